"""CPU checks of the boundary: the C-ABI library loads and exports every symbol include/paut.h declares;
the drop-in modules expose the reference's state_dict contract; host-side window tables match the oracle."""
import ctypes
import json
import os
import re

import pytest
import torch

import defectdetection_viaobjectdetection_b200 as paut
from defectdetection_viaobjectdetection_b200 import _lib, contract
from oracle import synth, windowing
from tests._golden import GOLDEN_DIR

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "paut.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(paut_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    declared = _header_symbols()
    assert len(declared) >= 16
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/paut.h but not exported"
    assert sorted(_lib.SYMBOLS) == declared
    assert _lib.load().paut_abi_version() == 1


def test_header_is_valid_c(tmp_path):
    """include/paut.h is a C header: a C99 translation unit includes it, binds the entry points with dlsym and
    runs the host-only calls (window tables, JSON loader, the no-GPU failure of ctx_create)."""
    import subprocess
    exe = str(tmp_path / "abi_smoke")
    src = os.path.join(ROOT, "tests", "c_abi", "abi_smoke.c")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), src, "-o", exe,
                    "-ldl"], check=True)
    r = subprocess.run([exe, _lib.LIB_PATH, os.path.join(GOLDEN_DIR, "json_volume", "a.json")], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert "c abi ok" in r.stdout


def test_detection_record_layout():
    assert _lib.DETECTION.itemsize == 48
    assert _lib.DETECTION.fields["confidence"][1] == 40


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    lib = _lib.load()
    h = ctypes.c_void_p()
    assert lib.paut_ctx_create(0, None, ctypes.byref(h)) == -2          # PAUT_ERR_CUDA
    assert b"no CUDA device" in lib.paut_last_error(None)
    m = paut.TwoStageDefectDetector(320).eval()
    with pytest.raises(RuntimeError):
        m(torch.rand(1, 50, 320))


from defectdetection_viaobjectdetection_b200.modules import FACTORIES as MODELS  # noqa: E402


def test_modules_expose_reference_state_dict_contract():
    with open(os.path.join(GOLDEN_DIR, "state_manifest.json")) as f:
        manifest = json.load(f)
    for key, entries in manifest.items():
        kind, cfg = key.split(":", 1)
        cfg = json.loads(cfg)
        m = MODELS[kind](cfg)
        sd = m.state_dict()
        assert list(sd.keys()) == [k for k, _ in entries], kind
        for k, shape in entries:
            assert list(sd[k].shape) == shape, (kind, k)
        # a reference-format checkpoint loads strictly (both checkpoint formats)
        ref_sd = synth.synth_state_dict(kind, seed=0, **cfg)
        m.load_state_dict(paut.load_checkpoint_state({"model_state_dict": ref_sd, "epoch": 3}), strict=True)
        m.load_state_dict(paut.load_checkpoint_state(ref_sd), strict=True)


def test_contract_matches_oracle_spec():
    for kind in synth.KINDS:
        a = [(k, tuple(s)) for k, s, _ in contract.state_spec(kind)]
        b = [(k, tuple(v[0])) for k, v in synth.state_spec(kind).items()]
        assert a == b, kind


def test_constructor_errors_match_reference():
    with pytest.raises(AssertionError):
        paut.MultiSignalClassifier_N(320, [128, 64, 32], 9)          # training_01.py:116 num_heads=9
    with pytest.raises(AssertionError):
        paut.SignalSequenceDetector(d_model=128, nhead=7)
    with pytest.raises(AssertionError):
        paut.ImprovedMultiSignalClassifier(320, [128, 64, 32], num_heads=9)


def test_window_table_host_matches_oracle():
    for n in list(range(0, 160)) + [299, 300, 301, 1000, 160000]:
        assert paut.window_table("ssd", n, 50) == ([] if n == 0 else windowing.ssd_windows(n, 50)), n
        assert paut.window_table("msc", n, 50) == windowing.msc_windows(n, 50), n
    for n in (10, 299, 300, 301, 900, 1000):
        assert paut.window_table("msc", n, 300) == windowing.msc_windows(n, 300)
        assert paut.window_table("ssd", n, 300) == windowing.ssd_windows(n, 300)
