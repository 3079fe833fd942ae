"""Pin the CPU oracle (oracle/) on the golden vectors produced by the reference classes."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import models as om
from oracle import postprocess as opp
from oracle import synth, windowing
from tests._golden import GOLDEN_DIR, case_id, case_state_dict, flatten, golden_files, load_case

FILES = golden_files()


def test_golden_files_present():
    assert len(FILES) >= 26


@pytest.mark.parametrize("path", FILES, ids=case_id)
def test_oracle_forward_matches_reference(path):
    c = load_case(path)
    sd = case_state_dict(c)
    with torch.no_grad():
        out = om.FORWARD[c["kind"]](sd, torch.from_numpy(c["x"]))
    flat = flatten(c["kind"], out)
    assert set(flat) == set(c["outs"])
    for k, ref in c["outs"].items():
        assert flat[k].shape == ref.shape, k
        # same ATen primitives in a different composition: accumulation-order noise only
        np.testing.assert_allclose(flat[k], ref, rtol=0, atol=2e-5, err_msg=f"{c['kind']}:{k}")


@pytest.mark.parametrize("path", [p for p in FILES if np.load(p).files.count("rec0")], ids=case_id)
def test_oracle_postprocess_bit_exact_on_reference_outputs(path):
    """Integer stage: fed the reference's own forward outputs, the restated predict() must return
    the same records (flags, positions, integer sample indices bit-exact; fp64 confidence equal)."""
    c = load_case(path)
    outs = {k: v for k, v in c["outs"].items()}
    for thr, ref in zip(c["thresholds"], c["recs"]):
        got = opp.postprocess(c["kind"], outs, float(thr), c["S"])
        assert len(got) == len(ref)
        for f in ("set_index", "position", "cls", "start_index", "end_index"):
            np.testing.assert_array_equal(got[f], ref[f], err_msg=f)
        for f in ("start", "end", "uncertainty", "anomaly"):
            np.testing.assert_array_equal(got[f], ref[f], err_msg=f)
        # class_score goes through a 2- or 3-way fp32 softmax: allow 1 ulp between exp implementations
        np.testing.assert_allclose(got["score"], ref["score"], rtol=2e-7, atol=0)
        np.testing.assert_allclose(got["confidence"], ref["confidence"], rtol=2e-7, atol=0)


def test_state_manifest_matches_spec():
    with open(os.path.join(GOLDEN_DIR, "state_manifest.json")) as f:
        manifest = json.load(f)
    assert len(manifest) >= 11
    for key, shapes in manifest.items():
        kind, cfg = key.split(":", 1)
        spec = synth.state_spec(kind, **json.loads(cfg))
        assert list(spec.keys()) == [k for k, _ in shapes], kind      # same keys, same order
        for (k, (shape, _, _)), (_, ref_shape) in zip(spec.items(), shapes):
            assert list(shape) == ref_shape, (kind, k)


def test_windowing_known_answers():
    with open(os.path.join(GOLDEN_DIR, "windowing.json")) as f:
        vec = json.load(f)
    for n, ref in vec["ssd"].items():
        assert [list(w) for w in windowing.ssd_windows(int(n), 50)] == ref, n
    for n, ref in vec["msc"].items():
        assert [list(w) for w in windowing.msc_windows(int(n), 50)] == ref, n
    assert vec["ssd_all_zero_dropped"] is True
    # SURVEY.md §8(c) known answers
    assert [s for s, _ in windowing.ssd_windows(120)] == [0, 17, 34, 51, 68, 70]
    assert windowing.ssd_windows(30) == [(0, 30)]
    assert [s for s, _ in windowing.msc_windows(120)] == [0, 50, 70]
    assert windowing.msc_windows(49) == []


def test_gather_windows_semantics():
    vol = np.arange(2 * 120 * 4, dtype=np.float32).reshape(2, 120, 4) + 1
    vol[1] = 0
    sets, table = windowing.gather_windows(vol, "ssd")
    assert table[:, 0].tolist() == [0] * 6                       # all-zero run dropped
    np.testing.assert_array_equal(sets[5], vol[0, 70:120])
    short = np.ones((1, 30, 4), np.float32)
    sets, table = windowing.gather_windows(short, "ssd")
    assert sets.shape == (1, 50, 4) and sets[0, 30:].sum() == 0 and table[0].tolist() == [0, 0, 30]
    sets, table = windowing.gather_windows(vol, "msc")
    assert table.tolist() == [[0, 0, 50], [0, 50, 50], [0, 70, 50], [1, 0, 50], [1, 50, 50], [1, 70, 50]]


def test_sample_index_is_float32_product():
    # a value where fp32 and fp64 products truncate differently
    rng = np.random.default_rng(0)
    x = rng.random(2_000_000, dtype=np.float32)
    i32 = opp.sample_index(x, 320)
    i64 = np.trunc(x.astype(np.float64) * 320).astype(np.int32)
    assert (i32 != i64).any()
    ref = np.array([int(np.float32(v) * 320) for v in x[:5000]], dtype=np.int32)
    np.testing.assert_array_equal(i32[:5000], ref)


BF16_FILES = sorted(__import__("glob").glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "bf16", "*__*.npz")))


@pytest.mark.parametrize("path", [f for f in BF16_FILES if "p334" not in f], ids=lambda p: os.path.basename(p)[:-4])
def test_oracle_bf16_emulation_matches_reference_on_rounded_operands(path):
    """oracle precision='bf16' (fp32 math on bf16-rounded inputs and matrix weights) against what the REFERENCE
    classes produce on the same rounded operands (make_golden.py --bf16): the emulation is pinned, not assumed."""
    import hashlib
    import json
    z = np.load(path)
    meta = json.loads(bytes(z["meta"]).decode())
    spec = meta["input"]
    x = synth.synth_paut_sets(spec["B"], spec["N"], spec["S"], seed=spec["seed"], defect_frac=0.2)
    if spec.get("transpose"):
        x = np.ascontiguousarray(x.transpose(0, 2, 1))
    assert hashlib.sha256(x.tobytes()).hexdigest() == meta["input_sha256"]
    kind = meta["kind"]
    sd = synth.synth_state_dict(kind, seed=0)
    with torch.no_grad():
        got = flatten(kind, om.FORWARD[kind](sd, torch.from_numpy(x), precision="bf16"))
    for k in got:
        ref = z["out__" + k]
        assert np.abs(got[k] - ref).max() <= 5e-5, (kind, k, np.abs(got[k] - ref).max())
