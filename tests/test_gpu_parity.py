"""GPU parity: the CUDA path (through the drop-in modules -> C ABI -> sm_100a kernels) against the golden
vectors of the reference classes and against the CPU oracle on fresh seeded inputs."""
import ctypes as C

import numpy as np
import pytest
import torch

import defectdetection_viaobjectdetection_b200 as paut
from defectdetection_viaobjectdetection_b200 import runtime
from defectdetection_viaobjectdetection_b200._lib import Outputs
from oracle import models as om
from oracle import postprocess as opp
from oracle import synth, windowing
from tests._golden import case_id, case_state_dict, flatten, golden_files, load_case
from tests.test_abi import MODELS

pytestmark = pytest.mark.gpu
FILES = golden_files()
FP32_ATOL = 1e-4     # north_star: fp32 mode within 1e-4 absolute
BF16_ATOL = 1e-2     # north_star: bf16 I/O within 1e-2 absolute


def build(kind, cfg, precision="fp32", sd=None):
    m = MODELS[kind](cfg)
    m.load_state_dict(sd if sd is not None else synth.synth_state_dict(kind, seed=0, **cfg), strict=True)
    m = m.cuda().eval()
    m.precision = precision
    return m


def run_flat(m, kind, x):
    out = m(x)
    torch.cuda.synchronize()
    return flatten(kind, out)


@pytest.mark.parametrize("path", FILES, ids=case_id)
def test_forward_fp32_matches_reference_golden(path):
    c = load_case(path)
    m = build(c["kind"], c["cfg"], sd=case_state_dict(c))
    got = run_flat(m, c["kind"], torch.from_numpy(c["x"]).cuda())
    assert set(got) == set(c["outs"])
    for k, ref in c["outs"].items():
        assert got[k].shape == ref.shape, k
        err = np.abs(got[k] - ref).max()
        assert err <= FP32_ATOL, f"{c['kind']}:{k} max abs err {err:.3e}"


@pytest.mark.parametrize("kind,shape", [
    ("msc", (3, 300, 320)), ("msc_n", (2, 170, 320)), ("conv1d_msc", (2, 298, 320)),
    ("ssd", (5, 50, 320)), ("enhanced", (3, 50, 320)), ("two_stage", (7, 50, 320)),
    ("msc_legacy", (3, 298, 320)), ("msc_legacy", (2, 170, 360)), ("improved", (3, 300, 320)),
    ("hybrid", (2, 300, 320)), ("hybrid", (3, 50, 384)), ("complex", (2, 300, 320)), ("complex", (3, 61, 200)),
])
def test_forward_fp32_matches_oracle_fresh_inputs(kind, shape):
    B, N, S = shape
    x = synth.synth_paut_sets(B, N, S, seed=99, defect_frac=0.1)
    # hybrid / complex: the conv stack is length-agnostic (shared_layer.0 sees the 128 resampled values)
    cfg = dict(signal_length=S) if kind not in ("hybrid", "complex") else {}
    sd = synth.synth_state_dict(kind, seed=0, **cfg)
    xin = np.ascontiguousarray(x.transpose(0, 2, 1)) if kind == "conv1d_msc" else x
    with torch.no_grad():
        ref = flatten(kind, om.FORWARD[kind](sd, torch.from_numpy(xin)))
    m = build(kind, cfg)
    got = run_flat(m, kind, torch.from_numpy(xin).cuda())
    for k in ref:
        err = np.abs(got[k] - ref[k]).max()
        assert err <= FP32_ATOL, f"{kind}:{k} max abs err {err:.3e}"


@pytest.mark.parametrize("kind", ["msc", "two_stage", "enhanced"])
def test_chunked_equals_unchunked_bit_exact(kind):
    """Processing in resident chunks of whole sets must not change a single bit (sets are independent)."""
    N = 300 if kind == "msc" else 50
    B = 40 if kind == "msc" else 64
    x = torch.from_numpy(synth.synth_uniform_sets(B, N, 320, seed=3)).cuda()
    m = build(kind, dict(signal_length=320))
    ctx = paut.get_context(x.device)
    ctx.set_workspace_limit(64 << 30)
    full = run_flat(m, kind, x)
    ctx.set_workspace_limit(64 << 20)
    chunked = run_flat(m, kind, x)
    ctx.set_workspace_limit(4 << 30)
    for k in full:
        assert np.array_equal(full[k], chunked[k]), k
    # and a shard of the batch equals the same rows of the whole batch
    part = run_flat(m, kind, x[B // 2:].contiguous())
    for k in full:
        ref = full[k][:, B // 2:] if k == "attention_weights" and kind == "enhanced" else full[k][B // 2:]
        assert np.array_equal(part[k], ref), k


def _struct_from(kind, outs, device):
    """paut_outputs filled from given numpy arrays (to feed the reference's own forward outputs)."""
    s = Outputs()
    keep = []
    for i, (name, _) in enumerate(runtime.OUTPUT_SLOTS[kind]):
        if name in outs:
            t = torch.from_numpy(np.ascontiguousarray(outs[name])).to(device)
            keep.append(t)
            s.slot[i] = t.data_ptr()
    return s, keep


@pytest.mark.parametrize("path", [p for p in FILES if "rec0" in np.load(p).files], ids=case_id)
def test_postprocess_bit_exact_on_reference_outputs(path):
    """Integer stage fed the reference's forward outputs: records identical to the reference predict()."""
    c = load_case(path)
    m = build(c["kind"], c["cfg"], sd=case_state_dict(c))
    x = torch.from_numpy(c["x"]).cuda()
    native = m._native_for(x)
    B, N, S = c["x"].shape
    struct, keep = _struct_from(c["kind"], c["outs"], x.device)
    for thr, ref in zip(c["thresholds"], c["recs"]):
        det, count = native.postprocess(struct, B, N, S, float(thr), x.device)
        got = runtime.records_to_numpy(det, count)
        assert len(got) == len(ref)
        for f in ("set_index", "position", "cls", "start_index", "end_index", "start", "end", "uncertainty",
                  "anomaly"):
            np.testing.assert_array_equal(got[f], ref[f], err_msg=f)
        np.testing.assert_allclose(got["score"], ref["score"], rtol=2e-7, atol=0)
        np.testing.assert_allclose(got["confidence"], ref["confidence"], rtol=2e-7, atol=0)


def test_postprocess_msc_and_large_batch_order():
    """MSC threshold rule + ordering/compaction over many sets against the oracle."""
    rng = np.random.default_rng(5)
    B, N, S = 3000, 300, 320
    prob = rng.random((B, N), dtype=np.float32)
    start = rng.random((B, N), dtype=np.float32)
    end = rng.random((B, N), dtype=np.float32)
    ref = opp.msc_postprocess(prob, start, end, 0.9, S)
    m = build("msc", dict(signal_length=S))
    native = m._native_for(torch.zeros(1, device="cuda"))
    struct, keep = _struct_from("msc", dict(defect_prob=prob, defect_start=start, defect_end=end), "cuda")
    det, count = native.postprocess(struct, B, N, S, 0.9, torch.device("cuda"))
    got = runtime.records_to_numpy(det, count)
    assert len(got) == len(ref) and len(ref) > 50000
    for f in ("set_index", "position", "start_index", "end_index", "start", "end", "score"):
        np.testing.assert_array_equal(got[f], ref[f], err_msg=f)


@pytest.mark.parametrize("kind", ["ssd", "enhanced", "two_stage"])
def test_predict_end_to_end_flag_agreement(kind):
    """predict() through the CUDA forward: >= 99.9 % keep/skip agreement with the oracle, positions exact
    wherever the forward outputs round the same way."""
    B, N, S = 40, 50, 320
    x = synth.synth_paut_sets(B, N, S, seed=21, defect_frac=0.1)
    sd = synth.synth_state_dict(kind, seed=0, signal_length=S)
    with torch.no_grad():
        ref_out = om.FORWARD[kind](sd, torch.from_numpy(x))
    flat = flatten(kind, ref_out)
    if kind == "two_stage":
        conf = flat["defect_probs"][..., 1] / (1.0 + flat["defect_uncertainty"][..., 1])
    else:
        conf = opp.softmax_f32(flat["class_preds"]).max(-1)
    thr = float(np.median(conf))
    ref = opp.postprocess(kind, ref_out, thr, S)
    m = build(kind, dict(signal_length=S))
    preds = m.predict(torch.from_numpy(x).cuda(), threshold=thr)
    got_keys = {(b, r["position"]) for b, rows in enumerate(preds) for r in rows}
    ref_keys = {(int(r["set_index"]), int(r["position"])) for r in ref}
    disagree = len(got_keys ^ ref_keys)
    assert disagree <= 0.001 * B * N, f"{disagree} of {B * N} flags differ"
    assert len(preds) == B and all(isinstance(rows, list) for rows in preds)


def test_window_gather_matches_oracle():
    rng = np.random.default_rng(11)
    for rule, n, L in (("ssd", 120, 50), ("ssd", 30, 50), ("msc", 120, 50), ("msc", 1000, 300), ("ssd", 50, 50)):
        vol = rng.random((5, n, 320), dtype=np.float32)
        vol[3] = 0
        keep = ~np.all(vol.reshape(5, -1) == 0, axis=1) if rule == "ssd" else None
        ref_sets, ref_table = windowing.gather_windows(vol, rule, L)
        sets, table = paut.gather_windows(torch.from_numpy(vol).cuda(), rule, L, keep_groups=keep)
        np.testing.assert_array_equal(table, ref_table)
        np.testing.assert_array_equal(sets.cpu().numpy(), ref_sets)
        sets16, _ = paut.gather_windows(torch.from_numpy(vol).cuda(), rule, L, out_dtype=torch.bfloat16,
                                        keep_groups=keep)
        np.testing.assert_array_equal(sets16.float().cpu().numpy(),
                                      torch.from_numpy(ref_sets).to(torch.bfloat16).float().numpy())


def test_error_behaviour_matches_reference():
    m = build("msc", dict(signal_length=320))
    with pytest.raises(RuntimeError):                       # view() on a non-contiguous input raises (NN_models.py:111)
        m(torch.rand(2, 320, 300, device="cuda").transpose(1, 2))
    with pytest.raises(RuntimeError):                       # more than 300 signals: position table too short
        m(torch.rand(1, 301, 320, device="cuda"))
    with pytest.raises(NotImplementedError):
        build("two_stage", dict(signal_length=320))(torch.rand(1, 50, 320, device="cuda"), targets=[{}])


@pytest.mark.parametrize("kind,precision", [("two_stage", "fp32"), ("msc", "bf16")])
def test_volume_scanner_equals_resident_run(kind, precision):
    """Streaming from host memory in chunks on two streams returns exactly the records of one resident call."""
    from defectdetection_viaobjectdetection_b200.streaming import VolumeScanner, to_predictions
    N = 300 if kind == "msc" else 50
    B = 37
    x = torch.from_numpy(synth.synth_paut_sets(B, N, 320, seed=8, defect_frac=0.1))
    if precision == "bf16":
        x = x.to(torch.bfloat16)
    m = build(kind, dict(signal_length=320), precision=precision)
    everything = m.predict_records(x.cuda(), threshold=-1.0)
    thr = float(np.median(everything["confidence"]))          # keeps about half of the A-scans
    whole = m.predict_records(x.cuda(), threshold=thr)
    scanner = VolumeScanner(m, chunk_sets=8)
    reusing = VolumeScanner(m, chunk_sets=8, reuse_output=True)
    first = reusing.scan(x, threshold=thr).copy()
    again = reusing.scan(x, threshold=thr)
    for f in whole.dtype.names:                               # the arena of the previous scan is not clobbered
        np.testing.assert_array_equal(first[f], whole[f], err_msg=f)
        np.testing.assert_array_equal(again[f], whole[f], err_msg=f)
    for host in (x.pin_memory(), x):
        got = scanner.scan(host, threshold=thr)
        assert len(got) == len(whole) > 0
        for f in whole.dtype.names:
            np.testing.assert_array_equal(got[f], whole[f], err_msg=f)
    assert scanner.h2d_bytes == x.numel() * x.element_size()
    preds = to_predictions(got, B)
    assert sum(len(p) for p in preds) == len(got)
    # lifetime: close() destroys the lanes' contexts (workspaces outside torch's allocator) and the packed weights made
    # for them; the scanner works again afterwards
    before = len(runtime._contexts)
    natives = len(m._native)
    scanner.close()
    assert len(runtime._contexts) == before - scanner.num_lanes
    assert len(m._native) == natives - scanner.num_lanes
    with scanner:
        again2 = scanner.scan(x.pin_memory(), threshold=thr)
        for f in whole.dtype.names:
            np.testing.assert_array_equal(again2[f], whole[f], err_msg=f)
    assert len(runtime._contexts) == before - scanner.num_lanes


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_shapes_alternate_across_streams(precision):
    """The opt-in to more than 48 KB of dynamic shared memory is a property of (device, kernel), not of a context
    (round-1 advice): a large-shape model on the default stream, a small-shape model on a second stream (its own
    context and workspace), the large shape again -- every launch must succeed and reproduce its first result."""
    big = build("msc", dict(signal_length=320), precision=precision)
    small = build("two_stage", dict(signal_length=320), precision=precision)
    xb = torch.from_numpy(synth.synth_paut_sets(3, 300, 320, seed=5, defect_frac=0.1)).cuda()
    xs = torch.from_numpy(synth.synth_paut_sets(2, 50, 320, seed=6, defect_frac=0.1)).cuda()
    if precision == "bf16":
        xb, xs = xb.to(torch.bfloat16), xs.to(torch.bfloat16)
    first_big = run_flat(big, "msc", xb)
    first_small = run_flat(small, "two_stage", xs)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    for _ in range(2):
        with torch.cuda.stream(side):
            on_side_small = run_flat(small, "two_stage", xs)
            on_side_big = run_flat(big, "msc", xb[:1])
        again_big = run_flat(big, "msc", xb)
        again_small = run_flat(small, "two_stage", xs)
        for k in first_big:
            np.testing.assert_array_equal(again_big[k], first_big[k], err_msg=k)
            np.testing.assert_array_equal(on_side_big[k], first_big[k][:1], err_msg=k)
        for k in first_small:
            np.testing.assert_array_equal(again_small[k], first_small[k], err_msg=k)
            np.testing.assert_array_equal(on_side_small[k], first_small[k], err_msg=k)


@pytest.mark.parametrize("n_sets,kind", [(3334, "msc"), (26667, "msc"), (20000, "two_stage")])
def test_full_size_volume_invariants(n_sets, kind):
    """BASELINE.json's full sizes (1 M A-scans of configs[1] / configs[3], the 8 M-A-scan whole-weld volume of
    configs[4]) through size-independent properties: the volume is a 256-set block repeated, so every repetition
    must reproduce the block's outputs bit for bit wherever it falls relative to the resident chunks (sets are
    independent); records of the whole volume == two half-volume shards concatenated; and, at 1 M A-scans, == the
    records streamed from host memory by VolumeScanner.  The block itself is checked against the oracle."""
    from defectdetection_viaobjectdetection_b200.streaming import VolumeScanner
    N = 300 if kind == "msc" else 50
    block = synth.synth_paut_sets(256, N, 320, seed=17, defect_frac=0.05)
    sd = synth.synth_state_dict(kind, seed=0)
    xb = torch.from_numpy(block).to(torch.bfloat16)
    reps = (n_sets + 255) // 256
    x = xb.cuda().repeat(reps, 1, 1)[:n_sets].contiguous()
    m = build(kind, dict(signal_length=320), precision="bf16")
    key = "defect_prob" if kind == "msc" else "defect_logits"
    out = run_flat(m, kind, x)[key]
    head = out[:256]
    with torch.no_grad():
        ref = flatten(kind, om.FORWARD[kind](sd, xb[:16].float(), precision="bf16"))[key]
    assert np.abs(head[:16] - ref).max() <= BF16_ATOL
    for k in range(1, reps):
        part = out[k * 256:(k + 1) * 256]
        assert np.array_equal(part, head[:len(part)]), f"repetition {k} differs from the first block"
    thr = float(np.median(m.predict_records(x[:256].contiguous(), threshold=-1.0)["confidence"]))   # keeps about half
    whole = m.predict_records(x, threshold=thr)
    h = (n_sets // 2) // 256 * 256                              # shard boundary on a block (and chunk-alignment) boundary
    lo, hi = m.predict_records(x[:h].contiguous(), thr), m.predict_records(x[h:].contiguous(), thr)
    hi["set_index"] += h
    both = np.concatenate([lo, hi])
    assert len(both) == len(whole) > 0
    for f in whole.dtype.names:
        np.testing.assert_array_equal(both[f], whole[f], err_msg=f)
    if n_sets <= 3334:
        got = VolumeScanner(m, chunk_sets=128).scan(x.cpu(), threshold=thr)
        for f in whole.dtype.names:
            np.testing.assert_array_equal(got[f], whole[f], err_msg=f)
