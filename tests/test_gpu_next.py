"""GPU parity of the SURVEY section-8 "next" rows: f3 (legacy MSC with the shipped trained weights, difference
matrix), f4 (detection-level metrics) and the keep rules of the f2 models -- all through the C ABI."""
import os

import numpy as np
import pytest
import torch

import defectdetection_viaobjectdetection_b200 as paut
from defectdetection_viaobjectdetection_b200 import runtime
from oracle import metrics as omx
from oracle import postprocess as opp
from oracle import synth
from tests._golden import GOLDEN_DIR, case_state_dict, load_case
from tests.test_gpu_parity import _struct_from, build

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def vec():
    return np.load(os.path.join(GOLDEN_DIR, "metrics_vectors.npz"))


def _records_on_device(rec):
    det = torch.from_numpy(rec.view(np.uint8).copy()).cuda()
    count = torch.tensor([len(rec)], dtype=torch.int32, device="cuda")
    return det, count


def test_metrics_match_reference_vectors(vec):
    """Integer counts bit-exact with the reference's calculate_metrics functions; fp64 sums to rounding."""
    rec, label, tpos = vec["rec"], vec["label"], vec["tpos"]
    B, N = label.shape
    det, count = _records_on_device(rec)
    lab_d, pos_d = torch.from_numpy(label).cuda(), torch.from_numpy(tpos).cuda()
    for thr, counts, err in zip(vec["rule0_thr"], vec["rule0_counts"], vec["rule0_mean_err"]):
        m = paut.metrics_match("position", det, count, B, N, lab_d, pos_d, float(thr))
        assert [m["tp"], m["fp"], m["fn"]] == counts.tolist()
        assert m["sum_position_error"] / max(m["tp"], 1) == pytest.approx(float(err), abs=1e-6)
    m = paut.metrics_match("class", det, count, B, N, lab_d, pos_d, 0.5)
    assert [m["tp"], m["fp"], m["fn"]] == vec["rule1_counts"].tolist()
    assert m["sum_iou"] / max(m["tp"], 1) == pytest.approx(float(vec["rule1_mean_iou"]), abs=1e-6)


def test_metrics_match_oracle_large_random():
    rng = np.random.default_rng(31)
    B, N = 4000, 50
    label = (rng.random((B, N)) < 0.3).astype(np.int32) * rng.integers(1, 4, size=(B, N)).astype(np.int32)
    tpos = np.sort(rng.random((B, N, 2), dtype=np.float32), axis=-1)
    ppos = (tpos + rng.normal(0, 0.1, size=(B, N, 2)).astype(np.float32)).astype(np.float32)
    keep = rng.random((B, N)) < 0.3
    keep[5] = False                                           # a set without predictions
    keep[6] = True
    b, i = np.nonzero(keep)
    rec = np.zeros(b.size, dtype=opp.DETECTION)
    rec["set_index"], rec["position"] = b, i
    rec["cls"] = rng.integers(1, 4, size=b.size)
    rec["start"], rec["end"] = ppos[b, i, 0], ppos[b, i, 1]
    det, count = _records_on_device(rec)
    lab_d, pos_d = torch.from_numpy(label).cuda(), torch.from_numpy(tpos).cuda()
    sub = 300                                                 # the Python oracle loops: check a prefix exactly ...
    preds, targets = omx.records_to_predictions(rec[rec["set_index"] < sub], sub), omx.targets_from_dense(label[:sub], tpos[:sub])
    det_s, count_s = _records_on_device(rec[rec["set_index"] < sub])
    for rule, fn in (("position", omx.match_same_position), ("class", omx.match_first_class)):
        for thr in (0.5, 0.25):
            r = fn(preds, targets, thr)
            m = paut.metrics_match(rule, det_s, count_s, sub, N, lab_d[:sub].contiguous(), pos_d[:sub].contiguous(), thr)
            assert [m["tp"], m["fp"], m["fn"]] == [r["true_positives"], r["false_positives"], r["false_negatives"]]
            key = "sum_position_error" if rule == "position" else "sum_iou"
            assert m[key] == pytest.approx(r[key], rel=1e-12)
    # ... and the whole volume through size-independent identities: tp + fp = predictions, tp + fn = targets,
    # and the sum over two halves equals the whole (sets are independent)
    whole = paut.metrics_match("class", det, count, B, N, lab_d, pos_d, 0.5)
    assert whole["tp"] + whole["fp"] == len(rec) and whole["tp"] + whole["fn"] == int((label > 0).sum())
    h = B // 2
    lo, hi = rec[rec["set_index"] < h], rec[rec["set_index"] >= h].copy()
    hi["set_index"] -= h
    parts = [paut.metrics_match("class", *_records_on_device(r), h, N, lab_d[s].contiguous(), pos_d[s].contiguous(), 0.5)
             for r, s in ((lo, slice(0, h)), (hi, slice(h, B)))]
    for k in ("tp", "fp", "fn"):
        assert parts[0][k] + parts[1][k] == whole[k]


def test_metrics_confusion(vec):
    prob, lab = torch.from_numpy(vec["conf_prob"]).cuda(), torch.from_numpy(vec["conf_label"]).cuda()
    for thr, counts in zip(vec["conf_thr"], vec["conf_counts"]):
        m = paut.metrics_confusion(prob, lab, float(thr), ge=True)
        assert [m["tp"], m["fp"], m["fn"], m["tn"]] == counts.tolist()
    rng = np.random.default_rng(2)
    p = rng.random(5_000_001, dtype=np.float32)
    y = (rng.random(5_000_001) < 0.2).astype(np.float32)
    for ge in (True, False):
        m = paut.metrics_confusion(torch.from_numpy(p).cuda(), torch.from_numpy(y).cuda(), 0.3, ge=ge)
        c = omx.confusion(p, y, 0.3, ge=ge)
        assert [m["tp"], m["fp"], m["fn"], m["tn"]] == [c["TP"], c["FP"], c["FN"], c["TN"]]


def test_difference_matrix_bit_exact(vec):
    x = synth.synth_paut_sets(3, 40, 320, seed=int(vec["diff_x_seed"]), defect_frac=0.3)
    prob = torch.from_numpy(vec["diff_prob"]).cuda()
    ref, diff, healthy = paut.difference_matrix(torch.from_numpy(x).cuda(), prob, 0.5)
    np.testing.assert_array_equal(healthy.cpu().numpy(), vec["diff_healthy"])
    # the reference computes in float64; the library returns the float32 rounding of the same float64 values
    np.testing.assert_array_equal(ref.cpu().numpy(), vec["diff_ref"].astype(np.float32))
    np.testing.assert_array_equal(diff.cpu().numpy(), vec["diff_mat"].astype(np.float32))
    # bf16 volume: same rule on the bf16-rounded samples
    xb = torch.from_numpy(x).to(torch.bfloat16)
    ref_b, diff_b, _ = paut.difference_matrix(xb.cuda(), prob, 0.5)
    for s in range(2):
        r, d = omx.difference_matrix(xb[s].float().numpy(), [float(v) for v in vec["diff_prob"][s]], 0.5)
        np.testing.assert_array_equal(ref_b[s].cpu().numpy(), r.astype(np.float32))
        np.testing.assert_array_equal(diff_b[s].cpu().numpy(), d.astype(np.float32))


@pytest.mark.parametrize("kind", ["msc_legacy", "improved", "hybrid", "complex"])
def test_keep_rules_on_device(kind):
    """Each single-probability model thresholds the way its reference caller does (include/paut.h "Keep rule")."""
    rng = np.random.default_rng(9)
    B, N, S = 64, 300, 320
    prob = rng.random((B, N), dtype=np.float32)
    prob[:, :7] = np.float32(0.7)                               # float32(0.7) < 0.7: fp32 and fp64 rules disagree here
    prob[:, 7:9] = np.float32(0.5)
    start, end = rng.random((B, N), dtype=np.float32), rng.random((B, N), dtype=np.float32)
    m = build(kind, {})
    native = m._native_for(torch.zeros(1, device="cuda"))
    outs = dict(defect_prob=prob, defect_start=start, defect_end=end) if kind == "improved" else dict(defect_prob=prob)
    struct, keep = _struct_from(kind, outs, "cuda")
    for thr in (0.7, 0.5, 0.123):
        ref = opp.postprocess(kind, outs, thr, S)
        det, count = native.postprocess(struct, B, N, S, thr, torch.device("cuda"))
        got = runtime.records_to_numpy(det, count)
        assert len(got) == len(ref) > 0
        for f in ("set_index", "position", "start_index", "end_index", "start", "end", "score", "confidence"):
            np.testing.assert_array_equal(got[f], ref[f], err_msg=f)


def test_legacy_real_weights_end_to_end():
    """The checkpoint the reference ships: forward in both modes, predict records, difference matrix."""
    c = load_case(os.path.join(GOLDEN_DIR, "msc_legacy__realFPDp2x170.npz"))
    sd = case_state_dict(c)
    m = build("msc_legacy", c["cfg"], sd=sd)
    x = torch.from_numpy(c["x"]).cuda()
    prob, ref, diff, healthy = m.difference_matrix(x, 0.5)
    assert np.abs(prob.cpu().numpy() - c["outs"]["defect_prob"]).max() <= 1e-4
    p = prob.cpu().numpy()
    for s in range(2):
        r, d = omx.difference_matrix(c["x"][s], [float(v) for v in p[s]], 0.5)
        assert int(healthy[s]) == int((p[s].astype(np.float64) < 0.5).sum())
        if r is None:                                          # no healthy A-scan in the set: the reference skips it
            assert int(healthy[s]) == 0 and not ref[s].any() and not diff[s].any()
            continue
        np.testing.assert_array_equal(ref[s].cpu().numpy(), r.astype(np.float32))
        np.testing.assert_array_equal(diff[s].cpu().numpy(), d.astype(np.float32))
    assert int(healthy.sum()) > 0
    pmap = m.prediction_map(x, 100)                             # GNN_testing_multi_v2_MAP.py:38-67: first 100 signals per folder
    with torch.no_grad():
        from oracle import models as om
        want_map = om.msc_legacy_forward(sd, torch.from_numpy(c["x"][:, :100]))
    assert pmap.shape == (2, 100) and np.abs(pmap.cpu().numpy() - want_map.numpy()).max() <= 1e-4
    rec = m.predict_records(x, 0.5)
    want = opp.postprocess("msc_legacy", p, 0.5, 360)
    np.testing.assert_array_equal(rec["position"], want["position"])
    m16 = build("msc_legacy", c["cfg"], precision="bf16", sd=sd)
    p16 = m16(x.to(torch.bfloat16)).cpu().numpy()
    assert np.abs(p16 - c["outs"]["defect_prob"]).max() <= 2e-2


def test_all_zero_run_drop_on_device():
    """dataset_preparation.py:205 (np.all(signals == 0)) as a device reduction, then the SSD windowing."""
    from oracle import windowing
    rng = np.random.default_rng(3)
    vol = rng.random((9, 120, 320), dtype=np.float32)
    vol[2] = 0
    vol[5] = -0.0                                    # -0.0 == 0
    vol[7] = 0
    vol[7, 119, 319] = 1e-30                         # one non-zero sample at the very end keeps the run
    vol[8] = 0
    vol[8, 0, 0] = np.nan                            # nan != 0
    want = ~np.all(vol.reshape(9, -1) == 0, axis=1)
    for dt in (torch.float32, torch.bfloat16):
        v = torch.from_numpy(vol).to(dt)
        want_dt = ~np.all(v.float().numpy().reshape(9, -1) == 0, axis=1)
        np.testing.assert_array_equal(paut.group_nonzero(v.cuda()), want_dt)
    ref_sets, ref_table = windowing.gather_windows(vol, "ssd", 50)
    sets, table = paut.gather_windows(torch.from_numpy(vol).cuda(), "ssd", 50, drop_all_zero=True)
    np.testing.assert_array_equal(table, ref_table)
    np.testing.assert_array_equal(sets.cpu().numpy(), ref_sets)
    assert want.sum() == 7
