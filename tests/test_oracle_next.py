"""Pin the oracle's restatement of the SURVEY section-8 "next" rows f3 / f4 (difference matrix, detection metrics)
on known answers produced by the reference's own functions (tests/golden/make_golden.py --metrics)."""
import os

import numpy as np
import pytest

from oracle import metrics as omx
from oracle import postprocess as opp
from oracle import synth
from tests._golden import GOLDEN_DIR


@pytest.fixture(scope="module")
def vec():
    return np.load(os.path.join(GOLDEN_DIR, "metrics_vectors.npz"))


def test_match_same_position_equals_reference(vec):
    B = vec["label"].shape[0]
    preds = omx.records_to_predictions(vec["rec"], B)
    targets = omx.targets_from_dense(vec["label"], vec["tpos"])
    for thr, counts, err in zip(vec["rule0_thr"], vec["rule0_counts"], vec["rule0_mean_err"]):
        r = omx.match_same_position(preds, targets, float(thr))
        assert [r["true_positives"], r["false_positives"], r["false_negatives"]] == counts.tolist()
        assert r["mean_position_error"] == pytest.approx(float(err), rel=0, abs=0)      # same float32 np.mean


def test_match_first_class_equals_reference(vec):
    B = vec["label"].shape[0]
    preds = omx.records_to_predictions(vec["rec"], B)
    targets = omx.targets_from_dense(vec["label"], vec["tpos"])
    r = omx.match_first_class(preds, targets)
    assert [r["true_positives"], r["false_positives"], r["false_negatives"]] == vec["rule1_counts"].tolist()
    assert r["mean_iou"] == float(vec["rule1_mean_iou"])


def test_confusion_equals_reference(vec):
    for thr, counts in zip(vec["conf_thr"], vec["conf_counts"]):
        c = omx.confusion(vec["conf_prob"], vec["conf_label"], float(thr), ge=True)
        assert [c["TP"], c["FP"], c["FN"], c["TN"]] == counts.tolist()
    # the fp32 tensor comparison differs from an fp64 comparison exactly on the planted values
    p = vec["conf_prob"]
    assert (p >= np.float32(0.7)).sum() - (p.astype(np.float64) >= 0.7).sum() == 5


def test_difference_matrix_equals_reference(vec):
    x = synth.synth_paut_sets(3, 40, 320, seed=int(vec["diff_x_seed"]), defect_frac=0.3)
    for s in range(3):
        ref, diff = omx.difference_matrix(x[s], [float(v) for v in vec["diff_prob"][s]], 0.5)
        if vec["diff_healthy"][s] == 0:
            assert ref is None and not diff.any()
        else:
            np.testing.assert_array_equal(ref, vec["diff_ref"][s])
            np.testing.assert_array_equal(diff, vec["diff_mat"][s])


def test_keep_rules():
    s = np.array([0.7, 0.5, 0.4999, 0.70001], np.float32)
    assert opp.keep_rule(s, 0.7, "ge32").tolist() == [True, False, False, True]
    assert opp.keep_rule(s, 0.7, "ge64").tolist() == [False, False, False, True]    # float32(0.7) < 0.7
    assert opp.keep_rule(s, 0.5, "gt32").tolist() == [True, False, False, True]
    assert opp.keep_rule(s, 0.5, "ge64").tolist() == [True, True, False, True]


def test_detection_metrics_formulas_match_reference(vec):
    """The host-side summary of the device counts uses the reference's formulas and key names."""
    from defectdetection_viaobjectdetection_b200.runtime import detection_metrics
    B = vec["label"].shape[0]
    preds = omx.records_to_predictions(vec["rec"], B)
    targets = omx.targets_from_dense(vec["label"], vec["tpos"])
    r0 = omx.match_same_position(preds, targets, 0.5)
    got0 = detection_metrics("position", dict(tp=r0["true_positives"], fp=r0["false_positives"], fn=r0["false_negatives"],
                                              sum_position_error=r0["sum_position_error"], sum_iou=0.0))
    for k in ("precision", "recall", "f1_score", "true_positives", "false_positives", "false_negatives"):
        assert got0[k] == r0[k], k
    assert got0["mean_position_error"] == pytest.approx(r0["mean_position_error"], abs=1e-6)
    r1 = omx.match_first_class(preds, targets)
    got1 = detection_metrics("class", dict(tp=r1["true_positives"], fp=r1["false_positives"], fn=r1["false_negatives"],
                                           sum_iou=r1["sum_iou"], sum_position_error=0.0))
    for k in ("precision", "recall", "f1", "true_positives", "false_positives", "false_negatives"):
        assert got1[k] == r1[k], k
    assert got1["mean_iou"] == pytest.approx(r1["mean_iou"], abs=1e-6)
    empty = detection_metrics("class", dict(tp=0, fp=0, fn=0, sum_iou=0.0, sum_position_error=0.0))
    assert empty["precision"] == 0 and empty["f1"] == 0 and empty["mean_iou"] == 0
