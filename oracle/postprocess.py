"""CPU restatement of the reference post-processing (test infrastructure).

The reference ``predict()`` methods loop in Python over (set b, position i),
call ``.item()`` (fp32 -> Python float = fp64) and keep a record when a
confidence test passes; viewers then turn ``defect_position`` (numpy float32)
into integer sample indices with ``int(start * len(signal))``.  Under
NumPy >= 2 a float32 scalar times a Python int stays float32, so

    index = trunc_toward_zero( RN_fp32( start_fp32 * fp32(S) ) )

while the confidence arithmetic (division, comparison) is fp64.  Both are
restated here vectorised; ``tests/golden`` pins them on the reference's own
``predict()`` output.

Record layout (numpy structured dtype ``DETECTION``) mirrors ``paut_detection``
in include/paut.h.
"""
from __future__ import annotations

import numpy as np

DETECTION = np.dtype([
    ("set_index", "<i4"), ("position", "<i4"), ("cls", "<i4"),
    ("start_index", "<i4"), ("end_index", "<i4"),
    ("start", "<f4"), ("end", "<f4"),
    ("score", "<f4"), ("uncertainty", "<f4"), ("anomaly", "<f4"),
    ("confidence", "<f8"),
], align=True)
assert DETECTION.itemsize == 48


def sample_index(frac_f32, signal_length):
    """predict.py:111-113, signal_visualizer.py:409-410: int(start * len(signal)), float32 product."""
    prod = np.asarray(frac_f32, dtype=np.float32) * np.float32(signal_length)
    return np.trunc(prod).astype(np.int32)


def softmax_f32(logits):
    """F.softmax(class_preds[b, i], dim=0) on fp32 (model.py:448): exp(x - max) / sum."""
    x = np.asarray(logits, dtype=np.float32)
    e = np.exp(x - x.max(axis=-1, keepdims=True), dtype=np.float32)
    return (e / e.sum(axis=-1, keepdims=True, dtype=np.float32)).astype(np.float32)


def _emit(keep, cls, score, unc, anomaly, conf, pos, signal_length):
    b, i = np.nonzero(keep)                      # row-major == the reference's (b, i) loop order
    out = np.zeros(b.size, dtype=DETECTION)
    out["set_index"], out["position"] = b, i
    out["cls"] = cls[b, i]
    out["start"], out["end"] = pos[b, i, 0], pos[b, i, 1]
    out["start_index"] = sample_index(pos[b, i, 0], signal_length)
    out["end_index"] = sample_index(pos[b, i, 1], signal_length)
    out["score"], out["uncertainty"], out["anomaly"] = score[b, i], unc[b, i], anomaly[b, i]
    out["confidence"] = conf[b, i]
    return out


def ssd_postprocess(class_preds, position_preds, anomaly_scores, threshold, signal_length, class_probs=None):
    """SignalSequenceDetector.predict, model.py:446-473.
    keep iff class_score > thr or (pred_class > 0 and anomaly > thr), comparisons in fp64."""
    probs = softmax_f32(class_preds) if class_probs is None else np.asarray(class_probs, np.float32)
    cls = probs.argmax(axis=-1).astype(np.int32)                       # first maximum, like torch.argmax
    score = np.take_along_axis(probs, cls[..., None].astype(np.int64), axis=-1)[..., 0]
    an = np.asarray(anomaly_scores, np.float32)[..., 0]
    conf = score.astype(np.float64)
    keep = (conf > threshold) | ((cls > 0) & (an.astype(np.float64) > threshold))
    return _emit(keep, cls, score, np.zeros_like(score), an, conf,
                 np.asarray(position_preds, np.float32), signal_length)


def enhanced_postprocess(class_preds, class_uncertainty, position_preds, anomaly_scores, threshold,
                         signal_length, class_probs=None):
    """EnhancedSignalSequenceDetector.predict, enhanced_model.py:766-805.
    adjusted = class_score / (1.0 + class_unc[pred_class]) in fp64 (:789);
    keep iff adjusted > thr or (pred_class > 0 and anomaly > thr) (:792)."""
    probs = softmax_f32(class_preds) if class_probs is None else np.asarray(class_probs, np.float32)
    cls = probs.argmax(axis=-1).astype(np.int32)
    idx = cls[..., None].astype(np.int64)
    score = np.take_along_axis(probs, idx, axis=-1)[..., 0]
    unc = np.take_along_axis(np.asarray(class_uncertainty, np.float32), idx, axis=-1)[..., 0]
    an = np.asarray(anomaly_scores, np.float32)[..., 0]
    conf = score.astype(np.float64) / (1.0 + unc.astype(np.float64))
    keep = (conf > threshold) | ((cls > 0) & (an.astype(np.float64) > threshold))
    return _emit(keep, cls, score, unc, an, conf, np.asarray(position_preds, np.float32), signal_length)


def two_stage_postprocess(defect_probs, defect_uncertainty, position_preds, threshold, signal_length):
    """TwoStageDefectDetector.predict, two_stage_model.py:477-499.
    adjusted = defect_prob / (1.0 + defect_unc[..., 1]) in fp64 (:486); keep iff adjusted > thr (:489)."""
    score = np.asarray(defect_probs, np.float32)[..., 1]
    unc = np.asarray(defect_uncertainty, np.float32)[..., 1]
    conf = score.astype(np.float64) / (1.0 + unc.astype(np.float64))
    keep = conf > threshold
    cls = np.ones(score.shape, np.int32)
    return _emit(keep, cls, score, unc, np.zeros_like(score), conf,
                 np.asarray(position_preds, np.float32), signal_length)


def keep_rule(score_f32, threshold, cmp):
    """The four ways the reference thresholds a single probability:
    'gt64' model_pred.py:84 / evalMSC.py:91 (`.item()`-style fp64 `>`); 'ge64' improved_model.py:178-181,
    teststtt.py:60,66 (Python floats, `>=`); 'ge32' acc_metrics_hybrid_binary_dynamic_.py:84 and 'gt32'
    test_detection.py:77 (tensor comparison: the Python scalar is rounded to fp32 first)."""
    s = np.asarray(score_f32, np.float32)
    if cmp == "gt64":
        return s.astype(np.float64) > threshold
    if cmp == "ge64":
        return s.astype(np.float64) >= threshold
    if cmp == "ge32":
        return s >= np.float32(threshold)
    if cmp == "gt32":
        return s > np.float32(threshold)
    raise ValueError(cmp)


KEEP_RULE = {"msc": "gt64", "msc_n": "gt64", "conv1d_msc": "gt64", "msc_legacy": "ge64", "improved": "ge64",
             "hybrid": "ge32", "complex": "gt32"}


def msc_postprocess(defect_prob, defect_start, defect_end, threshold, signal_length, cmp="gt64"):
    """MSC callers: model_pred.py:82-85 / evalMSC.py:91 -- `prob > 0.5` (strict) marks the A-scan
    defective; start/end are the model's fractional positions.  Other single-probability models: keep_rule."""
    score = np.asarray(defect_prob, np.float32)
    conf = score.astype(np.float64)
    keep = keep_rule(score, threshold, cmp)
    pos = np.stack([np.asarray(defect_start, np.float32), np.asarray(defect_end, np.float32)], axis=-1)
    return _emit(keep, np.ones(score.shape, np.int32), score, np.zeros_like(score), np.zeros_like(score),
                 conf, pos, signal_length)


def postprocess(kind, outputs, threshold, signal_length):
    """Dispatch on model kind; ``outputs`` is what oracle.models.FORWARD[kind] returns."""
    def npy(t):
        return t.detach().cpu().numpy() if hasattr(t, "detach") else np.asarray(t)
    if kind in ("msc", "msc_n", "improved"):
        if isinstance(outputs, dict):
            outputs = (outputs["defect_prob"], outputs["defect_start"], outputs["defect_end"])
        return msc_postprocess(*(npy(o) for o in outputs), threshold, signal_length, KEEP_RULE[kind])
    if kind in ("conv1d_msc", "msc_legacy", "hybrid", "complex"):
        p = npy(outputs["defect_prob"] if isinstance(outputs, dict) else outputs)
        z = np.zeros_like(p)
        return msc_postprocess(p, z, z, threshold, signal_length, KEEP_RULE[kind])
    if kind == "ssd":
        return ssd_postprocess(npy(outputs["class_preds"]), npy(outputs["position_preds"]),
                               npy(outputs["anomaly_scores"]), threshold, signal_length)
    if kind == "enhanced":
        return enhanced_postprocess(npy(outputs["class_preds"]), npy(outputs["class_uncertainty"]),
                                    npy(outputs["position_preds"]), npy(outputs["anomaly_scores"]),
                                    threshold, signal_length)
    if kind == "two_stage":
        return two_stage_postprocess(npy(outputs["defect_probs"]), npy(outputs["defect_uncertainty"]),
                                     npy(outputs["position_preds"]), threshold, signal_length)
    raise ValueError(kind)
