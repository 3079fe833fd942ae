"""CPU restatement of the reference's detection-level metrics (test infrastructure).

The loops below follow the reference line by line (plain Python, small cases only) and keep its arithmetic:
``predict()`` returns ``defect_position`` as a numpy float32 array and the targets go through
``tensor.cpu().numpy()``, so every IoU operation is a float32 operation under NumPy >= 2.
"""
from __future__ import annotations

import numpy as np


def _iou_f32(ps, pe, ts, te):
    """two_stage_train.py:329-336 / train.py:325-331 on numpy float32 scalars."""
    ps, pe, ts, te = np.float32(ps), np.float32(pe), np.float32(ts), np.float32(te)
    intersection = max(0, min(pe, te) - max(ps, ts))
    union = max(pe, te) - min(ps, ts)
    if union > 0:
        return np.float32(intersection / union)
    return None


def targets_from_dense(label, pos):
    """Dense targets (label [B,N] int, pos [B,N,2] float32) -> the per-set lists the reference builds
    (two_stage_train.py:249-262: one entry per A-scan with label > 0, 'position' = its index)."""
    out = []
    for b in range(label.shape[0]):
        out.append([{"position": i, "class": int(label[b, i]), "defect_position": pos[b, i].astype(np.float32)}
                    for i in range(label.shape[1]) if label[b, i] > 0])
    return out


def records_to_predictions(rec, B):
    """paut_detection records -> the per-set lists predict() returns (fields used by the metrics only)."""
    out = [[] for _ in range(B)]
    for r in rec:
        out[int(r["set_index"])].append({"position": int(r["position"]), "class": int(r["cls"]),
                                         "defect_position": np.array([r["start"], r["end"]], np.float32)})
    return out


def match_same_position(predictions, targets, iou_threshold=0.5):
    """two_stage_train.py:284-375 calculate_metrics."""
    tp = fp = fn = 0
    errors = []
    for batch_idx in range(len(predictions)):
        batch_targets = targets[batch_idx] if batch_idx < len(targets) else []
        matched = set()
        for pred in predictions[batch_idx]:
            best_iou, best_idx = 0, -1
            for ti, t in enumerate(batch_targets):
                if ti in matched or pred["position"] != t["position"]:
                    continue
                iou = _iou_f32(*pred["defect_position"], *t["defect_position"])
                if iou is not None and iou > best_iou:
                    best_iou, best_idx = iou, ti
            if best_iou > iou_threshold:
                tp += 1
                matched.add(best_idx)
                ps, pe = pred["defect_position"]
                ts, te = batch_targets[best_idx]["defect_position"]
                errors.append((abs(np.float32(ps) - np.float32(ts)) + abs(np.float32(pe) - np.float32(te))) / 2)
            else:
                fp += 1
        fn += len(batch_targets) - len(matched)
    precision = tp / max(tp + fp, 1)
    recall = tp / max(tp + fn, 1)
    return dict(true_positives=tp, false_positives=fp, false_negatives=fn, precision=precision, recall=recall,
                f1_score=2 * precision * recall / max(precision + recall, 1e-8),
                sum_position_error=float(np.sum(np.asarray(errors, np.float64))) if errors else 0.0,
                mean_position_error=float(np.mean(errors)) if errors else 0)


def match_first_class(predictions, targets, iou_threshold=0.5):
    """train.py:279-361 calculate_metrics (the 0.5 is hard-coded there)."""
    tp = fp = fn = 0
    ious = []
    for preds, tgts in zip(predictions, targets):
        matched = set()
        for pred in preds:
            hit = False
            for i, t in enumerate(tgts):
                if i in matched:
                    continue
                if pred["class"] == t["class"]:
                    iou = _iou_f32(*pred["defect_position"], *t["defect_position"])
                    if iou is not None and iou > iou_threshold:
                        tp += 1
                        matched.add(i)
                        hit = True
                        ious.append(iou)
                        break
            if not hit:
                fp += 1
        fn += len(tgts) - len(matched)
    precision = tp / (tp + fp) if tp + fp > 0 else 0
    recall = tp / (tp + fn) if tp + fn > 0 else 0
    return dict(true_positives=tp, false_positives=fp, false_negatives=fn, precision=precision, recall=recall,
                f1=2 * precision * recall / (precision + recall) if precision + recall > 0 else 0,
                sum_iou=float(np.sum(np.asarray(ious, np.float64))) if ious else 0.0,
                mean_iou=float(np.mean(ious)) if ious else 0)


def confusion(prob, label, threshold=0.5, ge=True):
    """acc_metrics_hybrid_binary_dynamic_.py:73-94 (ge) / test_detection.py:77 (strict): fp32 tensor comparison."""
    p = np.asarray(prob, np.float32).reshape(-1)
    y = np.asarray(label, np.float32).reshape(-1) > np.float32(0.5)
    pred = (p >= np.float32(threshold)) if ge else (p > np.float32(threshold))
    return dict(TP=int((pred & y).sum()), FP=int((pred & ~y).sum()), FN=int((~pred & y).sum()),
                TN=int((~pred & ~y).sum()))


def difference_matrix(signals, predictions, threshold=0.5):
    """teststtt.py:54-69 for one set: signals [N,S] (the reference holds float64 rows from np.loadtxt),
    predictions [N] -> (reference_signal [S] or None, difference matrix [N,S]); float64 arithmetic."""
    signals = np.asarray(signals, np.float64)
    healthy = [s for p, s in zip(predictions, signals) if p < threshold]
    if not healthy:
        return None, np.zeros_like(signals)
    ref = np.mean(healthy, axis=0)
    rows = [np.abs(s - ref) if p >= threshold else np.zeros_like(s) for p, s in zip(predictions, signals)]
    return ref, np.array(rows)
