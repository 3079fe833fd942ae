"""Generate the golden fixtures by running the REFERENCE classes (build container only).

    python tests/golden/make_golden.py            # needs /root/reference

For every hot-path model it imports the reference class from /root/reference,
loads the synthetic weights of oracle/synth.py with a strict load_state_dict,
runs ``forward`` (and ``predict`` where the class has one) on synthetic inputs
and stores outputs + records under tests/golden/.  Inputs and weights are not
stored (they are regenerated from seeds; a sha256 of each input is kept).
``DefectDetectionModel`` lives in a script that trains at import, so only its
class source range (signals/MSC_Conv1D_training.py:50-89) is exec'd.

The windowing known-answer vectors come from the reference's own
``SignalSequencePreparation.create_beam_sequences`` (matplotlib stubbed) and
from the rule in json_dataset.py:84-103.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("PAUT_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)

from oracle import synth  # noqa: E402
from oracle.postprocess import DETECTION  # noqa: E402

# (kind, case name, ctor cfg, input spec)
CASES = [
    ("msc", "u2x300", dict(), dict(gen="uniform", B=2, N=300, S=320, seed=1234)),
    ("msc", "p1x37", dict(), dict(gen="paut", B=1, N=37, S=320, seed=7)),
    ("msc_n", "u2x300", dict(), dict(gen="uniform", B=2, N=300, S=320, seed=1234)),
    ("msc_n", "p1x37", dict(), dict(gen="paut", B=1, N=37, S=320, seed=7)),
    ("conv1d_msc", "u2x300", dict(), dict(gen="uniform", B=2, N=300, S=320, seed=1234, transpose=True)),
    ("conv1d_msc", "p1x37", dict(), dict(gen="paut", B=1, N=37, S=320, seed=7, transpose=True)),
    ("ssd", "u2x50", dict(), dict(gen="uniform", B=2, N=50, S=320, seed=1234)),
    ("ssd", "p3x37", dict(), dict(gen="paut", B=3, N=37, S=320, seed=7)),
    ("ssd", "c3s100", dict(num_classes=3, signal_length=100), dict(gen="uniform", B=2, N=50, S=100, seed=5)),
    ("enhanced", "u2x50", dict(), dict(gen="uniform", B=2, N=50, S=320, seed=1234)),
    ("enhanced", "p1x37", dict(), dict(gen="paut", B=1, N=37, S=320, seed=7)),
    ("enhanced", "c3s100", dict(num_classes=3, signal_length=100), dict(gen="uniform", B=2, N=50, S=100, seed=5)),
    ("two_stage", "u2x50", dict(), dict(gen="uniform", B=2, N=50, S=320, seed=1234)),
    ("two_stage", "p3x37", dict(), dict(gen="paut", B=3, N=37, S=320, seed=7)),
    ("two_stage", "s100", dict(signal_length=100), dict(gen="uniform", B=2, N=50, S=100, seed=5)),
]


def make_input(spec):
    if spec["gen"] == "uniform":
        x = synth.synth_uniform_sets(spec["B"], spec["N"], spec["S"], seed=spec["seed"])
    else:
        x = synth.synth_paut_sets(spec["B"], spec["N"], spec["S"], seed=spec["seed"], defect_frac=0.2)
    if spec.get("transpose"):
        x = np.ascontiguousarray(x.transpose(0, 2, 1))
    return x


def reference_classes():
    sys.path.insert(0, os.path.join(REF, "SignalSequenceDetection"))
    sys.path.insert(0, os.path.join(REF, "signals", "multisignalNN"))
    from model import SignalSequenceDetector
    from enhanced_model import EnhancedSignalSequenceDetector
    from two_stage_model import TwoStageDefectDetector
    from NN_models import MultiSignalClassifier, MultiSignalClassifier_N
    src = open(os.path.join(REF, "signals", "MSC_Conv1D_training.py")).read().split("\n")
    ns = {}
    exec("import torch\nimport torch.nn as nn\n" + "\n".join(src[49:89]), ns)
    return {
        "msc": lambda **c: MultiSignalClassifier(c.get("signal_length", 320), [128, 64, 32], 4),
        "msc_n": lambda **c: MultiSignalClassifier_N(c.get("signal_length", 320), [128, 64, 32], 4),
        "conv1d_msc": lambda **c: ns["DefectDetectionModel"](320, 300),
        "ssd": lambda **c: SignalSequenceDetector(**c),
        "enhanced": lambda **c: EnhancedSignalSequenceDetector(**c),
        "two_stage": lambda **c: TwoStageDefectDetector(c.get("signal_length", 320)),
    }


def records_to_array(kind, preds, S):
    rows = []
    for b, seq in enumerate(preds):
        for r in seq:
            start, end = r["defect_position"]              # numpy float32 scalars
            rec = np.zeros((), dtype=DETECTION)
            rec["set_index"], rec["position"] = b, r["position"]
            rec["start"], rec["end"] = start, end
            rec["start_index"] = int(start * S)            # predict.py:111-113 verbatim expression
            rec["end_index"] = int(end * S)
            if kind == "two_stage":
                rec["cls"] = 1
                rec["score"], rec["uncertainty"] = r["defect_prob"], r["defect_uncertainty"]
                rec["confidence"] = r["adjusted_confidence"]
            else:
                rec["cls"], rec["score"], rec["anomaly"] = r["class"], r["class_score"], r["anomaly_score"]
                if kind == "enhanced":
                    rec["uncertainty"] = r["class_uncertainty"]
                    rec["confidence"] = r["adjusted_confidence"]
                else:
                    rec["confidence"] = r["class_score"]
            rows.append(rec)
    return np.array(rows, dtype=DETECTION) if rows else np.zeros(0, dtype=DETECTION)


def flatten_outputs(kind, out):
    if kind in ("msc", "msc_n"):
        return {"defect_prob": out[0], "defect_start": out[1], "defect_end": out[2]}
    if kind == "conv1d_msc":
        return {"defect_prob": out}
    flat = {}
    for k, v in out.items():
        if v is None:
            continue
        if isinstance(v, (list, tuple)):
            flat[k] = torch.stack(list(v), dim=0)          # [layers, B, N, N]
        else:
            flat[k] = v
    return flat


def windowing_vectors():
    sys.modules.setdefault("matplotlib", types.ModuleType("matplotlib"))
    sys.modules.setdefault("matplotlib.pyplot", types.ModuleType("matplotlib.pyplot"))
    sys.path.insert(0, os.path.join(REF, "SignalSequenceDetection"))
    from dataset_preparation import SignalSequencePreparation
    vec = {"ssd": {}, "msc": {}}
    for n in (1, 30, 49, 50, 51, 75, 99, 100, 101, 120, 149, 150, 151, 300, 301, 1000):
        prep = SignalSequencePreparation.__new__(SignalSequencePreparation)
        prep.seq_length = 50
        prep.number_false_signals = 0
        sig = np.arange(1, n * 4 + 1, dtype=np.float32).reshape(n, 4)
        prep.all_sequences = {"f": {"k": sig}}
        prep.all_annotations = {"f": {"k": [{"bbox": [0.0, 1.0, 0.1, 0.2], "label": "d"}]}}
        seqs = prep.create_beam_sequences()
        out = []
        for s in seqs:
            start = int(s.get("start_idx", 0))
            valid = int(s.get("original_length", 50))
            assert np.array_equal(s["signals"][:valid], sig[start:start + valid])
            out.append([start, valid])
        vec["ssd"][str(n)] = out
        # json_dataset.py:51-52,84-103 (the loader needs files on disk; the rule is four lines)
        import math
        w = []
        if n >= 50:
            num = math.ceil(n / 50)
            for i in range(num):
                if i < num - 1:
                    a = i * 50
                else:
                    a = n - 50
                w.append([a, 50])
        vec["msc"][str(n)] = w
    # all-zero run is dropped (dataset_preparation.py:205)
    prep = SignalSequencePreparation.__new__(SignalSequencePreparation)
    prep.seq_length, prep.number_false_signals = 50, 0
    prep.all_sequences = {"f": {"k": np.zeros((60, 4), np.float32)}}
    prep.all_annotations = {"f": {"k": [{"bbox": [0.0, 1.0, 0.1, 0.2], "label": "d"}]}}
    vec["ssd_all_zero_dropped"] = len(prep.create_beam_sequences()) == 0
    return vec


def main():
    torch.manual_seed(0)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    classes = reference_classes()
    manifest = {}
    for kind, case, cfg, spec in CASES:
        model = classes[kind](**cfg)
        spec_cfg = {k: v for k, v in cfg.items() if k in ("num_classes", "signal_length")}
        sd = synth.synth_state_dict(kind, seed=0, **spec_cfg)
        ref_keys = list(model.state_dict().keys())
        assert ref_keys == list(sd.keys()), f"{kind}: state_dict key/order mismatch"
        for k, v in model.state_dict().items():
            assert tuple(v.shape) == tuple(sd[k].shape), (kind, k)
        model.load_state_dict(sd, strict=True)
        model.eval()
        manifest[f"{kind}:{json.dumps(spec_cfg, sort_keys=True)}"] = [[k, list(v.shape)] for k, v in sd.items()]
        x_np = make_input(spec)
        x = torch.from_numpy(x_np)
        with torch.no_grad():
            out = model(x)
        flat = {k: v.detach().numpy().astype(np.float32) for k, v in flatten_outputs(kind, out).items()}
        save = {"out__" + k: v for k, v in flat.items()}
        meta = dict(kind=kind, case=case, cfg=cfg, input=spec,
                    input_sha256=hashlib.sha256(x_np.tobytes()).hexdigest(),
                    torch=torch.__version__, numpy=np.__version__)
        if hasattr(model, "predict"):
            S = spec["S"]
            if kind == "two_stage":
                conf = flat["defect_probs"][..., 1].astype(np.float64) / (1.0 + flat["defect_uncertainty"][..., 1].astype(np.float64))
            elif kind == "enhanced":
                p = torch.softmax(torch.from_numpy(flat["class_preds"]), -1).numpy()
                conf = p.max(-1).astype(np.float64) / (1.0 + flat["class_uncertainty"].mean(-1))
            else:
                conf = torch.softmax(torch.from_numpy(flat["class_preds"]), -1).numpy().max(-1).astype(np.float64)
            thresholds = [0.5, float(np.median(conf)), float(np.quantile(conf, 0.9))]
            for ti, thr in enumerate(thresholds):
                preds = model.predict(x, threshold=thr)
                save[f"rec{ti}"] = records_to_array(kind, preds, S)
            save["thresholds"] = np.asarray(thresholds, dtype=np.float64)
        save["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
        path = os.path.join(HERE, f"{kind}__{case}.npz")
        np.savez_compressed(path, **save)
        print(f"wrote {path}: {[(k, v.shape) for k, v in flat.items()]}"
              + (f" records {[len(save[f'rec{i}']) for i in range(3)]}" if "rec0" in save else ""))
    with open(os.path.join(HERE, "state_manifest.json"), "w") as f:
        json.dump(manifest, f, indent=0, sort_keys=True)
    with open(os.path.join(HERE, "windowing.json"), "w") as f:
        json.dump(windowing_vectors(), f, sort_keys=True)
    print("ok")


if __name__ == "__main__":
    main()
