"""Helpers to load the committed golden fixtures (tests/golden/*.npz) and rebuild their inputs."""
import glob
import hashlib
import json
import os

import numpy as np

from oracle import synth

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_files():
    return sorted(glob.glob(os.path.join(GOLDEN_DIR, "*__*.npz")))


def case_id(path):
    return os.path.basename(path)[:-4]


def load_case(path):
    z = np.load(path)
    meta = json.loads(bytes(z["meta"]).decode())
    spec = meta["input"]
    if spec["gen"] == "uniform":
        x = synth.synth_uniform_sets(spec["B"], spec["N"], spec["S"], seed=spec["seed"])
    else:
        x = synth.synth_paut_sets(spec["B"], spec["N"], spec["S"], seed=spec["seed"], defect_frac=0.2)
    if spec.get("transpose"):
        x = np.ascontiguousarray(x.transpose(0, 2, 1))
    if "scale" in spec:
        x = (x * np.float32(spec["scale"])).astype(np.float32)
    assert hashlib.sha256(x.tobytes()).hexdigest() == meta["input_sha256"], "synthetic input drifted"
    cfg = {k: v for k, v in meta["cfg"].items() if k in ("num_classes", "signal_length", "hidden_sizes", "d_model",
                                                          "num_layers")}
    weights = meta["cfg"].get("weights")
    outs = {k[5:]: z[k] for k in z.files if k.startswith("out__")}
    recs = [z[f"rec{i}"] for i in range(3)] if "rec0" in z.files else None
    thr = z["thresholds"] if "thresholds" in z.files else None
    return dict(kind=meta["kind"], case=meta["case"], cfg=cfg, x=x, outs=outs, recs=recs, thresholds=thr,
                S=spec["S"], meta=meta, weights=weights)


def case_state_dict(c):
    """Weights of a golden case: synthetic (seeded), or the reference's shipped checkpoint stored as a fixture."""
    import torch
    if c.get("weights"):
        z = np.load(os.path.join(GOLDEN_DIR, "weights_" + c["weights"][:-4] + ".npz"))
        return {k: torch.from_numpy(z[k]) for k in z.files}
    return synth.synth_state_dict(c["kind"], seed=0, **c["cfg"])


def flatten(kind, out):
    """Oracle / library outputs -> {name: np.ndarray} with the golden naming."""
    def npy(t):
        return t.detach().cpu().float().numpy() if hasattr(t, "detach") else np.asarray(t)
    if kind in ("msc", "msc_n", "improved"):
        return {"defect_prob": npy(out[0]), "defect_start": npy(out[1]), "defect_end": npy(out[2])}
    if kind in ("conv1d_msc", "msc_legacy", "hybrid", "complex"):
        return {"defect_prob": npy(out)}
    flat = {}
    for k, v in out.items():
        if v is None:
            continue
        if isinstance(v, (list, tuple)):
            flat[k] = np.stack([npy(t) for t in v], axis=0)
        else:
            flat[k] = npy(v)
    return flat
