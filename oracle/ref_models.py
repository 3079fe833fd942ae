"""The reference's own classes from oracle/_ref (see build_ref.py), for the CPU baseline arm of bench.py.
Returns None when the build product is absent: the caller then times the oracle port and says kind = "port"."""
import importlib.util
import os

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


def _load(name):
    path = os.path.join(REF_DIR, name + ".py")
    if not os.path.exists(path):
        return None
    spec = importlib.util.spec_from_file_location("paut_ref_" + name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def reference_model(kind, signal_length=320):
    """Reference nn.Module of `kind` (random init; load a state_dict afterwards), or None."""
    if kind in ("msc", "msc_n"):
        m = _load("NN_models")
        if m is None:
            return None
        cls = m.MultiSignalClassifier if kind == "msc" else m.MultiSignalClassifier_N
        return cls(signal_length, [128, 64, 32], 4)
    if kind == "conv1d_msc":
        m = _load("msc_conv1d_model")
        return None if m is None else m.DefectDetectionModel(signal_length, 300)
    table = {"ssd": ("model", "SignalSequenceDetector"), "enhanced": ("enhanced_model", "EnhancedSignalSequenceDetector"),
             "two_stage": ("two_stage_model", "TwoStageDefectDetector")}
    if kind in table:
        m = _load(table[kind][0])
        if m is None:
            return None
        cls = getattr(m, table[kind][1])
        return cls(signal_length) if kind == "two_stage" else cls(signal_length=signal_length)
    return None
