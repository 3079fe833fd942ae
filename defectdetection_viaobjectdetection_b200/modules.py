"""Drop-in ``nn.Module`` replacements for the reference's PAUT signal models.

Same constructor arguments, ``forward`` / ``predict`` signatures, return structures and
``state_dict()`` keys as the reference classes; the computation is libpaut.so (hand-written
sm_100a kernels) -- there is no PyTorch or CPU fallback, and training (``targets``) is out of scope.

    from defectdetection_viaobjectdetection_b200 import TwoStageDefectDetector
    m = TwoStageDefectDetector(signal_length=320).cuda().eval()
    m.load_state_dict(torch.load("two_stage.pth")["model_state_dict"])
    preds = m.predict(x_cuda, threshold=0.5)          # list[list[dict]] like the reference

``model.precision`` selects the arithmetic of the per-A-scan encoder: ``"fp32"`` (CUDA-core fp32,
logits within 1e-4 of the reference), ``"bf16"`` (tcgen05 tensor cores, bf16 operands, fp32
accumulate, within 1e-2) or ``"auto"`` (default: follows the dtype of ``x``).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn as nn

from . import contract
from .runtime import DETECTION, NativeModel, get_context, records_to_numpy


def _sinusoidal_pe(max_len, d):
    pe = torch.zeros(max_len, d)
    position = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
    div = torch.exp(torch.arange(0, d, 2).float() * (-math.log(10000.0) / d))
    pe[:, 0::2] = torch.sin(position * div)
    pe[:, 1::2] = torch.cos(position * div)
    return pe.unsqueeze(0)


def _init_tensor(shape, role):
    if role in ("w", "b"):
        fan_in = int(np.prod(shape[1:])) if len(shape) > 1 else max(int(shape[0]), 1)
        bound = 1.0 / math.sqrt(max(fan_in, 1))
        return torch.empty(shape).uniform_(-bound, bound)
    if role in ("nw", "bn_var"):
        return torch.ones(shape)
    if role in ("nb", "bn_mean"):
        return torch.zeros(shape)
    if role == "bn_count":
        return torch.tensor(0, dtype=torch.long)
    if role == "randn":
        return torch.randn(shape)
    if role == "pe":
        return _sinusoidal_pe(shape[1], shape[2])
    raise ValueError(role)


class _Node(nn.Module):
    """Anonymous container: only there so parameters get the reference's dotted names."""


class PautModule(nn.Module):
    """Base: builds the parameter tree from the contract and drives the native model."""

    _kind = None

    def __init__(self, cfg):
        super().__init__()
        self._cfg = dict(cfg)
        self.precision = "auto"
        self._native = {}
        self._fingerprint = {}
        spec_cfg = {k: v for k, v in cfg.items() if k in ("signal_length", "hidden_sizes", "d_model", "num_classes",
                                                         "num_layers", "dim_feedforward") and v}
        for key, shape, role in contract.state_spec(self._kind, **spec_cfg):
            *path, leaf = key.split(".")
            node = self
            for name in path:
                if name not in node._modules:
                    node.add_module(name, _Node())
                node = node._modules[name]
            t = _init_tensor(shape, role)
            if role in contract.BUFFER_ROLES:
                node.register_buffer(leaf, t)
            else:
                node.register_parameter(leaf, nn.Parameter(t))

    # ---- native model management -------------------------------------------------------------
    def _state_fingerprint(self):
        return tuple((t.data_ptr(), t._version) for t in self.state_dict(keep_vars=True).values())

    def _native_for(self, x):
        if self.training:
            raise RuntimeError("libpaut modules are inference-only: call .eval() (training is out of scope)")
        precision = self.precision
        if precision == "auto":
            precision = "bf16" if x.dtype == torch.bfloat16 else "fp32"
        if precision not in ("fp32", "bf16"):
            raise ValueError("precision must be 'fp32', 'bf16' or 'auto'")
        ctx = get_context(x.device)
        key = (ctx.device_index, ctx.stream_handle, precision)
        fp = self._state_fingerprint()
        native = self._native.get(key)
        if native is None or self._fingerprint.get(key) != fp:
            sd = self.state_dict()
            first = next(iter(sd.values()))
            if first.device != x.device:
                raise RuntimeError(f"module is on {first.device} but the input is on {x.device}")
            if native is None:
                native = NativeModel(ctx, self._kind, self._cfg, precision)
                self._native[key] = native
            native.load_state_dict(sd)
            self._fingerprint[key] = fp
        return native

    def release(self, device_index=None, stream_handle=None):
        """Close the packed native copies of this module made for one (device, stream) -- or all of them -- so that their
        context can be destroyed (``runtime.release_context``); the next call on that stream packs the weights again."""
        for key in list(self._native):
            if (device_index is None or key[0] == device_index) and (stream_handle is None or key[1] == stream_handle):
                self._native.pop(key).close()
                self._fingerprint.pop(key, None)

    def _run(self, x, wanted=None):
        native = self._native_for(x)
        nl = int(self._cfg.get("num_layers") or 6)
        return native, native.forward(x, wanted, nl)

    def _check_targets(self, targets):
        if targets is not None:
            raise NotImplementedError(
                "the training loss (forward with targets) is out of scope of the B200 inference path")

    # ---- post-processing ----------------------------------------------------------------------
    @torch.no_grad()
    def predict_records(self, x, threshold=0.5):
        """Device post-processing: numpy structured array of kept detections (dtype ``DETECTION``) in
        the reference's (set, position) order, with integer sample indices."""
        native, (outs, struct, (B, N, S)) = self._run(x)
        det, count = native.postprocess(struct, B, N, S, threshold, x.device)
        return records_to_numpy(det, count)


# ------------------------------------------------------------------------------------------------ signals/
class MultiSignalClassifier(PautModule):
    """signals/multisignalNN/NN_models.py:45-128.  forward(x[B,N,S]) -> (defect_prob, defect_start, defect_end)."""

    _kind = "msc"

    def __init__(self, signal_length, hidden_sizes, num_heads=4):
        hidden_sizes = list(hidden_sizes)
        if hidden_sizes[1] % num_heads != 0:
            raise AssertionError("embed_dim must be divisible by num_heads")
        super().__init__(dict(signal_length=signal_length, hidden_sizes=tuple(hidden_sizes[:3]), num_heads=num_heads))

    @torch.no_grad()
    def forward(self, x):
        _, (o, _, _) = self._run(x)
        return o["defect_prob"], o["defect_start"], o["defect_end"]


class MultiSignalClassifier_N(MultiSignalClassifier):
    """signals/multisignalNN/NN_models.py:198-246 (background extractor + local attention)."""

    _kind = "msc_n"


class DefectDetectionModel(PautModule):
    """signals/MSC_Conv1D_training.py:50-89 ("MSC Conv1D").  forward(x[B,S,N]) -> [B,N]."""

    _kind = "conv1d_msc"

    def __init__(self, signal_length, num_signals_per_set):
        super().__init__(dict(signal_length=signal_length))
        self.num_signals_per_set = num_signals_per_set

    @torch.no_grad()
    def forward(self, x):
        _, (o, _, _) = self._run(x)
        return o["defect_prob"]


# ------------------------------------------------------------------------------------------------ SignalSequenceDetection/
def _group_by_set(rec, B):
    if len(rec) == 0:
        return [[] for _ in range(B)]
    bounds = np.searchsorted(rec["set_index"], np.arange(B + 1))
    return [rec[bounds[b]:bounds[b + 1]] for b in range(B)]


class SignalSequenceDetector(PautModule):
    """SignalSequenceDetection/model.py:230-475."""

    _kind = "ssd"

    def __init__(self, signal_length=100, d_model=128, num_classes=2, nhead=8, num_layers=4, dim_feedforward=512,
                 dropout=0.1):
        if d_model % nhead != 0:
            raise AssertionError("embed_dim must be divisible by num_heads")
        super().__init__(dict(signal_length=signal_length, d_model=d_model, num_classes=num_classes, num_heads=nhead,
                              num_layers=num_layers, dim_feedforward=dim_feedforward))

    @torch.no_grad()
    def forward(self, x, targets=None):
        self._check_targets(targets)
        _, (o, _, _) = self._run(x)
        return {k: o[k] for k in ("class_preds", "position_preds", "anomaly_scores", "attention_weights")}

    def predict(self, x, threshold=0.5):
        """model.py:424-475 -- list (per set) of dicts for the kept A-scans."""
        B = x.shape[0]
        rec = self.predict_records(x, threshold)
        return [[{"position": int(r["position"]), "class": int(r["cls"]), "class_score": float(r["score"]),
                  "defect_position": np.array([r["start"], r["end"]], dtype=np.float32),
                  "anomaly_score": float(r["anomaly"])} for r in rows] for rows in _group_by_set(rec, B)]


class EnhancedSignalSequenceDetector(PautModule):
    """SignalSequenceDetection/enhanced_model.py:449-807."""

    _kind = "enhanced"

    def __init__(self, signal_length=100, d_model=256, num_classes=2, nhead=8, num_layers=6, dim_feedforward=1024,
                 dropout=0.1):
        if d_model % nhead != 0:
            raise AssertionError("embed_dim must be divisible by num_heads")
        super().__init__(dict(signal_length=signal_length, d_model=d_model, num_classes=num_classes, num_heads=nhead,
                              num_layers=num_layers, dim_feedforward=dim_feedforward))

    @torch.no_grad()
    def forward(self, x, targets=None):
        self._check_targets(targets)
        _, (o, _, _) = self._run(x)
        out = {k: o[k] for k in ("class_preds", "class_uncertainty", "position_preds", "position_uncertainty",
                                 "anomaly_scores", "anomaly_uncertainty")}
        out["attention_weights"] = list(o["attention_weights"].unbind(0))
        out["context_attention"] = o["context_attention"]
        out["cross_attention"] = o["cross_attention"]
        return out

    @torch.no_grad()
    def predict(self, x, threshold=0.5):
        """enhanced_model.py:741-807."""
        native, (o, struct, (B, N, S)) = self._run(x)
        det, count = native.postprocess(struct, B, N, S, threshold, x.device)
        rec = records_to_numpy(det, count)
        pos_unc = o["position_uncertainty"].cpu().numpy()
        an_unc = o["anomaly_uncertainty"].cpu().numpy()
        out = []
        for b, rows in enumerate(_group_by_set(rec, B)):
            out.append([{"position": int(r["position"]), "class": int(r["cls"]), "class_score": float(r["score"]),
                         "class_uncertainty": float(r["uncertainty"]),
                         "defect_position": np.array([r["start"], r["end"]], dtype=np.float32),
                         "position_uncertainty": pos_unc[b, int(r["position"])],
                         "anomaly_score": float(r["anomaly"]),
                         "anomaly_uncertainty": float(an_unc[b, int(r["position"]), 0]),
                         "adjusted_confidence": float(r["confidence"])} for r in rows])
        return out


class TwoStageDefectDetector(PautModule):
    """SignalSequenceDetection/two_stage_model.py:254-501."""

    _kind = "two_stage"

    def __init__(self, signal_length, d_model=128, num_classes=2):
        super().__init__(dict(signal_length=signal_length, d_model=d_model, num_classes=2))

    @torch.no_grad()
    def forward(self, x, targets=None):
        self._check_targets(targets)
        _, (o, _, _) = self._run(x)
        out = {k: o[k] for k in ("defect_logits", "defect_probs", "defect_uncertainty", "position_preds",
                                 "position_uncertainty")}
        out["attention_weights"] = None                                   # two_stage_model.py:172
        return out

    @torch.no_grad()
    def predict(self, x, threshold=0.5):
        """two_stage_model.py:454-501."""
        native, (o, struct, (B, N, S)) = self._run(x)
        det, count = native.postprocess(struct, B, N, S, threshold, x.device)
        rec = records_to_numpy(det, count)
        pos_unc = o["position_uncertainty"].cpu().numpy()
        return [[{"position": int(r["position"]), "defect_prob": float(r["score"]),
                  "defect_uncertainty": float(r["uncertainty"]),
                  "defect_position": np.array([r["start"], r["end"]], dtype=np.float32),
                  "position_uncertainty": pos_unc[b, int(r["position"])],
                  "adjusted_confidence": float(r["confidence"])} for r in rows]
                for b, rows in enumerate(_group_by_set(rec, B))]


# ------------------------------------------------------------------------------------------------ SURVEY section 8 "next" rows
class MultiSignalClassifierLegacy(PautModule):
    """The no-conv ``MultiSignalClassifier`` of signals/resaveModelOnnx.py:7-33 (same class in
    GNN_testing_multi_v2_MAP.py:16-36, teststtt.py) -- the one the repository's MultiSignalClassifier_model*.pth
    checkpoints load into.  forward(x[B,N,S]) -> outputs[B,N]."""

    _kind = "msc_legacy"

    def __init__(self, signal_length, hidden_sizes):
        super().__init__(dict(signal_length=signal_length, hidden_sizes=tuple(list(hidden_sizes)[:3])))

    @torch.no_grad()
    def forward(self, x):
        _, (o, _, _) = self._run(x)
        return o["defect_prob"]

    @torch.no_grad()
    def prediction_map(self, volume, num_signals_per_set):
        """GNN_testing_multi_v2_MAP.py:38-67 (load_and_predict + generate_prediction_map): one row per A-scan-index
        folder, computed from the folder's first ``num_signals_per_set`` signals.  volume [folders, n, S] -> [folders,
        num_signals_per_set] probabilities (the reference plots ``map * 100``)."""
        return self.forward(volume[:, :num_signals_per_set].contiguous())

    @torch.no_grad()
    def difference_matrix(self, x, threshold=0.5):
        """teststtt.py:54-69 for every set of x: (outputs [B,N], reference [B,S], diff [B,N,S], healthy_count [B])."""
        from .runtime import difference_matrix
        prob = self.forward(x)
        ref, diff, healthy = difference_matrix(x, prob, threshold)
        return prob, ref, diff, healthy


class ImprovedMultiSignalClassifier(PautModule):
    """signals/improved_multisignal/improved_model.py:69-197."""

    _kind = "improved"

    def __init__(self, signal_length, hidden_sizes, num_heads=8, dropout=0.1, num_transformer_layers=4):
        hidden_sizes = list(hidden_sizes)
        if hidden_sizes[1] % num_heads != 0:
            raise AssertionError("embed_dim must be divisible by num_heads")
        super().__init__(dict(signal_length=signal_length, hidden_sizes=tuple(hidden_sizes[:3]), num_heads=num_heads,
                              num_layers=num_transformer_layers))

    @torch.no_grad()
    def forward(self, x):
        _, (o, _, _) = self._run(x)
        return o["defect_prob"], o["defect_start"], o["defect_end"]

    def predict(self, x, threshold=0.5):
        """improved_model.py:160-197: keep iff prob >= threshold."""
        B = x.shape[0]
        rec = self.predict_records(x, threshold)
        return [[{"position": int(r["position"]), "defect_prob": float(r["score"]),
                  "defect_position": [float(r["start"]), float(r["end"])]} for r in rows]
                for rows in _group_by_set(rec, B)]


class HybridBinaryModel(PautModule):
    """signals/improved_multisignal/detection_models/hybrid_binary.py:83-168.  forward(x[B,N,S]) -> defect_prob[B,N]."""

    _kind = "hybrid"

    def __init__(self, signal_length=320, hidden_sizes=(256, 128, 48), num_heads=8, dropout=0.15,
                 num_transformer_layers=4):
        hidden_sizes = list(hidden_sizes)
        if hidden_sizes[1] % num_heads != 0:
            raise AssertionError("embed_dim must be divisible by num_heads")
        super().__init__(dict(signal_length=signal_length, hidden_sizes=tuple(hidden_sizes[:3]), num_heads=num_heads,
                              num_layers=num_transformer_layers))

    @torch.no_grad()
    def forward(self, x):
        _, (o, _, _) = self._run(x)
        return o["defect_prob"]


class ComplexDetectionModel(PautModule):
    """signals/improved_multisignal/detection_models/complex_detection_model.py:6-96.  forward -> detection_prob[B,N]."""

    _kind = "complex"

    def __init__(self, signal_length=320, d_model=64, num_heads=8, num_layers=4, dropout=0.1):
        if d_model % num_heads != 0:
            raise AssertionError("embed_dim must be divisible by num_heads")
        super().__init__(dict(signal_length=signal_length, d_model=d_model, num_heads=num_heads, num_layers=num_layers))

    @torch.no_grad()
    def forward(self, x):
        _, (o, _, _) = self._run(x)
        return o["defect_prob"]


# kind -> constructor from a cfg dict (bench.py and the tests build every model through this table)
FACTORIES = {
    "msc": lambda cfg: MultiSignalClassifier(cfg.get("signal_length", 320), [128, 64, 32], 4),
    "msc_n": lambda cfg: MultiSignalClassifier_N(cfg.get("signal_length", 320), [128, 64, 32], 4),
    "conv1d_msc": lambda cfg: DefectDetectionModel(320, 300),
    "ssd": lambda cfg: SignalSequenceDetector(**cfg),
    "enhanced": lambda cfg: EnhancedSignalSequenceDetector(**cfg),
    "two_stage": lambda cfg: TwoStageDefectDetector(cfg.get("signal_length", 320)),
    "msc_legacy": lambda cfg: MultiSignalClassifierLegacy(cfg.get("signal_length", 320), [128, 64, 32]),
    "improved": lambda cfg: ImprovedMultiSignalClassifier(cfg.get("signal_length", 320), [128, 64, 32], 8),
    "hybrid": lambda cfg: HybridBinaryModel(**{k: v for k, v in cfg.items() if k in ("signal_length", "hidden_sizes")}),
    "complex": lambda cfg: ComplexDetectionModel(),
}


def sample_indices(defect_position, signal_length):
    """predict.py:111-113 / signal_visualizer.py:409-410: int(start * len(signal)) with the float32 product."""
    p = np.asarray(defect_position, dtype=np.float32) * np.float32(signal_length)
    return np.trunc(p).astype(np.int32)


def load_checkpoint_state(ckpt):
    """Accept both checkpoint formats of the reference (bare state_dict, or dict with 'model_state_dict')."""
    return ckpt.get("model_state_dict", ckpt) if isinstance(ckpt, dict) else ckpt


__all__ = ["MultiSignalClassifier", "MultiSignalClassifier_N", "DefectDetectionModel", "SignalSequenceDetector",
           "EnhancedSignalSequenceDetector", "TwoStageDefectDetector", "MultiSignalClassifierLegacy",
           "ImprovedMultiSignalClassifier", "HybridBinaryModel", "ComplexDetectionModel", "sample_indices",
           "load_checkpoint_state", "DETECTION"]
