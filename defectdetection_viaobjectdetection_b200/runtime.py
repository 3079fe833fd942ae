"""Thin host layer over the C ABI: one library context per (GPU, CUDA stream), native model handles.

PyTorch is used for device memory and streams only; every computation is a libpaut kernel.
"""
from __future__ import annotations

import ctypes as C
import threading

import numpy as np
import torch

from . import _lib
from ._lib import DETECTION, KINDS, PRECISION, Metrics, ModelCfg, Outputs, check

_contexts = {}
_contexts_lock = threading.Lock()

# (name, trailing shape as a function of (N, C, L)) per output slot -- include/paut.h "Output slots"
OUTPUT_SLOTS = {
    "msc": [("defect_prob", "N"), ("defect_start", "N"), ("defect_end", "N")],
    "msc_n": [("defect_prob", "N"), ("defect_start", "N"), ("defect_end", "N")],
    "conv1d_msc": [("defect_prob", "N")],
    "ssd": [("class_preds", "NC"), ("position_preds", "N2"), ("anomaly_scores", "N1"), ("attention_weights", "N1")],
    "enhanced": [("class_preds", "NC"), ("class_uncertainty", "NC"), ("position_preds", "N2"),
                 ("position_uncertainty", "N2"), ("anomaly_scores", "N1"), ("anomaly_uncertainty", "N1"),
                 ("attention_weights", "LNN"), ("context_attention", "N"), ("cross_attention", "NN")],
    "two_stage": [("defect_logits", "N2"), ("defect_probs", "N2"), ("defect_uncertainty", "N2"),
                  ("position_preds", "N2"), ("position_uncertainty", "N2")],
    # SURVEY section 8 "next" rows f2 / f3
    "msc_legacy": [("defect_prob", "N")],
    "improved": [("defect_prob", "N"), ("defect_start", "N"), ("defect_end", "N")],
    "hybrid": [("defect_prob", "N")],
    "complex": [("defect_prob", "N")],
}


class Context:
    """paut_ctx bound to a device and a CUDA stream."""

    def __init__(self, device_index, stream_handle):
        self.lib = _lib.load()
        self.device_index = device_index
        self.stream_handle = stream_handle
        h = C.c_void_p()
        check(self.lib.paut_ctx_create(device_index, C.c_void_p(stream_handle), C.byref(h)))
        self.handle = h

    def set_workspace_limit(self, nbytes):
        check(self.lib.paut_ctx_set_workspace_limit(self.handle, int(nbytes)), self.handle)

    @property
    def launch_count(self):
        return int(self.lib.paut_ctx_launch_count(self.handle))

    def profile_begin(self):
        check(self.lib.paut_ctx_profile_begin(self.handle), self.handle)

    def profile_end(self):
        """{kernel name: (launches, total_ms)} measured with CUDA events on the ctx stream."""
        buf = C.create_string_buffer(1 << 16)
        check(self.lib.paut_ctx_profile_end(self.handle, buf, len(buf)), self.handle)
        out = {}
        for line in buf.value.decode().splitlines():
            name, n, ms = line.split()
            out[name] = (int(n), float(ms))
        return out

    def close(self):
        if self.handle:
            self.lib.paut_ctx_destroy(self.handle)
            self.handle = None


def get_context(device=None):
    """Context for `device` and torch's current stream on it (created on first use)."""
    if not torch.cuda.is_available():
        raise RuntimeError("libpaut needs a CUDA device (sm_100a); there is no CPU fallback")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError(f"libpaut runs on CUDA devices only, got {dev}")
    index = dev.index if dev.index is not None else torch.cuda.current_device()
    stream = torch.cuda.current_stream(index).cuda_stream
    key = (index, stream)
    with _contexts_lock:
        ctx = _contexts.get(key)
        if ctx is None:
            ctx = Context(index, stream)
            _contexts[key] = ctx
    return ctx


def release_context(device_index, stream_handle):
    """Destroy the context cached for (device, raw stream handle) and give its workspace back to the driver.  Call it
    when the stream goes away (``VolumeScanner.close`` does for its lanes): contexts are keyed by the raw handle, and a
    handle that CUDA reuses for a new stream would otherwise pick up the stale context and its grown workspace.
    The models created on the context must be closed first (``PautModule.release``)."""
    with _contexts_lock:
        ctx = _contexts.pop((int(device_index), int(stream_handle)), None)
    if ctx is not None:
        ctx.close()
    return ctx is not None


def total_launches():
    return sum(c.launch_count for c in _contexts.values())


class NativeModel:
    """paut_model handle: weights packed once, forward / postprocess on device buffers."""

    def __init__(self, ctx, kind, cfg, precision="fp32"):
        self.ctx, self.kind, self.precision = ctx, kind, precision
        self.lib = ctx.lib
        c = ModelCfg()
        c.signal_length = int(cfg.get("signal_length", 0) or 0)
        hs = cfg.get("hidden_sizes") or (0, 0, 0)
        for i in range(3):
            c.hidden_sizes[i] = int(hs[i])
        c.num_heads = int(cfg.get("num_heads", 0) or 0)
        c.d_model = int(cfg.get("d_model", 0) or 0)
        c.num_classes = int(cfg.get("num_classes", 0) or 0)
        c.num_layers = int(cfg.get("num_layers", 0) or 0)
        c.dim_feedforward = int(cfg.get("dim_feedforward", 0) or 0)
        c.precision = PRECISION[precision]
        self.num_classes = c.num_classes or 2
        self.num_layers = c.num_layers
        h = C.c_void_p()
        check(self.lib.paut_model_create(ctx.handle, KINDS[kind], C.byref(c), C.byref(h)), ctx.handle)
        self.handle = h

    def keys(self):
        out = []
        key, shape, ndim = C.c_char_p(), (C.c_int64 * 4)(), C.c_int()
        for i in range(self.lib.paut_model_num_keys(self.handle)):
            check(self.lib.paut_model_key(self.handle, i, C.byref(key), shape, C.byref(ndim)), self.ctx.handle)
            out.append((key.value.decode(), tuple(shape[j] for j in range(ndim.value))))
        return out

    def load_state_dict(self, state_dict):
        for key, _ in self.keys():
            if key not in state_dict:
                raise KeyError(f"state_dict is missing {key!r}")
            t = state_dict[key].detach()
            if t.dtype == torch.int64:
                t, dt = t.contiguous(), _lib.I64
            else:
                t, dt = t.to(torch.float32).contiguous(), _lib.F32
            if t.is_cuda:
                # set_tensor copies with a blocking cudaMemcpy on the legacy stream, which is NOT ordered behind a
                # conversion / contiguous kernel enqueued on a non-blocking torch stream (scanner lanes)
                torch.cuda.current_stream(t.device).synchronize()
            shape = (C.c_int64 * max(t.dim(), 1))(*t.shape)
            check(self.lib.paut_model_set_tensor(self.handle, key.encode(), C.c_void_p(t.data_ptr()), dt, shape,
                                                 t.dim()), self.ctx.handle)
        check(self.lib.paut_model_finalize(self.handle), self.ctx.handle)

    def _alloc_outputs(self, B, N, device, wanted=None, num_layers=1):
        outs, struct = {}, Outputs()
        shapes = {"N": (B, N), "NC": (B, N, self.num_classes), "N2": (B, N, 2), "N1": (B, N, 1),
                  "NN": (B, N, N), "LNN": (num_layers, B, N, N)}
        for i, (name, code) in enumerate(OUTPUT_SLOTS[self.kind]):
            if wanted is not None and name not in wanted:
                continue
            t = torch.empty(shapes[code], dtype=torch.float32, device=device)
            outs[name] = t
            struct.slot[i] = t.data_ptr()
        return outs, struct

    def forward(self, x, wanted=None, num_layers=1):
        """x: CUDA tensor [B,N,S] (conv1d_msc: [B,S,N]), fp32 or bf16, contiguous."""
        if not x.is_cuda:
            raise RuntimeError("libpaut forward needs a CUDA tensor (no CPU fallback)")
        if x.dim() != 3:
            raise ValueError("expected a 3-D input [batch, signals, samples]")
        if not x.is_contiguous():
            # the reference's x.view(...) raises on non-contiguous input (NN_models.py:111)
            raise RuntimeError("input must be contiguous")
        if x.dtype == torch.float32:
            dt = _lib.F32
        elif x.dtype == torch.bfloat16:
            dt = _lib.BF16
        else:
            raise TypeError(f"unsupported input dtype {x.dtype}")
        if self.kind == "conv1d_msc":
            B, S, N = x.shape
        else:
            B, N, S = x.shape
        outs, struct = self._alloc_outputs(B, N, x.device, wanted, num_layers)
        check(self.lib.paut_forward(self.handle, C.c_void_p(x.data_ptr()), dt, B, N, S, C.byref(struct)),
              self.ctx.handle)
        return outs, struct, (B, N, S)

    def debug_stage(self, stage, x, width):
        """Intermediate tensor of a fused kernel (paut_debug_stage) for kernel-level parity tests: fp32 [B*N, width]."""
        if not (x.is_cuda and x.is_contiguous() and x.dim() == 3):
            raise RuntimeError("x must be a contiguous CUDA tensor [B,N,S]")
        B, N, S = x.shape
        dt = {torch.float32: _lib.F32, torch.bfloat16: _lib.BF16}[x.dtype]
        out = torch.full((B * N, width), float("nan"), dtype=torch.float32, device=x.device)
        check(self.lib.paut_debug_stage(self.handle, stage, C.c_void_p(x.data_ptr()), dt, B, N, S,
                                        C.c_void_p(out.data_ptr())), self.ctx.handle)
        return out

    def postprocess(self, struct, B, N, S, threshold, device):
        """Device post-processing; returns (records tensor [B*N*48] uint8 on device, count tensor int32)."""
        det = torch.empty(B * N * DETECTION.itemsize, dtype=torch.uint8, device=device)
        count = torch.zeros(1, dtype=torch.int32, device=device)
        check(self.lib.paut_postprocess(self.handle, C.byref(struct), B, N, S, float(threshold),
                                        C.c_void_p(det.data_ptr()), C.c_void_p(count.data_ptr())), self.ctx.handle)
        return det, count

    def close(self):
        if self.handle:
            self.lib.paut_model_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def records_to_numpy(det, count):
    """D2H of the kept records only -> numpy structured array (dtype DETECTION)."""
    n = int(count.item())
    if n == 0:
        return np.zeros(0, dtype=DETECTION)
    raw = det[: n * DETECTION.itemsize].cpu().numpy()          # fresh host array: view it, no second copy
    return raw.view(DETECTION)


def window_table(rule, n, seq_length=50):
    """(start, valid_len) pairs of the reference windowing rules; rule 'msc' or 'ssd' (host arithmetic)."""
    lib = _lib.load()
    r = {"msc": 0, "ssd": 1}[rule]
    cnt = lib.paut_window_table_host(r, n, seq_length, None, 0)
    if cnt < 0:
        raise ValueError("bad window arguments")
    buf = (C.c_int32 * (2 * max(cnt, 1)))()
    lib.paut_window_table_host(r, n, seq_length, buf, cnt)
    return [(buf[2 * i], buf[2 * i + 1]) for i in range(cnt)]


def group_nonzero(volume):
    """dataset_preparation.py:205 on device: bool numpy mask [G], False where volume[g] is all zero."""
    if not volume.is_cuda or not volume.is_contiguous() or volume.dim() != 3:
        raise RuntimeError("volume must be a contiguous CUDA tensor [G, n, S]")
    G, n, S = volume.shape
    ctx = get_context(volume.device)
    dt = {torch.float32: _lib.F32, torch.bfloat16: _lib.BF16}[volume.dtype]
    flags = torch.empty(G, dtype=torch.int32, device=volume.device)
    check(ctx.lib.paut_group_nonzero(ctx.handle, C.c_void_p(volume.data_ptr()), dt, G, n, S,
                                     C.c_void_p(flags.data_ptr())), ctx.handle)
    return flags.cpu().numpy().astype(bool)


def gather_windows(volume, rule, seq_length=50, out_dtype=None, keep_groups=None, drop_all_zero=False):
    """Device windowing: volume [G, n, S] (CUDA, fp32/bf16) -> (sets [W, L, S], table int32 [W,3] on host).
    keep_groups: optional bool mask [G]; drop_all_zero: compute it on the device with the all-zero-run rule of
    dataset_preparation.py:205."""
    if not volume.is_cuda or not volume.is_contiguous():
        raise RuntimeError("volume must be a contiguous CUDA tensor")
    G, n, S = volume.shape
    if drop_all_zero and keep_groups is None:
        keep_groups = group_nonzero(volume)
    wins = window_table(rule, n, seq_length)
    groups = range(G) if keep_groups is None else [g for g in range(G) if bool(keep_groups[g])]
    table = np.array([(g, s, v) for g in groups for (s, v) in wins], dtype=np.int32).reshape(-1, 3)
    out_dtype = out_dtype or volume.dtype
    sets = torch.empty((len(table), seq_length, S), dtype=out_dtype, device=volume.device)
    if len(table) == 0:
        return sets, table
    ctx = get_context(volume.device)
    dt = {torch.float32: _lib.F32, torch.bfloat16: _lib.BF16}
    table_dev = torch.from_numpy(table).to(volume.device)
    check(ctx.lib.paut_window_gather(ctx.handle, C.c_void_p(volume.data_ptr()), dt[volume.dtype], G, n, S,
                                     C.c_void_p(table_dev.data_ptr()), len(table), seq_length,
                                     C.c_void_p(sets.data_ptr()), dt[out_dtype]), ctx.handle)
    return sets, table


# ------------------------------------------------------------------------------------------------ section 8 rows f3 / f4
def _f32c(t, name):
    if not (t.is_cuda and t.is_contiguous() and t.dtype == torch.float32):
        raise RuntimeError(f"{name} must be a contiguous fp32 CUDA tensor")
    return t


def difference_matrix(x, prob, threshold=0.5, want_diff=True):
    """signals/teststtt.py:54-69 on device for a batch of sets: x [B,N,S] (fp32/bf16), prob [B,N] fp32 ->
    (reference [B,S], diff [B,N,S] or None, healthy_count int32 [B]).  Sets without a healthy A-scan
    (the reference returns None) have healthy_count 0 and zero rows."""
    if not (x.is_cuda and x.is_contiguous() and x.dim() == 3):
        raise RuntimeError("x must be a contiguous CUDA tensor [B,N,S]")
    B, N, S = x.shape
    _f32c(prob, "prob")
    dt = {torch.float32: _lib.F32, torch.bfloat16: _lib.BF16}[x.dtype]
    ctx = get_context(x.device)
    ref = torch.empty((B, S), dtype=torch.float32, device=x.device)
    diff = torch.empty((B, N, S), dtype=torch.float32, device=x.device) if want_diff else None
    healthy = torch.empty(B, dtype=torch.int32, device=x.device)
    check(ctx.lib.paut_difference_matrix(ctx.handle, C.c_void_p(x.data_ptr()), dt, C.c_void_p(prob.data_ptr()), B, N, S,
                                         float(threshold), C.c_void_p(ref.data_ptr()),
                                         C.c_void_p(diff.data_ptr()) if want_diff else None,
                                         C.c_void_p(healthy.data_ptr())), ctx.handle)
    return ref, diff, healthy


def _metrics_out(device):
    return torch.zeros(C.sizeof(Metrics), dtype=torch.uint8, device=device)


def _metrics_read(buf):
    m = Metrics.from_buffer_copy(buf.cpu().numpy().tobytes())
    return dict(tp=m.tp, fp=m.fp, fn=m.fn, tn=m.tn, sum_iou=m.sum_iou, sum_position_error=m.sum_position_error)


def metrics_match(rule, det, count, B, N, target_label, target_pos, iou_threshold=0.5):
    """Detection-level matching on device.  rule 'position' = two_stage_train.py:284-375, 'class' =
    train.py:279-361.  det/count: what NativeModel.postprocess returned; target_label int32 [B,N],
    target_pos fp32 [B,N,2].  Returns the paut_metrics fields as a dict."""
    r = {"position": 0, "class": 1}[rule]
    if not (target_label.is_cuda and target_label.dtype == torch.int32 and target_label.is_contiguous()):
        raise RuntimeError("target_label must be a contiguous int32 CUDA tensor")
    _f32c(target_pos, "target_pos")
    ctx = get_context(det.device)
    out = _metrics_out(det.device)
    check(ctx.lib.paut_metrics_match(ctx.handle, r, C.c_void_p(det.data_ptr()), C.c_void_p(count.data_ptr()), B, N,
                                     C.c_void_p(target_label.data_ptr()), C.c_void_p(target_pos.data_ptr()),
                                     float(iou_threshold), C.c_void_p(out.data_ptr())), ctx.handle)
    return _metrics_read(out)


def metrics_confusion(prob, label, threshold=0.5, ge=True):
    """acc_metrics_hybrid_binary_dynamic_.py:73-94 on device: TP / FP / FN / TN of (prob >= thr) vs (label > 0.5)."""
    _f32c(prob, "prob")
    _f32c(label, "label")
    ctx = get_context(prob.device)
    out = _metrics_out(prob.device)
    check(ctx.lib.paut_metrics_confusion(ctx.handle, C.c_void_p(prob.data_ptr()), C.c_void_p(label.data_ptr()),
                                         prob.numel(), float(threshold), 1 if ge else 0, C.c_void_p(out.data_ptr())),
          ctx.handle)
    return _metrics_read(out)


def detection_metrics(rule, m):
    """paut_metrics counts -> the dict the reference's calculate_metrics returns (same keys, same formulas).
    rule 'position': two_stage_train.py:361-375; rule 'class': train.py:343-361."""
    tp, fp, fn = m["tp"], m["fp"], m["fn"]
    if rule == "position":
        precision = tp / max(tp + fp, 1)
        recall = tp / max(tp + fn, 1)
        return {"precision": precision, "recall": recall,
                "f1_score": 2 * precision * recall / max(precision + recall, 1e-8),
                "mean_position_error": m["sum_position_error"] / tp if tp else 0,
                "true_positives": tp, "false_positives": fp, "false_negatives": fn}
    if rule == "class":
        precision = tp / (tp + fp) if tp + fp > 0 else 0
        recall = tp / (tp + fn) if tp + fn > 0 else 0
        return {"precision": precision, "recall": recall,
                "f1": 2 * precision * recall / (precision + recall) if precision + recall > 0 else 0,
                "mean_iou": m["sum_iou"] / tp if tp else 0,
                "true_positives": tp, "false_positives": fp, "false_negatives": fn}
    raise ValueError(rule)
