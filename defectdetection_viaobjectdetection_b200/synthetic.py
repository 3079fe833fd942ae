"""Deterministic synthetic weights and inputs (bench.py, tests and golden fixtures; no reference arithmetic here).

The reference ships no checkpoint for any hot-path model, so parity is pinned
on synthetic weights.  They are generated per state_dict key from a
numpy PCG64 stream seeded by sha256(seed:key): independent of torch's RNG, of
key order and of the machine, so the build container (where the reference can
be imported) and the GPU box regenerate bit-identical tensors.

``state_spec(kind, **cfg)`` adds the fill rule to the state_dict contract (contract.state_spec, the single source) of
each reference class; ``tests/golden/make_golden.py`` verifies it against the
real classes (strict ``load_state_dict``) and stores the manifest.

Reference classes (paths relative to the reference root):
  msc        signals/multisignalNN/NN_models.py:45-128   MultiSignalClassifier
  msc_n      signals/multisignalNN/NN_models.py:198-246  MultiSignalClassifier_N
  conv1d_msc signals/MSC_Conv1D_training.py:50-89        DefectDetectionModel
  ssd        SignalSequenceDetection/model.py:230-285    SignalSequenceDetector
  enhanced   SignalSequenceDetection/enhanced_model.py:449-500
  two_stage  SignalSequenceDetection/two_stage_model.py:254-271
SURVEY section 8 "next" rows:
  msc_legacy signals/resaveModelOnnx.py:7-22 (the class the shipped MultiSignalClassifier_model*.pth belong to)
  improved   signals/improved_multisignal/improved_model.py:69-121   ImprovedMultiSignalClassifier
  hybrid     signals/improved_multisignal/detection_models/hybrid_binary.py:83-134   HybridBinaryModel
  complex    signals/improved_multisignal/detection_models/complex_detection_model.py:6-61
"""
from __future__ import annotations

import hashlib
import math
from collections import OrderedDict

import numpy as np

KINDS = ("msc", "msc_n", "conv1d_msc", "ssd", "enhanced", "two_stage",
         "msc_legacy", "improved", "hybrid", "complex")

# role -> how the tensor is filled
#   w   : U(-1/sqrt(fan_in), 1/sqrt(fan_in))   (fan_in = prod(shape[1:]))
#   b   : U(-1/sqrt(fan), 1/sqrt(fan)) with fan given explicitly
#   bnw / lnw : U(0.5, 1.5)     bnb / lnb : N(0, 0.1^2)
#   bnm : N(0, 0.1^2)           bnv : U(0.5, 1.5)        bnc : int64 scalar
#   randn : N(0, 1)             pe : sinusoidal table [1, max_len, d]


def state_spec(kind, **cfg):
    """Ordered {key: (shape, fill role, fan)} in the reference's registration order.  DERIVED from the one
    state_dict contract of the package (contract.state_spec: key, shape, parameter role); this function only adds how
    each tensor is filled: weights / biases U(+-1/sqrt(fan)) with the fan of the layer the tensor belongs to (a bias
    shares its weight's fan; recurrent tensors use the hidden size, as torch does), norms and BatchNorm statistics as
    listed above."""
    from . import contract
    entries = contract.state_spec(kind, **cfg)
    shapes = {k: tuple(sh) for k, sh, _ in entries}
    fill = {"nw": "bnw", "nb": "bnb", "bn_mean": "bnm", "bn_var": "bnv", "bn_count": "bnc", "pe": "pe", "randn": "randn"}

    def weight_fan(key):
        leaf = key.rsplit(".", 1)[-1]
        if leaf.startswith(("weight_ih_l", "weight_hh_l", "bias_ih_l", "bias_hh_l")):   # recurrent layer: hidden size
            hh = key.rsplit(".", 1)[0] + ".weight_hh_l" + leaf.split("_l", 1)[1]
            return shapes[hh][1]
        if leaf == "in_proj_bias":
            key = key[:-len("in_proj_bias")] + "in_proj_weight"
        elif leaf == "bias":
            key = key[:-len("bias")] + "weight"
        return int(np.prod(shapes[key][1:]))

    s = OrderedDict()
    for key, shape, role in entries:
        if role in ("w", "b"):
            s[key] = (tuple(shape), role, weight_fan(key))
        else:
            s[key] = (tuple(shape), fill[role], 0)
    return s


def _rng(seed, key):
    h = hashlib.sha256(f"{seed}:{key}".encode()).digest()
    return np.random.Generator(np.random.PCG64(int.from_bytes(h[:8], "little")))


def sinusoidal_pe(max_len, d):
    """The table of PositionalEncoding (model.py:14-21); it is a state_dict buffer."""
    pe = np.zeros((max_len, d), dtype=np.float32)
    pos = np.arange(max_len, dtype=np.float32)[:, None]
    div = np.exp(np.arange(0, d, 2, dtype=np.float32) * np.float32(-math.log(10000.0) / d)).astype(np.float32)
    pe[:, 0::2] = np.sin(pos * div)
    pe[:, 1::2] = np.cos(pos * div)
    return pe[None]


def synth_tensor(seed, key, shape, role, fan):
    g = _rng(seed, key)
    if role in ("w", "b"):
        bound = 1.0 / math.sqrt(max(fan, 1))
        return g.uniform(-bound, bound, size=shape).astype(np.float32)
    if role in ("bnw", "lnw", "bnv"):
        return g.uniform(0.5, 1.5, size=shape).astype(np.float32)
    if role in ("bnb", "lnb", "bnm"):
        return (0.1 * g.standard_normal(size=shape)).astype(np.float32)
    if role == "bnc":
        return np.array(7, dtype=np.int64)
    if role == "randn":
        return g.standard_normal(size=shape).astype(np.float32)
    if role == "pe":
        return sinusoidal_pe(shape[1], shape[2])
    raise ValueError(role)


def synth_state_dict(kind, seed=0, as_torch=True, **cfg):
    """Synthetic state_dict for a reference class (numpy, or torch tensors by default)."""
    out = OrderedDict()
    for key, (shape, role, fan) in state_spec(kind, **cfg).items():
        out[key] = synth_tensor(seed, key, shape, role, fan)
    if as_torch:
        import torch
        return OrderedDict((k, torch.from_numpy(v if v.ndim == 0 else np.ascontiguousarray(v))) for k, v in out.items())
    return out


def synth_paut_sets(n_sets, n_per_set, signal_length=320, seed=42, defect_frac=0.01, dtype=np.float32):
    """Synthetic A-scan sets in [0,1] after visualization/paut_data_generator.py:34-88:
    Gaussian back-wall echo (amp 0.6, sigma 10 samples) + N(0,0.05^2) noise per A-scan;
    ~defect_frac of the A-scans carry an extra echo (amp 0.56-1.08) at a random depth; clip to [0,1]."""
    g = np.random.Generator(np.random.PCG64(seed))
    t = np.arange(signal_length, dtype=np.float32)
    n = n_sets * n_per_set
    wall = 0.78 * signal_length + g.normal(0, 2.0, size=(n, 1)).astype(np.float32)
    x = 0.6 * np.exp(-0.5 * ((t[None] - wall) / 10.0) ** 2)
    x += g.normal(0, 0.05, size=(n, signal_length)).astype(np.float32)
    has = g.random(n) < defect_frac
    depth = g.uniform(0.15, 0.65, size=(n, 1)).astype(np.float32) * signal_length
    amp = g.uniform(0.56, 1.08, size=(n, 1)).astype(np.float32)
    echo = amp * np.exp(-0.5 * ((t[None] - depth) / 6.0) ** 2)
    x += echo * has[:, None]
    return np.clip(x, 0.0, 1.0).astype(dtype).reshape(n_sets, n_per_set, signal_length)


def synth_uniform_sets(n_sets, n_per_set, signal_length=320, seed=1234):
    g = np.random.Generator(np.random.PCG64(seed))
    return g.random((n_sets, n_per_set, signal_length), dtype=np.float32)
