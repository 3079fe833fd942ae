"""Deterministic synthetic weights and inputs (bench.py, tests and golden fixtures; no reference arithmetic here).

The reference ships no checkpoint for any hot-path model, so parity is pinned
on synthetic weights.  They are generated per state_dict key from a
numpy PCG64 stream seeded by sha256(seed:key): independent of torch's RNG, of
key order and of the machine, so the build container (where the reference can
be imported) and the GPU box regenerate bit-identical tensors.

``state_spec(kind, **cfg)`` is the state_dict contract (key -> shape, role) of
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


def _lin(spec, name, out_f, in_f):
    spec[name + ".weight"] = ((out_f, in_f), "w", in_f)
    spec[name + ".bias"] = ((out_f,), "b", in_f)


def _conv(spec, name, cout, cin_per_group, k):
    spec[name + ".weight"] = ((cout, cin_per_group, k), "w", cin_per_group * k)
    spec[name + ".bias"] = ((cout,), "b", cin_per_group * k)


def _bn(spec, name, c):
    spec[name + ".weight"] = ((c,), "bnw", 0)
    spec[name + ".bias"] = ((c,), "bnb", 0)
    spec[name + ".running_mean"] = ((c,), "bnm", 0)
    spec[name + ".running_var"] = ((c,), "bnv", 0)
    spec[name + ".num_batches_tracked"] = ((), "bnc", 0)


def _ln(spec, name, c):
    spec[name + ".weight"] = ((c,), "lnw", 0)
    spec[name + ".bias"] = ((c,), "lnb", 0)


def _mha(spec, name, d):
    spec[name + ".in_proj_weight"] = ((3 * d, d), "w", d)
    spec[name + ".in_proj_bias"] = ((3 * d,), "b", d)
    spec[name + ".out_proj.weight"] = ((d, d), "w", d)
    spec[name + ".out_proj.bias"] = ((d,), "b", d)


def _tel(spec, name, d, dff):
    """nn.TransformerEncoderLayer parameter order."""
    _mha(spec, name + ".self_attn", d)
    _lin(spec, name + ".linear1", dff, d)
    _lin(spec, name + ".linear2", d, dff)
    _ln(spec, name + ".norm1", d)
    _ln(spec, name + ".norm2", d)


def _rnn(spec, name, gates, in0, hidden, layers=2):
    for layer in range(layers):
        in_f = in0 if layer == 0 else 2 * hidden
        for suffix in ("", "_reverse"):
            spec[f"{name}.weight_ih_l{layer}{suffix}"] = ((gates * hidden, in_f), "w", hidden)
            spec[f"{name}.weight_hh_l{layer}{suffix}"] = ((gates * hidden, hidden), "w", hidden)
            spec[f"{name}.bias_ih_l{layer}{suffix}"] = ((gates * hidden,), "b", hidden)
            spec[f"{name}.bias_hh_l{layer}{suffix}"] = ((gates * hidden,), "b", hidden)


def state_spec(kind, signal_length=320, hidden_sizes=None, d_model=None,
               num_classes=2, num_layers=None, dim_feedforward=None, num_heads=None):
    """Ordered {key: (shape, role, fan)} following the reference's registration order."""
    s = OrderedDict()
    hidden_sizes = tuple(hidden_sizes or ((256, 128, 48) if kind == "hybrid" else (128, 64, 32)))
    if kind in ("msc", "msc_n"):
        h0, h1, h2 = hidden_sizes
        _conv(s, "conv1d.0", 8, 1, 3)
        _conv(s, "conv1d.2", 16, 8, 3)
        if kind == "msc_n":
            _conv(s, "background_extractor", 16, 1, 11)
        _lin(s, "shared_layer.0", h0, signal_length)
        _lin(s, "shared_layer.2", h1, h0)
        s["position_encoding.encoding"] = ((300, h1), "randn", 0)
        _mha(s, "transformer_encoder.self_attn", h1)
        if kind == "msc":
            _mha(s, "transformer_encoder.cross_attn", h1)
        else:
            _conv(s, "transformer_encoder.local_attn.local_conv", h1, 1, 5)
        _lin(s, "transformer_encoder.ffn.0", h2, h1)
        _lin(s, "transformer_encoder.ffn.2", h1, h2)
        for i in (1, 2, 3):
            _ln(s, f"transformer_encoder.norm{i}", h1)
        _lin(s, "classifier", 3, h1)
    elif kind == "conv1d_msc":
        _conv(s, "feature_extractor.0", 64, 1, 3)
        _conv(s, "feature_extractor.2", 128, 64, 3)
        _conv(s, "feature_extractor.4", 128, 128, 1)
        for i in range(4):
            _tel(s, f"transformer_encoder.layers.{i}", 128, 2048)
        _lin(s, "classifier.0", 64, 128)
        _lin(s, "classifier.2", 1, 64)
    elif kind == "ssd":
        d = d_model or 128
        nl = num_layers or 4
        dff = dim_feedforward or 512
        _conv(s, "signal_encoder.conv1", 64, 1, 7)
        _bn(s, "signal_encoder.bn1", 64)
        _conv(s, "signal_encoder.conv2", 128, 64, 5)
        _bn(s, "signal_encoder.bn2", 128)
        _conv(s, "signal_encoder.conv3", 256, 128, 3)
        _bn(s, "signal_encoder.bn3", 256)
        _lin(s, "signal_encoder.fc", d, 256)
        s["sequence_transformer.pos_encoder.pe"] = ((1, 5000, d), "pe", 0)
        for i in range(nl):
            _tel(s, f"sequence_transformer.transformer_encoder.layers.{i}", d, dff)
        _rnn(s, "context_aggregator.gru", 3, d, d // 2)
        _lin(s, "context_aggregator.projection", d, d)
        _lin(s, "anomaly_detector.anomaly_net.0", 64, 2 * d)
        _lin(s, "anomaly_detector.anomaly_net.3", 32, 64)
        _lin(s, "anomaly_detector.anomaly_net.5", 1, 32)
        _lin(s, "detection_head.class_head.0", d // 2, d)
        _lin(s, "detection_head.class_head.3", num_classes, d // 2)
        _lin(s, "detection_head.position_head.0", d // 2, d)
        _lin(s, "detection_head.position_head.3", 2, d // 2)
        _lin(s, "health_extractor.0", d // 2, d)
        _lin(s, "health_extractor.2", d // 4, d // 2)
        _lin(s, "health_extractor.4", d, d // 4)
        _lin(s, "attention.0", d // 4, d)
        _lin(s, "attention.2", 1, d // 4)
    elif kind == "enhanced":
        d = d_model or 256
        nl = num_layers or 6
        dff = dim_feedforward or 1024
        hd = 128  # EnhancedAnomalyDetector hidden_dim (enhanced_model.py:483)
        _conv(s, "signal_encoder.conv_init.0", 64, 1, 7)
        _bn(s, "signal_encoder.conv_init.1", 64)
        for b in (1, 2, 3, 4):
            _conv(s, f"signal_encoder.multi_scale.branch{b}", 32, 64, 3)
        _conv(s, "signal_encoder.multi_scale.combine.0", 128, 128, 1)
        _bn(s, "signal_encoder.multi_scale.combine.1", 128)
        for r in range(3):
            _conv(s, f"signal_encoder.res_blocks.{r}.conv_block.0", 128, 128, 3)
            _bn(s, f"signal_encoder.res_blocks.{r}.conv_block.1", 128)
            _conv(s, f"signal_encoder.res_blocks.{r}.conv_block.3", 128, 128, 3)
            _bn(s, f"signal_encoder.res_blocks.{r}.conv_block.4", 128)
        _conv(s, "signal_encoder.pyramid_1", 256, 128, 3)
        _bn(s, "signal_encoder.pyramid_bn1", 256)
        _conv(s, "signal_encoder.pyramid_2", 256, 256, 3)
        _bn(s, "signal_encoder.pyramid_bn2", 256)
        _lin(s, "signal_encoder.fc.0", d, 640)
        _ln(s, "signal_encoder.fc.1", d)
        s["sequence_transformer.pos_encoder.pe"] = ((1, 5000, d), "pe", 0)
        for i in range(nl):
            _tel(s, f"sequence_transformer.layers.{i}", d, dff)
        _ln(s, "sequence_transformer.norm", d)
        s["context_aggregator.attention_query"] = ((d,), "randn", 0)
        _rnn(s, "context_aggregator.lstm", 4, d, d // 2)
        _lin(s, "context_aggregator.attention_keys", d, d)
        _lin(s, "context_aggregator.attention_values", d, d)
        _lin(s, "context_aggregator.projection.0", d, 2 * d)
        _ln(s, "context_aggregator.projection.1", d)
        _lin(s, "anomaly_detector.health_extractor.0", hd, d)
        _ln(s, "anomaly_detector.health_extractor.1", hd)
        _lin(s, "anomaly_detector.health_extractor.4", hd // 2, hd)
        _ln(s, "anomaly_detector.health_extractor.5", hd // 2)
        _lin(s, "anomaly_detector.health_extractor.7", d, hd // 2)
        _lin(s, "anomaly_detector.anomaly_net.0", hd, 2 * d)
        _ln(s, "anomaly_detector.anomaly_net.1", hd)
        _lin(s, "anomaly_detector.anomaly_net.4", hd // 2, hd)
        _ln(s, "anomaly_detector.anomaly_net.5", hd // 2)
        _lin(s, "anomaly_detector.anomaly_net.7", 1, hd // 2)
        _lin(s, "anomaly_detector.uncertainty_net.0", hd, 2 * d)
        _ln(s, "anomaly_detector.uncertainty_net.1", hd)
        _lin(s, "anomaly_detector.uncertainty_net.4", 1, hd)
        for head, unc, nout in (("class_head", "class_uncertainty", num_classes),
                                ("position_head", "position_uncertainty", 2)):
            _lin(s, f"detection_head.{head}.0", d // 2, d)
            _ln(s, f"detection_head.{head}.1", d // 2)
            _lin(s, f"detection_head.{head}.4", d // 4, d // 2)
            _ln(s, f"detection_head.{head}.5", d // 4)
            _lin(s, f"detection_head.{head}.7", nout, d // 4)
            _lin(s, f"detection_head.{unc}.0", d // 4, d)
            _ln(s, f"detection_head.{unc}.1", d // 4)
            _lin(s, f"detection_head.{unc}.3", nout, d // 4)
        _mha(s, "cross_attention", d)
        _ln(s, "cross_norm", d)
        _lin(s, "sequence_integration.0", d, 2 * d)
        _ln(s, "sequence_integration.1", d)
    elif kind == "two_stage":
        d = d_model or 128
        q = d // 4
        for name, k in (("small", 3), ("medium", 5), ("large", 7), ("xlarge", 11)):
            _conv(s, f"signal_encoder.conv_{name}.0", q, 1, k)
            _bn(s, f"signal_encoder.conv_{name}.1", q)
            _conv(s, f"signal_encoder.conv_{name}.3", q, q, k)
            _bn(s, f"signal_encoder.conv_{name}.4", q)
        _lin(s, "signal_encoder.projection.0", d, d)
        _ln(s, "signal_encoder.projection.1", d)
        s["sequence_transformer.pos_encoder.pe"] = ((1, 5000, d), "pe", 0)
        for i in range(4):
            _tel(s, f"sequence_transformer.transformer_encoder.layers.{i}", d, 512)
        _ln(s, "sequence_transformer.norm", d)
        for mod, sub in (("defect_classifier", "classifier"), ("defect_classifier", "uncertainty"),
                         ("position_predictor", "position_predictor"), ("position_predictor", "uncertainty")):
            _lin(s, f"{mod}.{sub}.0", 64, d)
            _ln(s, f"{mod}.{sub}.1", 64)
            _lin(s, f"{mod}.{sub}.4", 2, 64)
    elif kind == "msc_legacy":
        h0, h1, h2 = hidden_sizes
        _lin(s, "shared_layer.0", h0, signal_length)
        _lin(s, "shared_layer.2", h1, h0)
        _mha(s, "attention", h1)
        _lin(s, "classifier.0", h2, h1)
        _lin(s, "classifier.2", 1, h2)
    elif kind in ("improved", "hybrid"):
        hyb = kind == "hybrid"
        h0, h1, h2 = hidden_sizes
        nl = num_layers or 4
        if not hyb:
            _conv(s, "conv1d.0", 16, 1, 3)
            _bn(s, "conv1d.1", 16)
            _conv(s, "conv1d.3", 32, 16, 3)
            _bn(s, "conv1d.4", 32)
            _conv(s, "background_extractor", 32, 1, 15)
            _lin(s, "shared_layer.0", h0, signal_length)
        else:
            _conv(s, "conv_layers.0", 32, 1, 3)
            _bn(s, "conv_layers.1", 32)
            _conv(s, "conv_layers.3", 64, 32, 3)
            _bn(s, "conv_layers.4", 64)
            _conv(s, "conv_layers.6", 64, 64, 5)
            _bn(s, "conv_layers.7", 64)
            _lin(s, "shared_layer.0", h0, 256)
        _lin(s, "shared_layer.3", h1, h0)
        s["position_encoding.encoding"] = ((1200 if hyb else 300, h1), "randn", 0)
        for i in range(nl):
            t = f"transformer_layers.{i}."
            _mha(s, t + "self_attn", h1)
            _conv(s, t + "local_attn.local_conv", h1, 1, 11 if hyb else 9)
            if hyb:
                _conv(s, t + "local_attn.local_conv2", h1, 1, 5)
            _lin(s, t + "ffn.0", h2, h1)
            _lin(s, t + "ffn.3", h1, h2)
            for j in (1, 2, 3):
                _ln(s, f"{t}norm{j}", h1)
        _lin(s, "classifier", 1 if hyb else 3, h1)
    elif kind == "complex":
        d = d_model or 64
        nl = num_layers or 4
        s["positional_encoding"] = ((300, d), "randn", 0)
        _conv(s, "conv_layers.0", 32, 1, 3)
        _bn(s, "conv_layers.1", 32)
        _conv(s, "conv_layers.3", 64, 32, 7)
        _bn(s, "conv_layers.4", 64)
        _conv(s, "conv_layers.6", 64, 64, 15)
        _bn(s, "conv_layers.7", 64)
        _lin(s, "feature_projection.0", d, 128)
        for i in range(nl):
            _tel(s, f"transformer.layers.{i}", d, 2 * d)
        _lin(s, "detection_head.0", d // 2, d)
        _lin(s, "detection_head.3", 1, d // 2)
    else:
        raise ValueError(f"unknown model kind {kind!r}")
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
