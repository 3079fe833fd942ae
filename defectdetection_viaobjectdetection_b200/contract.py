"""state_dict contract of the reference classes this package replaces.

``state_spec(kind, **cfg)`` returns ordered ``(key, shape, role)`` entries in the reference's
registration order, so ``state_dict()`` of the drop-in modules lists the same keys in the same
order and ``load_state_dict(..., strict=True)`` of a reference checkpoint succeeds.
Roles: ``w`` weight matrix / conv kernel, ``b`` bias, ``nw``/``nb`` norm scale/shift,
``bn_mean``/``bn_var``/``bn_count`` BatchNorm buffers, ``pe`` sinusoidal buffer, ``randn`` learned table.
"""
from __future__ import annotations

BUFFER_ROLES = ("bn_mean", "bn_var", "bn_count", "pe")


class _Spec(list):
    def lin(self, n, out_f, in_f):
        self.append((n + ".weight", (out_f, in_f), "w"))
        self.append((n + ".bias", (out_f,), "b"))

    def conv(self, n, cout, cin, k):
        self.append((n + ".weight", (cout, cin, k), "w"))
        self.append((n + ".bias", (cout,), "b"))

    def bn(self, n, c):
        self.append((n + ".weight", (c,), "nw"))
        self.append((n + ".bias", (c,), "nb"))
        self.append((n + ".running_mean", (c,), "bn_mean"))
        self.append((n + ".running_var", (c,), "bn_var"))
        self.append((n + ".num_batches_tracked", (), "bn_count"))

    def ln(self, n, c):
        self.append((n + ".weight", (c,), "nw"))
        self.append((n + ".bias", (c,), "nb"))

    def mha(self, n, d):
        self.append((n + ".in_proj_weight", (3 * d, d), "w"))
        self.append((n + ".in_proj_bias", (3 * d,), "b"))
        self.append((n + ".out_proj.weight", (d, d), "w"))
        self.append((n + ".out_proj.bias", (d,), "b"))

    def tel(self, n, d, dff):
        self.mha(n + ".self_attn", d)
        self.lin(n + ".linear1", dff, d)
        self.lin(n + ".linear2", d, dff)
        self.ln(n + ".norm1", d)
        self.ln(n + ".norm2", d)

    def rnn(self, n, gates, in0, hidden):
        for layer in range(2):
            in_f = in0 if layer == 0 else 2 * hidden
            for suf in ("", "_reverse"):
                self.append((f"{n}.weight_ih_l{layer}{suf}", (gates * hidden, in_f), "w"))
                self.append((f"{n}.weight_hh_l{layer}{suf}", (gates * hidden, hidden), "w"))
                self.append((f"{n}.bias_ih_l{layer}{suf}", (gates * hidden,), "b"))
                self.append((f"{n}.bias_hh_l{layer}{suf}", (gates * hidden,), "b"))


def state_spec(kind, signal_length=320, hidden_sizes=None, d_model=None, num_classes=2,
               num_layers=None, dim_feedforward=None, num_heads=None):
    s = _Spec()
    hidden_sizes = tuple(hidden_sizes or ((256, 128, 48) if kind == "hybrid" else (128, 64, 32)))
    if kind in ("msc", "msc_n"):                      # signals/multisignalNN/NN_models.py:45-128, :198-246
        h0, h1, h2 = hidden_sizes
        s.conv("conv1d.0", 8, 1, 3)
        s.conv("conv1d.2", 16, 8, 3)
        if kind == "msc_n":
            s.conv("background_extractor", 16, 1, 11)
        s.lin("shared_layer.0", h0, signal_length)
        s.lin("shared_layer.2", h1, h0)
        s.append(("position_encoding.encoding", (300, h1), "randn"))
        s.mha("transformer_encoder.self_attn", h1)
        if kind == "msc":
            s.mha("transformer_encoder.cross_attn", h1)
        else:
            s.conv("transformer_encoder.local_attn.local_conv", h1, 1, 5)
        s.lin("transformer_encoder.ffn.0", h2, h1)
        s.lin("transformer_encoder.ffn.2", h1, h2)
        for i in (1, 2, 3):
            s.ln(f"transformer_encoder.norm{i}", h1)
        s.lin("classifier", 3, h1)
    elif kind == "conv1d_msc":                        # signals/MSC_Conv1D_training.py:50-76
        s.conv("feature_extractor.0", 64, 1, 3)
        s.conv("feature_extractor.2", 128, 64, 3)
        s.conv("feature_extractor.4", 128, 128, 1)
        for i in range(4):
            s.tel(f"transformer_encoder.layers.{i}", 128, 2048)
        s.lin("classifier.0", 64, 128)
        s.lin("classifier.2", 1, 64)
    elif kind == "ssd":                               # SignalSequenceDetection/model.py:230-285
        d, nl, dff = d_model or 128, num_layers or 4, dim_feedforward or 512
        s.conv("signal_encoder.conv1", 64, 1, 7)
        s.bn("signal_encoder.bn1", 64)
        s.conv("signal_encoder.conv2", 128, 64, 5)
        s.bn("signal_encoder.bn2", 128)
        s.conv("signal_encoder.conv3", 256, 128, 3)
        s.bn("signal_encoder.bn3", 256)
        s.lin("signal_encoder.fc", d, 256)
        s.append(("sequence_transformer.pos_encoder.pe", (1, 5000, d), "pe"))
        for i in range(nl):
            s.tel(f"sequence_transformer.transformer_encoder.layers.{i}", d, dff)
        s.rnn("context_aggregator.gru", 3, d, d // 2)
        s.lin("context_aggregator.projection", d, d)
        s.lin("anomaly_detector.anomaly_net.0", 64, 2 * d)
        s.lin("anomaly_detector.anomaly_net.3", 32, 64)
        s.lin("anomaly_detector.anomaly_net.5", 1, 32)
        s.lin("detection_head.class_head.0", d // 2, d)
        s.lin("detection_head.class_head.3", num_classes, d // 2)
        s.lin("detection_head.position_head.0", d // 2, d)
        s.lin("detection_head.position_head.3", 2, d // 2)
        s.lin("health_extractor.0", d // 2, d)
        s.lin("health_extractor.2", d // 4, d // 2)
        s.lin("health_extractor.4", d, d // 4)
        s.lin("attention.0", d // 4, d)
        s.lin("attention.2", 1, d // 4)
    elif kind == "enhanced":                          # SignalSequenceDetection/enhanced_model.py:449-500
        d, nl, dff = d_model or 256, num_layers or 6, dim_feedforward or 1024
        hd = 128
        e = "signal_encoder."
        s.conv(e + "conv_init.0", 64, 1, 7)
        s.bn(e + "conv_init.1", 64)
        for b in (1, 2, 3, 4):
            s.conv(f"{e}multi_scale.branch{b}", 32, 64, 3)
        s.conv(e + "multi_scale.combine.0", 128, 128, 1)
        s.bn(e + "multi_scale.combine.1", 128)
        for r in range(3):
            q = f"{e}res_blocks.{r}.conv_block."
            s.conv(q + "0", 128, 128, 3)
            s.bn(q + "1", 128)
            s.conv(q + "3", 128, 128, 3)
            s.bn(q + "4", 128)
        s.conv(e + "pyramid_1", 256, 128, 3)
        s.bn(e + "pyramid_bn1", 256)
        s.conv(e + "pyramid_2", 256, 256, 3)
        s.bn(e + "pyramid_bn2", 256)
        s.lin(e + "fc.0", d, 640)
        s.ln(e + "fc.1", d)
        s.append(("sequence_transformer.pos_encoder.pe", (1, 5000, d), "pe"))
        for i in range(nl):
            s.tel(f"sequence_transformer.layers.{i}", d, dff)
        s.ln("sequence_transformer.norm", d)
        s.append(("context_aggregator.attention_query", (d,), "randn"))
        s.rnn("context_aggregator.lstm", 4, d, d // 2)
        s.lin("context_aggregator.attention_keys", d, d)
        s.lin("context_aggregator.attention_values", d, d)
        s.lin("context_aggregator.projection.0", d, 2 * d)
        s.ln("context_aggregator.projection.1", d)
        a = "anomaly_detector."
        s.lin(a + "health_extractor.0", hd, d)
        s.ln(a + "health_extractor.1", hd)
        s.lin(a + "health_extractor.4", hd // 2, hd)
        s.ln(a + "health_extractor.5", hd // 2)
        s.lin(a + "health_extractor.7", d, hd // 2)
        s.lin(a + "anomaly_net.0", hd, 2 * d)
        s.ln(a + "anomaly_net.1", hd)
        s.lin(a + "anomaly_net.4", hd // 2, hd)
        s.ln(a + "anomaly_net.5", hd // 2)
        s.lin(a + "anomaly_net.7", 1, hd // 2)
        s.lin(a + "uncertainty_net.0", hd, 2 * d)
        s.ln(a + "uncertainty_net.1", hd)
        s.lin(a + "uncertainty_net.4", 1, hd)
        h = "detection_head."
        for head, unc, nout in (("class_head", "class_uncertainty", num_classes),
                                ("position_head", "position_uncertainty", 2)):
            s.lin(f"{h}{head}.0", d // 2, d)
            s.ln(f"{h}{head}.1", d // 2)
            s.lin(f"{h}{head}.4", d // 4, d // 2)
            s.ln(f"{h}{head}.5", d // 4)
            s.lin(f"{h}{head}.7", nout, d // 4)
            s.lin(f"{h}{unc}.0", d // 4, d)
            s.ln(f"{h}{unc}.1", d // 4)
            s.lin(f"{h}{unc}.3", nout, d // 4)
        s.mha("cross_attention", d)
        s.ln("cross_norm", d)
        s.lin("sequence_integration.0", d, 2 * d)
        s.ln("sequence_integration.1", d)
    elif kind == "two_stage":                         # SignalSequenceDetection/two_stage_model.py:254-271
        d = d_model or 128
        q = d // 4
        for name, k in (("small", 3), ("medium", 5), ("large", 7), ("xlarge", 11)):
            p = f"signal_encoder.conv_{name}."
            s.conv(p + "0", q, 1, k)
            s.bn(p + "1", q)
            s.conv(p + "3", q, q, k)
            s.bn(p + "4", q)
        s.lin("signal_encoder.projection.0", d, d)
        s.ln("signal_encoder.projection.1", d)
        s.append(("sequence_transformer.pos_encoder.pe", (1, 5000, d), "pe"))
        for i in range(4):
            s.tel(f"sequence_transformer.transformer_encoder.layers.{i}", d, 512)
        s.ln("sequence_transformer.norm", d)
        for m in ("defect_classifier.classifier", "defect_classifier.uncertainty",
                  "position_predictor.position_predictor", "position_predictor.uncertainty"):
            s.lin(m + ".0", 64, d)
            s.ln(m + ".1", 64)
            s.lin(m + ".4", 2, 64)
    elif kind == "msc_legacy":                        # signals/resaveModelOnnx.py:7-22
        h0, h1, h2 = hidden_sizes
        s.lin("shared_layer.0", h0, signal_length)
        s.lin("shared_layer.2", h1, h0)
        s.mha("attention", h1)
        s.lin("classifier.0", h2, h1)
        s.lin("classifier.2", 1, h2)
    elif kind in ("improved", "hybrid"):              # improved_model.py:69-121, hybrid_binary.py:83-134
        hyb = kind == "hybrid"
        h0, h1, h2 = hidden_sizes
        if not hyb:
            s.conv("conv1d.0", 16, 1, 3)
            s.bn("conv1d.1", 16)
            s.conv("conv1d.3", 32, 16, 3)
            s.bn("conv1d.4", 32)
            s.conv("background_extractor", 32, 1, 15)
            s.lin("shared_layer.0", h0, signal_length)
        else:
            s.conv("conv_layers.0", 32, 1, 3)
            s.bn("conv_layers.1", 32)
            s.conv("conv_layers.3", 64, 32, 3)
            s.bn("conv_layers.4", 64)
            s.conv("conv_layers.6", 64, 64, 5)
            s.bn("conv_layers.7", 64)
            s.lin("shared_layer.0", h0, 256)
        s.lin("shared_layer.3", h1, h0)
        s.append(("position_encoding.encoding", (1200 if hyb else 300, h1), "randn"))
        for i in range(num_layers or 4):
            t = f"transformer_layers.{i}."
            s.mha(t + "self_attn", h1)
            s.conv(t + "local_attn.local_conv", h1, 1, 11 if hyb else 9)
            if hyb:
                s.conv(t + "local_attn.local_conv2", h1, 1, 5)
            s.lin(t + "ffn.0", h2, h1)
            s.lin(t + "ffn.3", h1, h2)
            for j in (1, 2, 3):
                s.ln(f"{t}norm{j}", h1)
        s.lin("classifier", 1 if hyb else 3, h1)
    elif kind == "complex":                           # detection_models/complex_detection_model.py:6-61
        d = d_model or 64
        s.append(("positional_encoding", (300, d), "randn"))
        s.conv("conv_layers.0", 32, 1, 3)
        s.bn("conv_layers.1", 32)
        s.conv("conv_layers.3", 64, 32, 7)
        s.bn("conv_layers.4", 64)
        s.conv("conv_layers.6", 64, 64, 15)
        s.bn("conv_layers.7", 64)
        s.lin("feature_projection.0", d, 128)
        for i in range(num_layers or 4):
            s.tel(f"transformer.layers.{i}", d, 2 * d)
        s.lin("detection_head.0", d // 2, d)
        s.lin("detection_head.3", 1, d // 2)
    else:
        raise ValueError(f"unknown model kind {kind!r}")
    return list(s)
