"""Functional fp32 CPU restatement of the reference forwards (test infrastructure).

Every function takes a ``state_dict`` with the reference's keys plus the input
tensor and returns what the reference ``forward`` returns in eval mode
(dropout = identity, BatchNorm = running statistics).  No ``nn.Module`` of the
reference is used: only ``torch.nn.functional`` primitives and explicit loops
for the recurrences, so each step that a CUDA kernel must reproduce is spelled
out.  Citations are relative to the reference repository root.

``precision='bf16'`` emulates the library's bf16-I/O mode the way SURVEY.md
§2.2 prescribes: fp32 math on bf16-rounded inputs and weights (matrix
operands), used only to size the 1e-2 tolerance in tests.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

BN_EPS = 1e-5
LN_EPS = 1e-5


# --------------------------------------------------------------------------- primitives
def linear(sd, name, x):
    return F.linear(x, sd[name + ".weight"], sd[name + ".bias"])


def layer_norm(sd, name, x):
    w = sd[name + ".weight"]
    return F.layer_norm(x, (w.numel(),), w, sd[name + ".bias"], LN_EPS)


def conv1d(sd, name, x, padding=0, dilation=1, stride=1, groups=1):
    return F.conv1d(x, sd[name + ".weight"], sd[name + ".bias"], stride=stride,
                    padding=padding, dilation=dilation, groups=groups)


def bn_eval(sd, name, x):
    """BatchNorm1d in eval mode: (x-mean)/sqrt(var+eps)*gamma+beta over channel dim 1."""
    mean = sd[name + ".running_mean"][None, :, None]
    var = sd[name + ".running_var"][None, :, None]
    g = sd[name + ".weight"][None, :, None]
    b = sd[name + ".bias"][None, :, None]
    return (x - mean) / torch.sqrt(var + BN_EPS) * g + b


def mha(sd, name, q_in, kv_in, nhead, need_weights=False):
    """nn.MultiheadAttention(batch_first=True), eval: packed in_proj rows [q|k|v],
    scale 1/sqrt(head_dim), softmax over keys, weights averaged over heads."""
    B, Nq, D = q_in.shape
    Nk = kv_in.shape[1]
    w, b = sd[name + ".in_proj_weight"], sd[name + ".in_proj_bias"]
    q = F.linear(q_in, w[:D], b[:D])
    k = F.linear(kv_in, w[D:2 * D], b[D:2 * D])
    v = F.linear(kv_in, w[2 * D:], b[2 * D:])
    hd = D // nhead
    q = q.view(B, Nq, nhead, hd).transpose(1, 2)
    k = k.view(B, Nk, nhead, hd).transpose(1, 2)
    v = v.view(B, Nk, nhead, hd).transpose(1, 2)
    s = torch.matmul(q, k.transpose(-1, -2)) * (1.0 / math.sqrt(hd))
    p = torch.softmax(s, dim=-1)
    o = torch.matmul(p, v).transpose(1, 2).reshape(B, Nq, D)
    o = F.linear(o, sd[name + ".out_proj.weight"], sd[name + ".out_proj.bias"])
    return o, (p.mean(dim=1) if need_weights else None)


def encoder_layer(sd, name, x, nhead, act="relu", need_weights=False):
    """Post-norm nn.TransformerEncoderLayer (norm_first=False): x = LN1(x + SA(x));
    x = LN2(x + W2 act(W1 x)).  Also enhanced_model.py:195-208 (act = exact-erf GELU)."""
    a, w = mha(sd, name + ".self_attn", x, x, nhead, need_weights)
    x = layer_norm(sd, name + ".norm1", x + a)
    h = linear(sd, name + ".linear1", x)
    h = F.relu(h) if act == "relu" else F.gelu(h)
    x = layer_norm(sd, name + ".norm2", x + linear(sd, name + ".linear2", h))
    return x, w


def gru_bidir(sd, name, x, hidden, layers=2):
    """nn.GRU(batch_first, bidirectional); gate rows [r|z|n];
    n = tanh(W_in x + b_in + r*(W_hn h + b_hn)); h' = (1-z)*n + z*h."""
    B, T, _ = x.shape
    inp = x
    for layer in range(layers):
        outs = []
        for suffix, order in (("", range(T)), ("_reverse", range(T - 1, -1, -1))):
            wi = sd[f"{name}.weight_ih_l{layer}{suffix}"]
            wh = sd[f"{name}.weight_hh_l{layer}{suffix}"]
            bi = sd[f"{name}.bias_ih_l{layer}{suffix}"]
            bh = sd[f"{name}.bias_hh_l{layer}{suffix}"]
            gi_all = F.linear(inp, wi, bi)
            h = x.new_zeros(B, hidden)
            out = x.new_zeros(B, T, hidden)
            for t in order:
                gi = gi_all[:, t]
                gh = F.linear(h, wh, bh)
                r = torch.sigmoid(gi[:, :hidden] + gh[:, :hidden])
                z = torch.sigmoid(gi[:, hidden:2 * hidden] + gh[:, hidden:2 * hidden])
                n = torch.tanh(gi[:, 2 * hidden:] + r * gh[:, 2 * hidden:])
                h = (1.0 - z) * n + z * h
                out[:, t] = h
            outs.append(out)
        inp = torch.cat(outs, dim=-1)
    return inp


def lstm_bidir(sd, name, x, hidden, layers=2):
    """nn.LSTM(batch_first, bidirectional); gate rows [i|f|g|o];
    c' = sig(f)*c + sig(i)*tanh(g); h' = sig(o)*tanh(c')."""
    B, T, _ = x.shape
    inp = x
    for layer in range(layers):
        outs = []
        for suffix, order in (("", range(T)), ("_reverse", range(T - 1, -1, -1))):
            wi = sd[f"{name}.weight_ih_l{layer}{suffix}"]
            wh = sd[f"{name}.weight_hh_l{layer}{suffix}"]
            bi = sd[f"{name}.bias_ih_l{layer}{suffix}"]
            bh = sd[f"{name}.bias_hh_l{layer}{suffix}"]
            gi_all = F.linear(inp, wi, bi)
            h = x.new_zeros(B, hidden)
            c = x.new_zeros(B, hidden)
            out = x.new_zeros(B, T, hidden)
            for t in order:
                g = gi_all[:, t] + F.linear(h, wh, bh)
                i_g = torch.sigmoid(g[:, :hidden])
                f_g = torch.sigmoid(g[:, hidden:2 * hidden])
                g_g = torch.tanh(g[:, 2 * hidden:3 * hidden])
                o_g = torch.sigmoid(g[:, 3 * hidden:])
                c = f_g * c + i_g * g_g
                h = o_g * torch.tanh(c)
                out[:, t] = h
            outs.append(out)
        inp = torch.cat(outs, dim=-1)
    return inp


def _prep(sd, x, precision):
    """fp32 oracle, or the bf16-I/O emulation: round input and every matrix weight to bf16."""
    sd = {k: (v.float() if v.is_floating_point() else v) for k, v in sd.items()}
    x = x.float()
    if precision == "bf16":
        x = x.to(torch.bfloat16).float()
        sd = {k: (v.to(torch.bfloat16).float() if (v.is_floating_point() and v.dim() >= 2 and not k.endswith(".pe")
                                                   and "encoding" not in k) else v)
              for k, v in sd.items()}
    elif precision != "fp32":
        raise ValueError(precision)
    return sd, x


# --------------------------------------------------------------------------- signals/ family
def _msc_front(sd, x, with_background):
    """NN_models.py:111-116 / :227-235: per A-scan Conv1d 1->8 k3 ReLU, 8->16 k3 ReLU,
    [MSC_N: minus depthwise k11 background], mean over the 16 channels, MLP S->h0->h1 (ReLU)."""
    B, N, S = x.shape
    h = x.reshape(B * N, 1, S)
    h = F.relu(conv1d(sd, "conv1d.0", h, padding=1))
    h = F.relu(conv1d(sd, "conv1d.2", h, padding=1))
    if with_background:
        h = h - conv1d(sd, "background_extractor", h, padding=5, groups=16)
    h = h.mean(dim=1)
    h = F.relu(linear(sd, "shared_layer.0", h))
    h = F.relu(linear(sd, "shared_layer.2", h))
    h = h.view(B, N, -1)
    return h + sd["position_encoding.encoding"][:N][None]          # NN_models.py:11-14


def _msc_head(sd, h):
    o = linear(sd, "classifier", h)                                  # NN_models.py:123-127
    return torch.sigmoid(o[..., 0]), torch.tanh(o[..., 1]) * 0.5 + 0.5, torch.tanh(o[..., 2]) * 0.5 + 0.5


def msc_forward(sd, x, num_heads=4, precision="fp32"):
    """MultiSignalClassifier.forward, NN_models.py:108-128 with TransformerEncoder.forward :31-42."""
    sd, x = _prep(sd, x, precision)
    h = _msc_front(sd, x, with_background=False)
    te = "transformer_encoder"
    a, _ = mha(sd, te + ".self_attn", h, h, num_heads)
    h = layer_norm(sd, te + ".norm1", h + a)
    shifted = torch.cat([h[:, 1:], h[:, -1:]], dim=1)                # NN_models.py:35
    a, _ = mha(sd, te + ".cross_attn", h, shifted, num_heads)
    h = layer_norm(sd, te + ".norm2", h + a)
    f = linear(sd, te + ".ffn.2", F.relu(linear(sd, te + ".ffn.0", h)))
    h = layer_norm(sd, te + ".norm3", h + f)
    return _msc_head(sd, h)


def msc_n_forward(sd, x, num_heads=4, precision="fp32"):
    """MultiSignalClassifier_N.forward, NN_models.py:225-246 with TransformerEncoder_N.forward :184-195
    and LocalAttention_N.forward :162-167 (depthwise k5 conv along the set axis)."""
    sd, x = _prep(sd, x, precision)
    h = _msc_front(sd, x, with_background=True)
    te = "transformer_encoder"
    a, _ = mha(sd, te + ".self_attn", h, h, num_heads)
    h = layer_norm(sd, te + ".norm1", h + a)
    loc = conv1d(sd, te + ".local_attn.local_conv", h.permute(0, 2, 1), padding=2, groups=h.shape[-1])
    h = layer_norm(sd, te + ".norm2", h + loc.permute(0, 2, 1))
    f = linear(sd, te + ".ffn.2", F.relu(linear(sd, te + ".ffn.0", h)))
    h = layer_norm(sd, te + ".norm3", h + f)
    return _msc_head(sd, h)


def conv1d_msc_forward(sd, x, nhead=4, precision="fp32"):
    """DefectDetectionModel.forward, MSC_Conv1D_training.py:78-89.  x is [B, S, N]."""
    sd, x = _prep(sd, x, precision)
    B, S, N = x.shape
    h = x.permute(0, 2, 1).contiguous().view(-1, 1, S)
    h = F.relu(conv1d(sd, "feature_extractor.0", h, padding=1))
    h = F.relu(conv1d(sd, "feature_extractor.2", h, padding=1))
    h = F.relu(conv1d(sd, "feature_extractor.4", h))
    h = h.mean(dim=2).view(B, N, -1)
    for i in range(4):
        h, _ = encoder_layer(sd, f"transformer_encoder.layers.{i}", h, nhead)
    h = F.relu(linear(sd, "classifier.0", h))
    return torch.sigmoid(linear(sd, "classifier.2", h)).squeeze(-1)


# --------------------------------------------------------------------------- SignalSequenceDetection
def _n_layers(sd, prefix):
    idx = {int(k[len(prefix):].split(".")[0]) for k in sd if k.startswith(prefix)}
    return max(idx) + 1


def ssd_forward(sd, x, nhead=8, precision="fp32"):
    """SignalSequenceDetector.forward, model.py:287-343 (targets=None)."""
    sd, x = _prep(sd, x, precision)
    B, N, S = x.shape
    h = x.reshape(B * N, 1, S)                                                     # model.py:65
    h = F.relu(bn_eval(sd, "signal_encoder.bn1", conv1d(sd, "signal_encoder.conv1", h, padding=3)))
    h = F.relu(bn_eval(sd, "signal_encoder.bn2", conv1d(sd, "signal_encoder.conv2", h, padding=2)))
    h = F.relu(bn_eval(sd, "signal_encoder.bn3", conv1d(sd, "signal_encoder.conv3", h, padding=1)))
    h = linear(sd, "signal_encoder.fc", h.mean(dim=2)).view(B, N, -1)              # model.py:73-79
    d = h.shape[-1]
    seq = h + sd["sequence_transformer.pos_encoder.pe"][:, :N]                     # model.py:31
    pre = "sequence_transformer.transformer_encoder.layers."
    for i in range(_n_layers(sd, pre)):
        seq, _ = encoder_layer(sd, f"{pre}{i}", seq, nhead)
    ctx = gru_bidir(sd, "context_aggregator.gru", seq, d // 2)                     # model.py:186
    ctx = linear(sd, "context_aggregator.projection", ctx)
    health = linear(sd, "health_extractor.4", F.relu(linear(sd, "health_extractor.2",
                    F.relu(linear(sd, "health_extractor.0", seq)))))
    att = linear(sd, "attention.2", F.relu(linear(sd, "attention.0", seq)))
    att = torch.softmax(att, dim=1)                                                # model.py:314 (over N)
    enh = seq * att + ctx                                                          # model.py:317
    comb = torch.cat([enh, health], dim=-1)
    an = F.relu(linear(sd, "anomaly_detector.anomaly_net.0", comb))
    an = F.relu(linear(sd, "anomaly_detector.anomaly_net.3", an))
    an = torch.sigmoid(linear(sd, "anomaly_detector.anomaly_net.5", an))
    logits = linear(sd, "detection_head.class_head.3", F.relu(linear(sd, "detection_head.class_head.0", enh)))
    pos = torch.sigmoid(linear(sd, "detection_head.position_head.3",
                               F.relu(linear(sd, "detection_head.position_head.0", enh))))
    if logits.shape[-1] > 1:                                                       # model.py:327-334
        logits = logits.clone()
        logits[:, :, 1:] = logits[:, :, 1:] + an
    return {"class_preds": logits, "position_preds": pos, "anomaly_scores": an, "attention_weights": att}


def _mlp_ln_gelu(sd, name, x, idx):
    """Linear -> LayerNorm -> GELU(erf) at indices (idx, idx+1) of an nn.Sequential."""
    return F.gelu(layer_norm(sd, f"{name}.{idx + 1}", linear(sd, f"{name}.{idx}", x)))


def enhanced_encoder(sd, x):
    """EnhancedSignalEncoder.forward, enhanced_model.py:135-175."""
    B, N, S = x.shape
    p = "signal_encoder."
    h = x.reshape(B * N, 1, S)
    h = F.relu(bn_eval(sd, p + "conv_init.1", conv1d(sd, p + "conv_init.0", h, padding=3)))
    br = [conv1d(sd, f"{p}multi_scale.branch{b + 1}", h, padding=2 ** b, dilation=2 ** b) for b in range(4)]
    h = torch.cat(br, dim=1)                                                       # :82-89
    h = F.relu(bn_eval(sd, p + "multi_scale.combine.1", conv1d(sd, p + "multi_scale.combine.0", h)))
    for r, dil in enumerate((1, 2, 4)):                                            # :110-114, :54-58
        q = f"{p}res_blocks.{r}.conv_block."
        o = F.relu(bn_eval(sd, q + "1", conv1d(sd, q + "0", h, padding=dil, dilation=dil)))
        o = bn_eval(sd, q + "4", conv1d(sd, q + "3", o, padding=dil, dilation=dil))
        h = F.relu(o + h)
    f0 = h.mean(dim=2)
    x1 = F.relu(bn_eval(sd, p + "pyramid_bn1", conv1d(sd, p + "pyramid_1", h, padding=1, stride=2)))
    f1 = x1.mean(dim=2)
    x2 = F.relu(bn_eval(sd, p + "pyramid_bn2", conv1d(sd, p + "pyramid_2", x1, padding=1, stride=2)))
    f2 = x2.mean(dim=2)
    f = torch.cat([f0, f1, f2], dim=1)
    return F.relu(layer_norm(sd, p + "fc.1", linear(sd, p + "fc.0", f))).view(B, N, -1)


def enhanced_forward(sd, x, nhead=8, precision="fp32"):
    """EnhancedSignalSequenceDetector.forward, enhanced_model.py:502-566 (targets=None)."""
    sd, x = _prep(sd, x, precision)
    B, N, S = x.shape
    sf = enhanced_encoder(sd, x)
    d = sf.shape[-1]
    seq = sf + sd["sequence_transformer.pos_encoder.pe"][:, :N]                    # :27
    attn_ws = []
    pre = "sequence_transformer.layers."
    for i in range(_n_layers(sd, pre)):                                            # :240-246
        seq, w = encoder_layer(sd, f"{pre}{i}", seq, nhead, act="gelu", need_weights=True)
        attn_ws.append(w)
    seq = layer_norm(sd, "sequence_transformer.norm", seq)
    # EnhancedContextAggregator.forward :283-313
    ca = "context_aggregator."
    lo = lstm_bidir(sd, ca + "lstm", seq, d // 2)
    keys = linear(sd, ca + "attention_keys", lo)
    scores = torch.sum(keys * sd[ca + "attention_query"][None, None, :], dim=-1)
    cw = torch.softmax(scores, dim=1)
    vals = linear(sd, ca + "attention_values", lo) * cw.unsqueeze(-1)
    ctx = layer_norm(sd, ca + "projection.1", linear(sd, ca + "projection.0", torch.cat([lo, vals], dim=-1)))
    # cross attention :525-528 (always 8 heads, :491)
    co, cross_w = mha(sd, "cross_attention", sf, ctx, 8, need_weights=True)
    cf = layer_norm(sd, "cross_norm", sf + co)
    integ = _mlp_ln_gelu(sd, "sequence_integration", torch.cat([cf, seq], dim=-1), 0)
    # EnhancedAnomalyDetector.forward :358-380
    ad = "anomaly_detector."
    hl = _mlp_ln_gelu(sd, ad + "health_extractor", seq, 0)
    hl = _mlp_ln_gelu(sd, ad + "health_extractor", hl, 4)
    hl = linear(sd, ad + "health_extractor.7", hl)
    comb = torch.cat([integ, hl], dim=-1)
    an = _mlp_ln_gelu(sd, ad + "anomaly_net", comb, 0)
    an = _mlp_ln_gelu(sd, ad + "anomaly_net", an, 4)
    an = torch.sigmoid(linear(sd, ad + "anomaly_net.7", an))
    au = F.softplus(linear(sd, ad + "uncertainty_net.4", _mlp_ln_gelu(sd, ad + "uncertainty_net", comb, 0)))
    # EnhancedDefectDetectionHead.forward :433-448
    dh = "detection_head."
    cl = _mlp_ln_gelu(sd, dh + "class_head", integ, 0)
    cl = linear(sd, dh + "class_head.7", _mlp_ln_gelu(sd, dh + "class_head", cl, 4))
    cu = F.softplus(linear(sd, dh + "class_uncertainty.3", _mlp_ln_gelu(sd, dh + "class_uncertainty", integ, 0)))
    po = _mlp_ln_gelu(sd, dh + "position_head", integ, 0)
    po = torch.sigmoid(linear(sd, dh + "position_head.7", _mlp_ln_gelu(sd, dh + "position_head", po, 4)))
    pu = F.softplus(linear(sd, dh + "position_uncertainty.3",
                           _mlp_ln_gelu(sd, dh + "position_uncertainty", integ, 0)))
    if cl.shape[-1] > 1:                                                           # :545-552
        cl = cl.clone()
        cl[:, :, 1:] = cl[:, :, 1:] + an
    return {"class_preds": cl, "class_uncertainty": cu, "position_preds": po, "position_uncertainty": pu,
            "anomaly_scores": an, "anomaly_uncertainty": au, "attention_weights": attn_ws,
            "context_attention": cw, "cross_attention": cross_w}


def two_stage_forward(sd, x, nhead=8, precision="fp32"):
    """TwoStageDefectDetector.forward, two_stage_model.py:273-312 (targets=None)."""
    sd, x = _prep(sd, x, precision)
    B, N, S = x.shape
    h = x.reshape(B * N, 1, S)
    feats = []
    for name, k in (("small", 3), ("medium", 5), ("large", 7), ("xlarge", 11)):   # :102-114
        q = f"signal_encoder.conv_{name}."
        o = F.relu(bn_eval(sd, q + "1", conv1d(sd, q + "0", h, padding=k // 2)))
        o = F.relu(bn_eval(sd, q + "4", conv1d(sd, q + "3", o, padding=k // 2)))
        feats.append(o.mean(dim=2))
    f = torch.cat(feats, dim=1).view(B, N, -1)
    f = layer_norm(sd, "signal_encoder.projection.1", linear(sd, "signal_encoder.projection.0", f))
    seq = f + sd["sequence_transformer.pos_encoder.pe"][:, :N]
    pre = "sequence_transformer.transformer_encoder.layers."
    for i in range(_n_layers(sd, pre)):
        seq, _ = encoder_layer(sd, f"{pre}{i}", seq, nhead)
    seq = layer_norm(sd, "sequence_transformer.norm", seq)                         # :165

    def head(name):
        return linear(sd, name + ".4", F.relu(layer_norm(sd, name + ".1", linear(sd, name + ".0", seq))))

    logits = head("defect_classifier.classifier")
    dunc = F.softplus(head("defect_classifier.uncertainty")) + 1e-6               # :210
    pos = torch.sigmoid(head("position_predictor.position_predictor"))
    punc = F.softplus(head("position_predictor.uncertainty")) + 1e-6              # :249
    probs = torch.softmax(logits, dim=-1)
    return {"defect_logits": logits, "defect_probs": probs, "defect_uncertainty": dunc,
            "position_preds": pos * probs[:, :, 1:2], "position_uncertainty": punc,
            "attention_weights": None}


# --------------------------------------------------------------------------- SURVEY section 8 "next" rows
def msc_legacy_forward(sd, x, precision="fp32"):
    """Legacy no-conv MultiSignalClassifier.forward, signals/resaveModelOnnx.py:24-33 (identical copies in
    GNN_testing_multi_v2_MAP.py:16-36, teststtt.py): MLP, one self-attention (4 heads) with no residual, MLP head."""
    sd, x = _prep(sd, x, precision)
    h = F.relu(linear(sd, "shared_layer.0", x))
    h = F.relu(linear(sd, "shared_layer.2", h))
    a, _ = mha(sd, "attention", h, h, 4)
    o = torch.sigmoid(linear(sd, "classifier.2", F.relu(linear(sd, "classifier.0", a))))
    return o.squeeze(-1)


def _local_encoder_layer(sd, t, h, num_heads, kernels):
    """improved_model.py:55-67 / hybrid_binary.py:65-80: self-attention -> LN -> depthwise conv(s) along the set
    axis (LocalAttention) -> LN -> FFN -> LN, dropout = identity."""
    a, _ = mha(sd, t + "self_attn", h, h, num_heads)
    h = layer_norm(sd, t + "norm1", h + a)
    loc = h.permute(0, 2, 1)
    for name, k in kernels:
        loc = conv1d(sd, t + "local_attn." + name, loc, padding=k // 2, groups=loc.shape[1])
    h = layer_norm(sd, t + "norm2", h + loc.permute(0, 2, 1))
    f = linear(sd, t + "ffn.3", F.relu(linear(sd, t + "ffn.0", h)))
    return layer_norm(sd, t + "norm3", h + f)


def improved_forward(sd, x, num_heads=8, precision="fp32"):
    """ImprovedMultiSignalClassifier.forward, improved_model.py:123-157."""
    sd, x = _prep(sd, x, precision)
    B, N, S = x.shape
    h = x.reshape(B * N, 1, S)
    h = F.relu(bn_eval(sd, "conv1d.1", conv1d(sd, "conv1d.0", h, padding=1)))
    h = F.relu(bn_eval(sd, "conv1d.4", conv1d(sd, "conv1d.3", h, padding=1)))
    h = h - conv1d(sd, "background_extractor", h, padding=7, groups=32)          # :130-131
    h = h.mean(dim=1)
    h = F.relu(linear(sd, "shared_layer.0", h))
    h = F.relu(linear(sd, "shared_layer.3", h)).view(B, N, -1)
    h = h + sd["position_encoding.encoding"][:N][None]
    for i in range(_n_layers(sd, "transformer_layers.")):
        h = _local_encoder_layer(sd, f"transformer_layers.{i}.", h, num_heads, (("local_conv", 9),))
    o = linear(sd, "classifier", h)
    return torch.sigmoid(o[..., 0]), torch.clamp(o[..., 1], 0.0, 1.0), torch.clamp(o[..., 2], 0.0, 1.0)


def _conv_stack3(sd, h, pads):
    for idx, pad in zip((0, 3, 6), pads):
        h = F.relu(bn_eval(sd, f"conv_layers.{idx + 1}", conv1d(sd, f"conv_layers.{idx}", h, padding=pad)))
    return h


def hybrid_forward(sd, x, num_heads=8, precision="fp32"):
    """HybridBinaryModel.forward, hybrid_binary.py:136-168."""
    sd, x = _prep(sd, x, precision)
    B, N, S = x.shape
    h = _conv_stack3(sd, x.reshape(B * N, 1, S), (1, 1, 2))
    k = max(S // 128, 1)                                                         # :108-111
    h = F.avg_pool1d(h, kernel_size=k, stride=k)
    h = F.interpolate(h, size=128, mode="linear", align_corners=False)           # :145
    seq = h.mean(dim=1).view(B, N, -1)
    seq = torch.cat([seq, seq - seq.mean(dim=1, keepdim=True)], dim=-1)          # :147-149
    h = F.relu(linear(sd, "shared_layer.0", seq))
    h = F.relu(linear(sd, "shared_layer.3", h))
    h = h + sd["position_encoding.encoding"][:N][None]
    for i in range(_n_layers(sd, "transformer_layers.")):
        h = _local_encoder_layer(sd, f"transformer_layers.{i}.", h, num_heads, (("local_conv", 11), ("local_conv2", 5)))
    return torch.sigmoid(linear(sd, "classifier", h).squeeze(-1))


def complex_forward(sd, x, num_heads=8, precision="fp32"):
    """ComplexDetectionModel.forward, complex_detection_model.py:63-96."""
    sd, x = _prep(sd, x, precision)
    B, N, S = x.shape
    h = _conv_stack3(sd, x.reshape(B * N, 1, S), (1, 3, 7))
    h = F.adaptive_avg_pool1d(h, 128).mean(dim=1)                                # :71-75
    h = F.relu(linear(sd, "feature_projection.0", h)).view(B, N, -1)
    h = h + sd["positional_encoding"][:N][None]
    for i in range(_n_layers(sd, "transformer.layers.")):
        h, _ = encoder_layer(sd, f"transformer.layers.{i}", h, num_heads)
    h = F.relu(linear(sd, "detection_head.0", h))
    return torch.sigmoid(linear(sd, "detection_head.3", h).squeeze(-1))


FORWARD = {
    "msc_legacy": msc_legacy_forward,
    "improved": improved_forward,
    "hybrid": hybrid_forward,
    "complex": complex_forward,
    "msc": msc_forward,
    "msc_n": msc_n_forward,
    "conv1d_msc": conv1d_msc_forward,
    "ssd": ssd_forward,
    "enhanced": enhanced_forward,
    "two_stage": two_stage_forward,
}
