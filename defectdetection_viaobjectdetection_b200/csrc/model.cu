// state_dict contract, weight packing (BatchNorm folding, kernel layouts) and the forward graphs.
#include "model.cuh"

#include <algorithm>
#include <cmath>
#include <utility>
#include <cstdlib>
#include <cstring>

namespace paut {

// ------------------------------------------------------------------------------------------ contract
namespace {
struct SpecBuilder {
  std::vector<KeySpec>& s;
  void add(const std::string& k, std::vector<int64_t> shape) { s.push_back({k, std::move(shape)}); }
  void lin(const std::string& n, int out, int in) { add(n + ".weight", {out, in}); add(n + ".bias", {out}); }
  void conv(const std::string& n, int co, int ci, int k) { add(n + ".weight", {co, ci, k}); add(n + ".bias", {co}); }
  void bn(const std::string& n, int c) {
    add(n + ".weight", {c}); add(n + ".bias", {c}); add(n + ".running_mean", {c}); add(n + ".running_var", {c});
    add(n + ".num_batches_tracked", {});
  }
  void ln(const std::string& n, int c) { add(n + ".weight", {c}); add(n + ".bias", {c}); }
  void mha(const std::string& n, int d) {
    add(n + ".in_proj_weight", {3 * d, d}); add(n + ".in_proj_bias", {3 * d});
    add(n + ".out_proj.weight", {d, d}); add(n + ".out_proj.bias", {d});
  }
  void tel(const std::string& n, int d, int dff) {
    mha(n + ".self_attn", d); lin(n + ".linear1", dff, d); lin(n + ".linear2", d, dff);
    ln(n + ".norm1", d); ln(n + ".norm2", d);
  }
  void rnn(const std::string& n, int gates, int in0, int hidden) {
    for (int layer = 0; layer < 2; ++layer) {
      const int in_f = layer == 0 ? in0 : 2 * hidden;
      for (const char* suf : {"", "_reverse"}) {
        const std::string l = "_l" + std::to_string(layer) + suf;
        add(n + ".weight_ih" + l, {gates * hidden, in_f});
        add(n + ".weight_hh" + l, {gates * hidden, hidden});
        add(n + ".bias_ih" + l, {gates * hidden});
        add(n + ".bias_hh" + l, {gates * hidden});
      }
    }
  }
};
std::string istr(int i) { return std::to_string(i); }
}  // namespace

void Model::build_spec() {
  spec.clear();
  SpecBuilder b{spec};
  const int S = cfg.signal_length, d = cfg.d_model, C = cfg.num_classes;
  switch (kind) {
    case PAUT_MODEL_MSC:
    case PAUT_MODEL_MSC_N: {
      const int h0 = cfg.hidden_sizes[0], h1 = cfg.hidden_sizes[1], h2 = cfg.hidden_sizes[2];
      b.conv("conv1d.0", 8, 1, 3);
      b.conv("conv1d.2", 16, 8, 3);
      if (kind == PAUT_MODEL_MSC_N) b.conv("background_extractor", 16, 1, 11);
      b.lin("shared_layer.0", h0, S);
      b.lin("shared_layer.2", h1, h0);
      b.add("position_encoding.encoding", {300, h1});
      b.mha("transformer_encoder.self_attn", h1);
      if (kind == PAUT_MODEL_MSC) b.mha("transformer_encoder.cross_attn", h1);
      else b.conv("transformer_encoder.local_attn.local_conv", h1, 1, 5);
      b.lin("transformer_encoder.ffn.0", h2, h1);
      b.lin("transformer_encoder.ffn.2", h1, h2);
      for (int i = 1; i <= 3; ++i) b.ln("transformer_encoder.norm" + istr(i), h1);
      b.lin("classifier", 3, h1);
      break;
    }
    case PAUT_MODEL_CONV1D_MSC:
      b.conv("feature_extractor.0", 64, 1, 3);
      b.conv("feature_extractor.2", 128, 64, 3);
      b.conv("feature_extractor.4", 128, 128, 1);
      for (int i = 0; i < 4; ++i) b.tel("transformer_encoder.layers." + istr(i), 128, 2048);
      b.lin("classifier.0", 64, 128);
      b.lin("classifier.2", 1, 64);
      break;
    case PAUT_MODEL_SSD:
      b.conv("signal_encoder.conv1", 64, 1, 7);  b.bn("signal_encoder.bn1", 64);
      b.conv("signal_encoder.conv2", 128, 64, 5); b.bn("signal_encoder.bn2", 128);
      b.conv("signal_encoder.conv3", 256, 128, 3); b.bn("signal_encoder.bn3", 256);
      b.lin("signal_encoder.fc", d, 256);
      b.add("sequence_transformer.pos_encoder.pe", {1, 5000, d});
      for (int i = 0; i < cfg.num_layers; ++i)
        b.tel("sequence_transformer.transformer_encoder.layers." + istr(i), d, cfg.dim_feedforward);
      b.rnn("context_aggregator.gru", 3, d, d / 2);
      b.lin("context_aggregator.projection", d, d);
      b.lin("anomaly_detector.anomaly_net.0", 64, 2 * d);
      b.lin("anomaly_detector.anomaly_net.3", 32, 64);
      b.lin("anomaly_detector.anomaly_net.5", 1, 32);
      b.lin("detection_head.class_head.0", d / 2, d);
      b.lin("detection_head.class_head.3", C, d / 2);
      b.lin("detection_head.position_head.0", d / 2, d);
      b.lin("detection_head.position_head.3", 2, d / 2);
      b.lin("health_extractor.0", d / 2, d);
      b.lin("health_extractor.2", d / 4, d / 2);
      b.lin("health_extractor.4", d, d / 4);
      b.lin("attention.0", d / 4, d);
      b.lin("attention.2", 1, d / 4);
      break;
    case PAUT_MODEL_ENHANCED: {
      const int hd = 128;
      const std::string e = "signal_encoder.";
      b.conv(e + "conv_init.0", 64, 1, 7); b.bn(e + "conv_init.1", 64);
      for (int i = 1; i <= 4; ++i) b.conv(e + "multi_scale.branch" + istr(i), 32, 64, 3);
      b.conv(e + "multi_scale.combine.0", 128, 128, 1); b.bn(e + "multi_scale.combine.1", 128);
      for (int r = 0; r < 3; ++r) {
        const std::string q = e + "res_blocks." + istr(r) + ".conv_block.";
        b.conv(q + "0", 128, 128, 3); b.bn(q + "1", 128); b.conv(q + "3", 128, 128, 3); b.bn(q + "4", 128);
      }
      b.conv(e + "pyramid_1", 256, 128, 3); b.bn(e + "pyramid_bn1", 256);
      b.conv(e + "pyramid_2", 256, 256, 3); b.bn(e + "pyramid_bn2", 256);
      b.lin(e + "fc.0", d, 640); b.ln(e + "fc.1", d);
      b.add("sequence_transformer.pos_encoder.pe", {1, 5000, d});
      for (int i = 0; i < cfg.num_layers; ++i) b.tel("sequence_transformer.layers." + istr(i), d, cfg.dim_feedforward);
      b.ln("sequence_transformer.norm", d);
      b.add("context_aggregator.attention_query", {d});
      b.rnn("context_aggregator.lstm", 4, d, d / 2);
      b.lin("context_aggregator.attention_keys", d, d);
      b.lin("context_aggregator.attention_values", d, d);
      b.lin("context_aggregator.projection.0", d, 2 * d); b.ln("context_aggregator.projection.1", d);
      const std::string a = "anomaly_detector.";
      b.lin(a + "health_extractor.0", hd, d); b.ln(a + "health_extractor.1", hd);
      b.lin(a + "health_extractor.4", hd / 2, hd); b.ln(a + "health_extractor.5", hd / 2);
      b.lin(a + "health_extractor.7", d, hd / 2);
      b.lin(a + "anomaly_net.0", hd, 2 * d); b.ln(a + "anomaly_net.1", hd);
      b.lin(a + "anomaly_net.4", hd / 2, hd); b.ln(a + "anomaly_net.5", hd / 2);
      b.lin(a + "anomaly_net.7", 1, hd / 2);
      b.lin(a + "uncertainty_net.0", hd, 2 * d); b.ln(a + "uncertainty_net.1", hd);
      b.lin(a + "uncertainty_net.4", 1, hd);
      const std::string h = "detection_head.";
      const char* heads[2] = {"class_head", "position_head"};
      const char* uncs[2] = {"class_uncertainty", "position_uncertainty"};
      const int nouts[2] = {C, 2};
      for (int i = 0; i < 2; ++i) {
        b.lin(h + heads[i] + ".0", d / 2, d); b.ln(h + heads[i] + ".1", d / 2);
        b.lin(h + heads[i] + ".4", d / 4, d / 2); b.ln(h + heads[i] + ".5", d / 4);
        b.lin(h + heads[i] + ".7", nouts[i], d / 4);
        b.lin(h + uncs[i] + ".0", d / 4, d); b.ln(h + uncs[i] + ".1", d / 4);
        b.lin(h + uncs[i] + ".3", nouts[i], d / 4);
      }
      b.mha("cross_attention", d);
      b.ln("cross_norm", d);
      b.lin("sequence_integration.0", d, 2 * d); b.ln("sequence_integration.1", d);
      break;
    }
    case PAUT_MODEL_TWO_STAGE: {
      const int q = d / 4;
      const char* names[4] = {"small", "medium", "large", "xlarge"};
      const int ks[4] = {3, 5, 7, 11};
      for (int i = 0; i < 4; ++i) {
        const std::string p = std::string("signal_encoder.conv_") + names[i] + ".";
        b.conv(p + "0", q, 1, ks[i]); b.bn(p + "1", q); b.conv(p + "3", q, q, ks[i]); b.bn(p + "4", q);
      }
      b.lin("signal_encoder.projection.0", d, d); b.ln("signal_encoder.projection.1", d);
      b.add("sequence_transformer.pos_encoder.pe", {1, 5000, d});
      for (int i = 0; i < 4; ++i) b.tel("sequence_transformer.transformer_encoder.layers." + istr(i), d, 512);
      b.ln("sequence_transformer.norm", d);
      const char* mods[4] = {"defect_classifier.classifier", "defect_classifier.uncertainty",
                             "position_predictor.position_predictor", "position_predictor.uncertainty"};
      for (const char* m : mods) {
        b.lin(std::string(m) + ".0", 64, d); b.ln(std::string(m) + ".1", 64); b.lin(std::string(m) + ".4", 2, 64);
      }
      break;
    }
    case PAUT_MODEL_MSC_LEGACY: {                     // resaveModelOnnx.py:8-22
      const int h0 = cfg.hidden_sizes[0], h1 = cfg.hidden_sizes[1], h2 = cfg.hidden_sizes[2];
      b.lin("shared_layer.0", h0, S);
      b.lin("shared_layer.2", h1, h0);
      b.mha("attention", h1);
      b.lin("classifier.0", h2, h1);
      b.lin("classifier.2", 1, h2);
      break;
    }
    case PAUT_MODEL_IMPROVED:                         // improved_model.py:70-121
    case PAUT_MODEL_HYBRID: {                         // hybrid_binary.py:88-134
      const int h0 = cfg.hidden_sizes[0], h1 = cfg.hidden_sizes[1], h2 = cfg.hidden_sizes[2];
      const bool hyb = kind == PAUT_MODEL_HYBRID;
      if (!hyb) {
        b.conv("conv1d.0", 16, 1, 3); b.bn("conv1d.1", 16);
        b.conv("conv1d.3", 32, 16, 3); b.bn("conv1d.4", 32);
        b.conv("background_extractor", 32, 1, 15);
        b.lin("shared_layer.0", h0, S);
      } else {
        b.conv("conv_layers.0", 32, 1, 3); b.bn("conv_layers.1", 32);
        b.conv("conv_layers.3", 64, 32, 3); b.bn("conv_layers.4", 64);
        b.conv("conv_layers.6", 64, 64, 5); b.bn("conv_layers.7", 64);
        b.lin("shared_layer.0", h0, 256);
      }
      b.lin("shared_layer.3", h1, h0);
      b.add("position_encoding.encoding", {hyb ? 1200 : 300, h1});
      for (int i = 0; i < cfg.num_layers; ++i) {
        const std::string t = "transformer_layers." + istr(i) + ".";
        b.mha(t + "self_attn", h1);
        b.conv(t + "local_attn.local_conv", h1, 1, hyb ? 11 : 9);
        if (hyb) b.conv(t + "local_attn.local_conv2", h1, 1, 5);
        b.lin(t + "ffn.0", h2, h1);
        b.lin(t + "ffn.3", h1, h2);
        for (int j = 1; j <= 3; ++j) b.ln(t + "norm" + istr(j), h1);
      }
      b.lin("classifier", hyb ? 1 : 3, h1);
      break;
    }
    case PAUT_MODEL_COMPLEX:                          // complex_detection_model.py:11-61
      b.add("positional_encoding", {300, d});
      b.conv("conv_layers.0", 32, 1, 3); b.bn("conv_layers.1", 32);
      b.conv("conv_layers.3", 64, 32, 7); b.bn("conv_layers.4", 64);
      b.conv("conv_layers.6", 64, 64, 15); b.bn("conv_layers.7", 64);
      b.lin("feature_projection.0", d, 128);
      for (int i = 0; i < cfg.num_layers; ++i) b.tel("transformer.layers." + istr(i), d, 2 * d);
      b.lin("detection_head.0", d / 2, d);
      b.lin("detection_head.3", 1, d / 2);
      break;
    default:
      throw Error(PAUT_ERR_INVALID, "unknown model kind");
  }
}

// ------------------------------------------------------------------------------------------ tensors
Model::~Model() {
  for (void* p : dev_allocs) cudaFree(p);
}

void Model::set_tensor(const char* key, const void* ptr, int dtype, const int64_t* shape, int ndim) {
  PAUT_CHECK(key && ptr, PAUT_ERR_INVALID, "set_tensor: null key or pointer");
  const KeySpec* ks = nullptr;
  for (const auto& k : spec)
    if (k.key == key) { ks = &k; break; }
  PAUT_CHECK(ks, PAUT_ERR_INVALID, std::string("set_tensor: unexpected key '") + key + "'");
  PAUT_CHECK(ndim == (int)ks->shape.size(), PAUT_ERR_INVALID, std::string("set_tensor: rank mismatch for ") + key);
  int64_t n = 1;
  for (int i = 0; i < ndim; ++i) {
    PAUT_CHECK(shape[i] == ks->shape[i], PAUT_ERR_INVALID, std::string("set_tensor: shape mismatch for ") + key);
    n *= shape[i];
  }
  HostTensor t;
  t.shape.assign(shape, shape + ndim);
  t.data.resize(n);
  if (dtype == PAUT_F32) {
    PAUT_CUDA(cudaMemcpy(t.data.data(), ptr, n * sizeof(float), cudaMemcpyDefault));
  } else if (dtype == PAUT_I64) {
    std::vector<int64_t> tmp(n);
    PAUT_CUDA(cudaMemcpy(tmp.data(), ptr, n * sizeof(int64_t), cudaMemcpyDefault));
    for (int64_t i = 0; i < n; ++i) t.data[i] = (float)tmp[i];
  } else if (dtype == PAUT_BF16) {
    std::vector<uint16_t> tmp(n);
    PAUT_CUDA(cudaMemcpy(tmp.data(), ptr, n * sizeof(uint16_t), cudaMemcpyDefault));
    for (int64_t i = 0; i < n; ++i) {
      uint32_t u = (uint32_t)tmp[i] << 16;
      std::memcpy(&t.data[i], &u, 4);
    }
  } else {
    throw Error(PAUT_ERR_INVALID, "set_tensor: dtype must be F32, BF16 or I64");
  }
  host[key] = std::move(t);
  finalized = false;
}

const HostTensor& Model::H(const std::string& key) const {
  auto it = host.find(key);
  if (it == host.end()) throw Error(PAUT_ERR_MISSING, "state_dict tensor not set: " + key);
  return it->second;
}

const float* Model::upload(const std::vector<float>& v) {
  void* p = nullptr;
  PAUT_CUDA(cudaMalloc(&p, std::max<size_t>(v.size(), 4) * sizeof(float)));
  dev_allocs.push_back(p);
  PAUT_CUDA(cudaMemcpy(p, v.data(), v.size() * sizeof(float), cudaMemcpyHostToDevice));
  return static_cast<const float*>(p);
}

Lin Model::make_lin(const std::vector<float>& W, const std::vector<float>& bias, int rows, int K) {
  std::vector<float> Wt((size_t)rows * K);
  for (int n = 0; n < rows; ++n)
    for (int k = 0; k < K; ++k) Wt[(size_t)k * rows + n] = W[(size_t)n * K + k];
  Lin l;
  l.K = K;
  l.N = rows;
  l.W = upload(W);
  l.Wt = upload(Wt);
  l.b = upload(bias);
  pack_tc(l, W);
  return l;
}

Lin Model::pack_lin_rows(const std::string& wkey, const std::string& bkey, int row0, int rows) {
  const HostTensor& w = H(wkey);
  const HostTensor& b = H(bkey);
  const int K = (int)w.shape[1];
  std::vector<float> W(w.data.begin() + (size_t)row0 * K, w.data.begin() + (size_t)(row0 + rows) * K);
  std::vector<float> bias(b.data.begin() + row0, b.data.begin() + row0 + rows);
  return make_lin(W, bias, rows, K);
}

// bf16 mode: additionally pack the weight for the tcgen05 GEMM (chunked K-major bf16)
void Model::pack_tc(Lin& l, const std::vector<float>& W) {
  if (cfg.precision != PAUT_PRECISION_BF16) return;
  if (lin_res_ln_supported(l.N, l.K, l.K)) {           // candidates for the fused Linear + residual + LayerNorm kernel
    std::vector<uint16_t> hb(W.size());
    for (size_t i = 0; i < W.size(); ++i) {
      uint32_t u;
      memcpy(&u, &W[i], 4);
      hb[i] = (uint16_t)((u + 0x7fffu + ((u >> 16) & 1u)) >> 16);
    }
    void* dp = nullptr;
    PAUT_CUDA(cudaMalloc(&dp, hb.size() * sizeof(uint16_t)));
    dev_allocs.push_back(dp);
    PAUT_CUDA(cudaMemcpy(dp, hb.data(), hb.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
    l.Wrow = dp;
  }
  const int nt = tc_pick_ntile(l.N, l.K, &l.stream_b);
  if (nt == 0) return;
  std::vector<uint16_t> packed;
  int Kp = 0;
  tc_pack_weight(W.data(), l.N, l.K, nt, packed, &Kp);
  void* p = nullptr;
  PAUT_CUDA(cudaMalloc(&p, packed.size() * sizeof(uint16_t)));
  dev_allocs.push_back(p);
  PAUT_CUDA(cudaMemcpy(p, packed.data(), packed.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
  l.Wp = p;
  l.NT = nt;
}

Lin Model::pack_lin(const std::string& name) {
  return pack_lin_rows(name + ".weight", name + ".bias", 0, (int)H(name + ".weight").shape[0]);
}

// Conv1d (+ optional eval BatchNorm folded in): y = conv(x) * scale + shift,
// scale = gamma / sqrt(var + eps), shift = (bias - mean) * scale + beta.
ConvW Model::pack_conv(const std::string& cn, const std::string& bn, int max_dil, bool stride2) {
  const HostTensor& w = H(cn + ".weight");
  const HostTensor& b = H(cn + ".bias");
  const int Cout = (int)w.shape[0], Cin = (int)w.shape[1], taps = (int)w.shape[2];
  std::vector<float> scale(Cout, 1.f), shift(Cout);
  for (int c = 0; c < Cout; ++c) shift[c] = b.data[c];
  if (!bn.empty()) {
    const HostTensor& g = H(bn + ".weight");
    const HostTensor& be = H(bn + ".bias");
    const HostTensor& mu = H(bn + ".running_mean");
    const HostTensor& var = H(bn + ".running_var");
    for (int c = 0; c < Cout; ++c) {
      const double sc = (double)g.data[c] / std::sqrt((double)var.data[c] + 1e-5);
      scale[c] = (float)sc;
      shift[c] = (float)(((double)b.data[c] - (double)mu.data[c]) * sc + (double)be.data[c]);
    }
  }
  std::vector<float> p((size_t)taps * Cin * Cout);
  for (int co = 0; co < Cout; ++co)
    for (int ci = 0; ci < Cin; ++ci)
      for (int t = 0; t < taps; ++t)
        p[((size_t)t * Cin + ci) * Cout + co] = w.data[((size_t)co * Cin + ci) * taps + t] * scale[co];
  ConvW cw;
  cw.Cin = Cin;
  cw.Cout = Cout;
  cw.taps = taps;
  cw.w = upload(p);
  cw.shift = upload(shift);
  auto up16 = [&](const std::vector<uint16_t>& v) {
    void* d = nullptr;
    PAUT_CUDA(cudaMalloc(&d, v.size() * sizeof(uint16_t)));
    dev_allocs.push_back(d);
    PAUT_CUDA(cudaMemcpy(d, v.data(), v.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
    return d;
  };
  if (cfg.precision == PAUT_PRECISION_BF16 && conv_tc_plan(Cin, taps, Cout, max_dil, &cw.NT, &cw.CB)) {
    std::vector<uint16_t> packed;
    conv_tc_pack(p.data(), taps, Cin, Cout, cw.NT, cw.CB, packed);
    cw.Wp = up16(packed);
  }
  if (cfg.precision == PAUT_PRECISION_BF16 && stride2 && taps == 3 &&
      conv_tc_plan(2 * Cin, 2, Cout, 1, &cw.NT_s2, &cw.CB_s2)) {
    // stride 2 as a 2-tap stride-1 conv over row pairs: tap 0 = [0 | W0], tap 1 = [W1 | W2]   (ops_conv_tc.cu)
    std::vector<float> p2((size_t)2 * 2 * Cin * Cout, 0.f);
    for (int ci = 0; ci < Cin; ++ci)
      for (int co = 0; co < Cout; ++co) {
        p2[((size_t)0 * 2 * Cin + Cin + ci) * Cout + co] = p[((size_t)0 * Cin + ci) * Cout + co];
        p2[((size_t)1 * 2 * Cin + ci) * Cout + co] = p[((size_t)1 * Cin + ci) * Cout + co];
        p2[((size_t)1 * 2 * Cin + Cin + ci) * Cout + co] = p[((size_t)2 * Cin + ci) * Cout + co];
      }
    std::vector<uint16_t> packed;
    conv_tc_pack(p2.data(), 2, 2 * Cin, Cout, cw.NT_s2, cw.CB_s2, packed);
    cw.Wp_s2 = up16(packed);
  }
  return cw;
}

LNW Model::pack_ln(const std::string& name) {
  LNW l;
  l.D = (int)H(name + ".weight").shape[0];
  l.g = upload(H(name + ".weight").data);
  l.b = upload(H(name + ".bias").data);
  return l;
}

MHAW Model::pack_mha(const std::string& name, int heads) {
  MHAW m;
  m.D = (int)H(name + ".in_proj_weight").shape[1];
  m.H = heads;
  PAUT_CHECK(heads > 0 && m.D % heads == 0, PAUT_ERR_INVALID, "embed_dim must be divisible by num_heads");
  PAUT_CHECK(m.D / heads <= 64, PAUT_ERR_UNSUPPORTED, "attention head_dim must be <= 64");
  m.hd = m.D / heads;
  m.Dp = m.D;
  if (cfg.precision == PAUT_PRECISION_BF16 && m.hd == 8) {
    // bf16 mode, head_dim 8 (d_model 64 with 8 heads: improved_model.py:70, complex_detection_model.py:11): the
    // tensor-core attention kernel works on head_dim 16 / 32, so every head is zero-padded to 16 through the
    // weights -- in_proj emits [q | 0], [k | 0], [v | 0] per head (q.k and P.v are unchanged by zero columns),
    // out_proj gets zero columns for the padding.  The kernel's 1/sqrt(16) becomes the 1/sqrt(8) of the model by
    // scaling the q rows (and q bias) by sqrt(2).
    const HostTensor& w = H(name + ".in_proj_weight");
    const HostTensor& b = H(name + ".in_proj_bias");
    const HostTensor& wo = H(name + ".out_proj.weight");
    const int D = m.D, Dp = 2 * D;
    std::vector<float> W((size_t)3 * Dp * D, 0.f), bias((size_t)3 * Dp, 0.f), Wo((size_t)D * Dp, 0.f);
    const float qs = std::sqrt(2.0f);
    for (int part = 0; part < 3; ++part)
      for (int h = 0; h < heads; ++h)
        for (int j = 0; j < 8; ++j) {
          const int src = part * D + h * 8 + j, dst = part * Dp + h * 16 + j;
          const float sc = part == 0 ? qs : 1.f;
          bias[dst] = b.data[src] * sc;
          for (int k = 0; k < D; ++k) W[(size_t)dst * D + k] = w.data[(size_t)src * D + k] * sc;
        }
    for (int n = 0; n < D; ++n)
      for (int h = 0; h < heads; ++h)
        for (int j = 0; j < 8; ++j) Wo[(size_t)n * Dp + h * 16 + j] = wo.data[(size_t)n * D + h * 8 + j];
    m.in_proj = make_lin(W, bias, 3 * Dp, D);
    m.out_proj = make_lin(Wo, H(name + ".out_proj.bias").data, D, Dp);
    m.hd = 16;
    m.Dp = Dp;
    return m;
  }
  m.in_proj = pack_lin_rows(name + ".in_proj_weight", name + ".in_proj_bias", 0, 3 * m.D);
  m.out_proj = pack_lin(name + ".out_proj");
  return m;
}

TELW Model::pack_tel(const std::string& name, int heads) {
  TELW t;
  t.attn = pack_mha(name + ".self_attn", heads);
  t.l1 = pack_lin(name + ".linear1");
  t.l2 = pack_lin(name + ".linear2");
  t.n1 = pack_ln(name + ".norm1");
  t.n2 = pack_ln(name + ".norm2");
  return t;
}

RNNW Model::pack_rnn(const std::string& name, int layer, int G, int Hh) {
  RNNW r;
  r.G = G;
  r.H = Hh;
  const int GH = G * Hh;
  const std::string l = "_l" + std::to_string(layer);
  const HostTensor& wf = H(name + ".weight_ih" + l);
  const HostTensor& wr = H(name + ".weight_ih" + l + "_reverse");
  const int K = (int)wf.shape[1];
  std::vector<float> W((size_t)2 * GH * K), Wt((size_t)2 * GH * K), bias(2 * GH);
  for (int dir = 0; dir < 2; ++dir) {
    const HostTensor& w = dir == 0 ? wf : wr;
    const HostTensor& b = H(name + ".bias_ih" + l + (dir ? "_reverse" : ""));
    for (int n = 0; n < GH; ++n) {
      bias[dir * GH + n] = b.data[n];
      for (int k = 0; k < K; ++k) {
        const float v = w.data[(size_t)n * K + k];
        W[(size_t)(dir * GH + n) * K + k] = v;
        Wt[(size_t)k * 2 * GH + dir * GH + n] = v;
      }
    }
  }
  r.ih.K = K;
  r.ih.N = 2 * GH;
  r.ih.W = upload(W);
  r.ih.Wt = upload(Wt);
  r.ih.b = upload(bias);
  pack_tc(r.ih, W);
  std::vector<float> whh((size_t)2 * Hh * GH), bhh(2 * GH);
  for (int dir = 0; dir < 2; ++dir) {
    const HostTensor& w = H(name + ".weight_hh" + l + (dir ? "_reverse" : ""));
    const HostTensor& b = H(name + ".bias_hh" + l + (dir ? "_reverse" : ""));
    for (int n = 0; n < GH; ++n) {
      bhh[dir * GH + n] = b.data[n];
      for (int k = 0; k < Hh; ++k) whh[((size_t)dir * Hh + k) * GH + n] = w.data[(size_t)n * Hh + k];
    }
  }
  r.whh_t = upload(whh);
  r.bhh = upload(bhh);
  return r;
}

void Model::finalize() {
  for (const auto& k : spec)
    if (!host.count(k.key)) throw Error(PAUT_ERR_MISSING, "state_dict tensor not set: " + k.key);
  for (void* p : dev_allocs) cudaFree(p);
  dev_allocs.clear();
  lin.clear(); conv.clear(); ln.clear(); mha.clear(); tel.clear(); rnn.clear(); raw.clear();
  PAUT_CUDA(cudaSetDevice(ctx->device));

  auto L = [&](const std::string& n) { lin[n] = pack_lin(n); };
  auto N_ = [&](const std::string& n) { ln[n] = pack_ln(n); };
  auto C_ = [&](const std::string& n, const std::string& bn, int max_dil = 1, bool stride2 = false) {
    conv[n] = pack_conv(n, bn, max_dil, stride2);
  };
  auto R = [&](const std::string& n) { raw[n] = upload(H(n).data); };
  const int d = cfg.d_model;

  switch (kind) {
    case PAUT_MODEL_MSC:
    case PAUT_MODEL_MSC_N: {
      PAUT_CHECK(cfg.hidden_sizes[1] % cfg.num_heads == 0, PAUT_ERR_INVALID,
                 "embed_dim must be divisible by num_heads");
      for (const char* n : {"conv1d.0.weight", "conv1d.0.bias", "conv1d.2.weight", "conv1d.2.bias"}) R(n);
      if (kind == PAUT_MODEL_MSC_N) {
        R("background_extractor.weight"); R("background_extractor.bias");
        R("transformer_encoder.local_attn.local_conv.weight"); R("transformer_encoder.local_attn.local_conv.bias");
      }
      L("shared_layer.0"); L("shared_layer.2");
      R("position_encoding.encoding");
      enc_tc = EncTc{};
      if (cfg.precision == PAUT_PRECISION_BF16 && kind == PAUT_MODEL_MSC &&
          msc_encoder_tc_supported(cfg.signal_length, cfg.hidden_sizes[0], cfg.hidden_sizes[1])) {
        auto up16 = [&](const std::vector<uint16_t>& v) {
          void* p = nullptr;
          PAUT_CUDA(cudaMalloc(&p, v.size() * sizeof(uint16_t)));
          dev_allocs.push_back(p);
          PAUT_CUDA(cudaMemcpy(p, v.data(), v.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
          return p;
        };
        std::vector<uint16_t> packed;
        msc_pack_conv2(H("conv1d.2.weight").data.data(), H("conv1d.2.bias").data.data(), packed);
        enc_tc.Bc = up16(packed);
        std::vector<float> w1 = H("shared_layer.0.weight").data;
        for (float& v : w1) v *= (1.f / 32.f);        // (1/2 from relu = (y+|y|)/2) * (1/16 channel mean)
        int Kp = 0;
        tc_pack_weight(w1.data(), cfg.hidden_sizes[0], cfg.signal_length, cfg.hidden_sizes[0], packed, &Kp);
        enc_tc.W1p = up16(packed);
        tc_pack_weight(H("shared_layer.2.weight").data.data(), cfg.hidden_sizes[1], cfg.hidden_sizes[0],
                       cfg.hidden_sizes[1], packed, &Kp);
        enc_tc.W2p = up16(packed);
        enc_tc.ready = true;
      }
      mscn = MscnFront{};
      if (cfg.precision == PAUT_PRECISION_BF16 && kind == PAUT_MODEL_MSC_N && mscn_front_supported(cfg.signal_length)) {
        std::vector<uint16_t> packed;
        mscn_front_pack(H("conv1d.0.weight").data.data(), H("conv1d.0.bias").data.data(), H("conv1d.2.weight").data.data(),
                        H("background_extractor.weight").data.data(), packed);
        void* p = nullptr;
        PAUT_CUDA(cudaMalloc(&p, packed.size() * sizeof(uint16_t)));
        dev_allocs.push_back(p);
        PAUT_CUDA(cudaMemcpy(p, packed.data(), packed.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
        mscn.W = p;
        double fc = 0.0;
        for (int c2 = 0; c2 < 16; ++c2) {
          mscn.b2[c2] = H("conv1d.2.bias").data[c2];
          fc += (double)H("background_extractor.bias").data[c2];
        }
        mscn.f_const = (float)(fc / 16.0);
        mscn.ready = true;
      }
      mha["self"] = pack_mha("transformer_encoder.self_attn", cfg.num_heads);
      if (kind == PAUT_MODEL_MSC) mha["cross"] = pack_mha("transformer_encoder.cross_attn", cfg.num_heads);
      L("transformer_encoder.ffn.0"); L("transformer_encoder.ffn.2");
      for (int i = 1; i <= 3; ++i) N_("transformer_encoder.norm" + istr(i));
      L("classifier");
      set_tc = SetTc{};
      if (cfg.precision == PAUT_PRECISION_BF16 &&
          msc_set_tc_supported(1, cfg.hidden_sizes[1], cfg.num_heads, cfg.hidden_sizes[2])) {
        auto bf = [](float f) -> uint16_t {
          uint32_t u;
          std::memcpy(&u, &f, 4);
          return (uint16_t)((u + 0x7fffu + ((u >> 16) & 1u)) >> 16);
        };
        auto up_bf16 = [&](const std::vector<float>& w, size_t pad_to = 0) {
          std::vector<uint16_t> hb(std::max(w.size(), pad_to), 0);
          for (size_t i = 0; i < w.size(); ++i) hb[i] = bf(w[i]);
          void* p = nullptr;
          PAUT_CUDA(cudaMalloc(&p, hb.size() * sizeof(uint16_t)));
          dev_allocs.push_back(p);
          PAUT_CUDA(cudaMemcpy(p, hb.data(), hb.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
          return p;
        };
        const std::string te = "transformer_encoder.";
        set_tc.Wqkv_self = up_bf16(H(te + "self_attn.in_proj_weight").data);
        set_tc.Wo_self = up_bf16(H(te + "self_attn.out_proj.weight").data);
        if (kind == PAUT_MODEL_MSC) {
          set_tc.Wqkv_cross = up_bf16(H(te + "cross_attn.in_proj_weight").data);
          set_tc.Wo_cross = up_bf16(H(te + "cross_attn.out_proj.weight").data);
        }
        auto up_u16 = [&](const std::vector<uint16_t>& hb) {
          void* p = nullptr;
          PAUT_CUDA(cudaMalloc(&p, hb.size() * sizeof(uint16_t)));
          dev_allocs.push_back(p);
          PAUT_CUDA(cudaMemcpy(p, hb.data(), hb.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
          return p;
        };
        int ai = 0;
        for (const char* an : {"self_attn", "cross_attn"}) {
          if (ai == 1 && kind != PAUT_MODEL_MSC) break;
          std::vector<uint16_t> pk;
          msc_attn_tc_pack(H(te + an + ".in_proj_weight").data.data(), 0, 128, pk);
          set_tc.tc_w[ai][0] = up_u16(pk);
          msc_attn_tc_pack(H(te + an + ".in_proj_weight").data.data(), 128, 64, pk);
          set_tc.tc_w[ai][1] = up_u16(pk);
          msc_attn_tc_pack(H(te + an + ".out_proj.weight").data.data(), 0, 64, pk);
          set_tc.tc_w[ai][2] = up_u16(pk);
          ++ai;
        }
        set_tc.W1 = up_bf16(H(te + "ffn.0.weight").data);
        set_tc.W2 = up_bf16(H(te + "ffn.2.weight").data);
        set_tc.Wc = up_bf16(H("classifier.weight").data, 8 * 64);        // rows 3..7 zero
        set_tc.ready = true;
      }
      break;
    }
    case PAUT_MODEL_CONV1D_MSC:
      C_("feature_extractor.0", ""); C_("feature_extractor.2", ""); C_("feature_extractor.4", "");
      for (int i = 0; i < 4; ++i) tel.push_back(pack_tel("transformer_encoder.layers." + istr(i), 4));
      L("classifier.0"); L("classifier.2");
      break;
    case PAUT_MODEL_SSD:
      C_("signal_encoder.conv1", "signal_encoder.bn1");
      C_("signal_encoder.conv2", "signal_encoder.bn2");
      C_("signal_encoder.conv3", "signal_encoder.bn3");
      L("signal_encoder.fc");
      R("sequence_transformer.pos_encoder.pe");
      for (int i = 0; i < cfg.num_layers; ++i)
        tel.push_back(pack_tel("sequence_transformer.transformer_encoder.layers." + istr(i), cfg.num_heads));
      for (int l = 0; l < 2; ++l) rnn.push_back(pack_rnn("context_aggregator.gru", l, 3, d / 2));
      for (const char* n : {"context_aggregator.projection", "anomaly_detector.anomaly_net.0",
                            "anomaly_detector.anomaly_net.3", "anomaly_detector.anomaly_net.5",
                            "detection_head.class_head.0", "detection_head.class_head.3",
                            "detection_head.position_head.0", "detection_head.position_head.3", "health_extractor.0",
                            "health_extractor.2", "health_extractor.4", "attention.0", "attention.2"})
        L(n);
      break;
    case PAUT_MODEL_ENHANCED: {
      const std::string e = "signal_encoder.";
      C_(e + "conv_init.0", e + "conv_init.1");
      for (int i = 1; i <= 4; ++i) C_(e + "multi_scale.branch" + istr(i), "", 1 << (i - 1));
      C_(e + "multi_scale.combine.0", e + "multi_scale.combine.1");
      for (int r = 0; r < 3; ++r) {
        const std::string q = e + "res_blocks." + istr(r) + ".conv_block.";
        C_(q + "0", q + "1", 1 << r); C_(q + "3", q + "4", 1 << r);
      }
      C_(e + "pyramid_1", e + "pyramid_bn1", 1, true); C_(e + "pyramid_2", e + "pyramid_bn2", 1, true);
      enh_branches = BranchConv{};
      if (cfg.precision == PAUT_PRECISION_BF16) {
        // multi_scale.branch1..4 (64 -> 32 channels, 3 taps, dilation 1/2/4/8, no BN): one shared-input grouped launch
        std::vector<std::vector<float>> ws(4);
        std::vector<float> shift_all;
        const float* wp[4];
        int taps[4] = {3, 3, 3, 3};
        bool ok = true;
        for (int b = 0; b < 4; ++b) {
          const HostTensor& w = H(e + "multi_scale.branch" + istr(b + 1) + ".weight");
          const HostTensor& bi = H(e + "multi_scale.branch" + istr(b + 1) + ".bias");
          ok = ok && w.shape[0] == 32 && w.shape[1] == 64 && w.shape[2] == 3;
          ws[b].assign((size_t)3 * 64 * 32, 0.f);
          for (int co = 0; co < 32 && ok; ++co) {
            shift_all.push_back(bi.data[co]);
            for (int ci = 0; ci < 64; ++ci)
              for (int t = 0; t < 3; ++t) ws[b][((size_t)t * 64 + ci) * 32 + co] = w.data[((size_t)co * 64 + ci) * 3 + t];
          }
          wp[b] = ws[b].data();
        }
        if (ok) {
          std::vector<uint16_t> packed;
          conv_tc_pack_grouped(wp, taps, 4, 64, 32, packed, enh_branches.goff);
          void* dptr = nullptr;
          PAUT_CUDA(cudaMalloc(&dptr, packed.size() * sizeof(uint16_t)));
          dev_allocs.push_back(dptr);
          PAUT_CUDA(cudaMemcpy(dptr, packed.data(), packed.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
          enh_branches.Wp = dptr;
          enh_branches.shift = upload(shift_all);
          enh_branches.ready = true;
        }
      }
      L(e + "fc.0"); N_(e + "fc.1");
      R("sequence_transformer.pos_encoder.pe");
      for (int i = 0; i < cfg.num_layers; ++i)
        tel.push_back(pack_tel("sequence_transformer.layers." + istr(i), cfg.num_heads));
      N_("sequence_transformer.norm");
      R("context_aggregator.attention_query");
      for (int l = 0; l < 2; ++l) rnn.push_back(pack_rnn("context_aggregator.lstm", l, 4, d / 2));
      L("context_aggregator.attention_keys"); L("context_aggregator.attention_values");
      L("context_aggregator.projection.0"); N_("context_aggregator.projection.1");
      const std::string a = "anomaly_detector.";
      for (const char* n : {"health_extractor.0", "health_extractor.4", "health_extractor.7", "anomaly_net.0",
                            "anomaly_net.4", "anomaly_net.7", "uncertainty_net.0", "uncertainty_net.4"})
        L(a + n);
      for (const char* n : {"health_extractor.1", "health_extractor.5", "anomaly_net.1", "anomaly_net.5",
                            "uncertainty_net.1"})
        N_(a + n);
      const std::string h = "detection_head.";
      for (const char* n : {"class_head", "position_head"}) {
        L(h + n + ".0"); N_(h + n + ".1"); L(h + n + ".4"); N_(h + n + ".5"); L(h + n + ".7");
      }
      for (const char* n : {"class_uncertainty", "position_uncertainty"}) {
        L(h + n + ".0"); N_(h + n + ".1"); L(h + n + ".3");
      }
      PAUT_CHECK(d % 8 == 0, PAUT_ERR_INVALID, "cross_attention: d_model must be divisible by 8");
      lin["cross.q"] = pack_lin_rows("cross_attention.in_proj_weight", "cross_attention.in_proj_bias", 0, d);
      lin["cross.kv"] = pack_lin_rows("cross_attention.in_proj_weight", "cross_attention.in_proj_bias", d, 2 * d);
      L("cross_attention.out_proj");
      N_("cross_norm");
      L("sequence_integration.0"); N_("sequence_integration.1");
      break;
    }
    case PAUT_MODEL_TWO_STAGE: {
      for (const char* n : {"small", "medium", "large", "xlarge"}) {
        const std::string p = std::string("signal_encoder.conv_") + n + ".";
        C_(p + "0", p + "1"); C_(p + "3", p + "4");
      }
      ts_grouped = GroupedConv{};
      if (cfg.precision == PAUT_PRECISION_BF16 && (d / 4) % 16 == 0 && d <= 128) {
        // the four second convolutions (q -> q channels, 3/5/7/11 taps) as ONE grouped tcgen05 launch
        const int q = d / 4;
        std::vector<std::vector<float>> ws(4);
        std::vector<float> shift_all;
        const float* wp[4];
        int i = 0;
        for (const char* n : {"small", "medium", "large", "xlarge"}) {
          const std::string cn = std::string("signal_encoder.conv_") + n + ".3", bn = std::string("signal_encoder.conv_") + n + ".4";
          const HostTensor& w = H(cn + ".weight");
          const HostTensor& b = H(cn + ".bias");
          const HostTensor& g = H(bn + ".weight");
          const HostTensor& be = H(bn + ".bias");
          const HostTensor& mu = H(bn + ".running_mean");
          const HostTensor& var = H(bn + ".running_var");
          const int taps = (int)w.shape[2];
          ts_grouped.taps[i] = taps;
          ws[i].assign((size_t)taps * q * q, 0.f);
          for (int co = 0; co < q; ++co) {
            const double sc = (double)g.data[co] / std::sqrt((double)var.data[co] + 1e-5);
            shift_all.push_back((float)(((double)b.data[co] - (double)mu.data[co]) * sc + (double)be.data[co]));
            for (int ci = 0; ci < q; ++ci)
              for (int t = 0; t < taps; ++t)
                ws[i][((size_t)t * q + ci) * q + co] = w.data[((size_t)co * q + ci) * taps + t] * (float)sc;
          }
          wp[i] = ws[i].data();
          ++i;
        }
        std::vector<uint16_t> packed;
        conv_tc_pack_grouped(wp, ts_grouped.taps, 4, q, q, packed, ts_grouped.goff);
        void* dptr = nullptr;
        PAUT_CUDA(cudaMalloc(&dptr, packed.size() * sizeof(uint16_t)));
        dev_allocs.push_back(dptr);
        PAUT_CUDA(cudaMemcpy(dptr, packed.data(), packed.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
        ts_grouped.Wp = dptr;
        ts_grouped.shift = upload(shift_all);
        ts_grouped.ready = true;
      }
      ts_heads = TsHeads{};
      if (cfg.precision == PAUT_PRECISION_BF16 && ts_heads_supported(d, 64)) {
        const char* mods[4] = {"defect_classifier.classifier", "defect_classifier.uncertainty",
                               "position_predictor.position_predictor", "position_predictor.uncertainty"};
        for (int h = 0; h < 4; ++h) {
          const std::vector<float>& w = H(std::string(mods[h]) + ".0.weight").data;       // [64][128]
          std::vector<uint16_t> hb(w.size());
          for (size_t i = 0; i < w.size(); ++i) {
            uint32_t u;
            memcpy(&u, &w[i], 4);
            hb[i] = (uint16_t)((u + 0x7fffu + ((u >> 16) & 1u)) >> 16);
          }
          void* dp = nullptr;
          PAUT_CUDA(cudaMalloc(&dp, hb.size() * sizeof(uint16_t)));
          dev_allocs.push_back(dp);
          PAUT_CUDA(cudaMemcpy(dp, hb.data(), hb.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
          ts_heads.W0[h] = dp;
        }
        ts_heads.ready = true;
      }
      ts_enc = TsEnc{};
      if (cfg.precision == PAUT_PRECISION_BF16 && d == 128) {
        // operands of the fused encoder kernel: BatchNorm folded on the host in fp64, everything in fp16 (stem weights
        // and shifts as channel pairs for HFMA2, second convolutions in the K-major operand layout)
        std::vector<std::vector<float>> w1(4), s1(4), w2(4);
        std::vector<float> shift2;
        const float* w1p[4];
        const float* s1p[4];
        const float* w2p[4];
        const int ks[4] = {3, 5, 7, 11};
        bool ok = true;
        int i = 0;
        for (const char* n : {"small", "medium", "large", "xlarge"}) {
          const std::string base = std::string("signal_encoder.conv_") + n + ".";
          auto fold = [&](const std::string& cn, const std::string& bn, std::vector<float>& scale, std::vector<float>& shift) {
            const HostTensor& b = H(cn + ".bias");
            const HostTensor& g = H(bn + ".weight");
            const HostTensor& be = H(bn + ".bias");
            const HostTensor& mu = H(bn + ".running_mean");
            const HostTensor& var = H(bn + ".running_var");
            const int C = (int)b.data.size();
            scale.resize(C); shift.resize(C);
            for (int co = 0; co < C; ++co) {
              const double sc = (double)g.data[co] / std::sqrt((double)var.data[co] + 1e-5);
              scale[co] = (float)sc;
              shift[co] = (float)(((double)b.data[co] - (double)mu.data[co]) * sc + (double)be.data[co]);
            }
          };
          const HostTensor& wa = H(base + "0.weight");
          const HostTensor& wb = H(base + "3.weight");
          ok = ok && wa.shape[0] == 32 && wa.shape[2] == ks[i] && wb.shape[0] == 32 && wb.shape[1] == 32 && wb.shape[2] == ks[i];
          if (!ok) break;
          std::vector<float> sc1, sc2, sh2;
          fold(base + "0", base + "1", sc1, s1[i]);
          fold(base + "3", base + "4", sc2, sh2);
          shift2.insert(shift2.end(), sh2.begin(), sh2.end());
          w1[i].assign((size_t)ks[i] * 32, 0.f);
          w2[i].assign((size_t)ks[i] * 32 * 32, 0.f);
          for (int co = 0; co < 32; ++co)
            for (int t = 0; t < ks[i]; ++t) {
              w1[i][(size_t)t * 32 + co] = wa.data[(size_t)co * ks[i] + t] * sc1[co];
              for (int ci = 0; ci < 32; ++ci)
                w2[i][((size_t)t * 32 + ci) * 32 + co] = wb.data[((size_t)co * 32 + ci) * ks[i] + t] * sc2[co];
            }
          w1p[i] = w1[i].data(); s1p[i] = s1[i].data(); w2p[i] = w2[i].data();
          ++i;
        }
        if (ok) {
          std::vector<uint16_t> W2h, Wsth;
          ts_encoder_pack(w1p, s1p, w2p, Wsth, W2h);
          void* dptr = nullptr;
          PAUT_CUDA(cudaMalloc(&dptr, W2h.size() * sizeof(uint16_t)));
          dev_allocs.push_back(dptr);
          PAUT_CUDA(cudaMemcpy(dptr, W2h.data(), W2h.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
          ts_enc.W2 = dptr;
          void* wq = nullptr;
          PAUT_CUDA(cudaMalloc(&wq, Wsth.size() * sizeof(uint16_t)));
          dev_allocs.push_back(wq);
          PAUT_CUDA(cudaMemcpy(wq, Wsth.data(), Wsth.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
          ts_enc.Wst = wq;
          ts_enc.shift2 = upload(shift2);
          ts_enc.ready = true;
        }
      }
      L("signal_encoder.projection.0"); N_("signal_encoder.projection.1");
      R("sequence_transformer.pos_encoder.pe");
      for (int i = 0; i < 4; ++i)
        tel.push_back(pack_tel("sequence_transformer.transformer_encoder.layers." + istr(i), 8));
      N_("sequence_transformer.norm");
      for (const char* m : {"defect_classifier.classifier", "defect_classifier.uncertainty",
                            "position_predictor.position_predictor", "position_predictor.uncertainty"}) {
        L(std::string(m) + ".0"); N_(std::string(m) + ".1"); L(std::string(m) + ".4");
      }
      break;
    }
    case PAUT_MODEL_MSC_LEGACY:
      L("shared_layer.0"); L("shared_layer.2");
      mha["attention"] = pack_mha("attention", 4);
      L("classifier.0"); L("classifier.2");
      break;
    case PAUT_MODEL_IMPROVED:
    case PAUT_MODEL_HYBRID: {
      const bool hyb = kind == PAUT_MODEL_HYBRID;
      PAUT_CHECK(cfg.hidden_sizes[1] % cfg.num_heads == 0, PAUT_ERR_INVALID,
                 "embed_dim must be divisible by num_heads");
      if (!hyb) {
        C_("conv1d.0", "conv1d.1"); C_("conv1d.3", "conv1d.4");
        R("background_extractor.weight"); R("background_extractor.bias");
      } else {
        C_("conv_layers.0", "conv_layers.1"); C_("conv_layers.3", "conv_layers.4"); C_("conv_layers.6", "conv_layers.7");
      }
      L("shared_layer.0"); L("shared_layer.3");
      R("position_encoding.encoding");
      for (int i = 0; i < cfg.num_layers; ++i) {
        const std::string t = "transformer_layers." + istr(i) + ".";
        mha[t + "self_attn"] = pack_mha(t + "self_attn", cfg.num_heads);
        R(t + "local_attn.local_conv.weight"); R(t + "local_attn.local_conv.bias");
        if (hyb) { R(t + "local_attn.local_conv2.weight"); R(t + "local_attn.local_conv2.bias"); }
        L(t + "ffn.0"); L(t + "ffn.3");
        for (int j2 = 1; j2 <= 3; ++j2) N_(t + "norm" + istr(j2));
      }
      L("classifier");
      break;
    }
    case PAUT_MODEL_COMPLEX:
      C_("conv_layers.0", "conv_layers.1"); C_("conv_layers.3", "conv_layers.4"); C_("conv_layers.6", "conv_layers.7");
      L("feature_projection.0");
      R("positional_encoding");
      for (int i = 0; i < cfg.num_layers; ++i) tel.push_back(pack_tel("transformer.layers." + istr(i), cfg.num_heads));
      L("detection_head.0"); L("detection_head.3");
      break;
  }
  PAUT_CUDA(cudaDeviceSynchronize());
  finalized = true;
}

// ------------------------------------------------------------------------------------------ graph helpers
namespace {
struct G {
  Ctx& c;
  bool bf16 = false;     // PAUT_PRECISION_BF16: tensor-core kernels where they exist
  void attention(const float* q, int ldq, const float* k, int ldk, const float* v, int ldv, float* out, int ldo,
                 int64_t B, int Nq, int Nk, int H, int hd, bool kv_shift, float* avgw) {
    if (bf16 && attention_bf16_supported(Nq, Nk, hd, avgw != nullptr))
      op_attention_bf16(c, q, ldq, k, ldk, v, ldv, out, ldo, B, Nq, Nk, H, hd, kv_shift, avgw);
    else
      op_attention(c, q, ldq, k, ldk, v, ldv, out, ldo, B, Nq, Nk, H, hd, kv_shift, avgw);
  }
  float* linear(const float* A, int lda, const Lin& L, int64_t M, int act = ACT_NONE, const float* res = nullptr,
                int ldr = 0, float* out = nullptr, int ldc = 0, int coff = 0, float eps = 0.f,
                const float* table = nullptr, int mod = 1) {
    LinArgs a;
    a.A = A; a.lda = lda; a.Wt = L.Wt; a.W = L.W; a.bias = L.b; a.M = M; a.K = L.K; a.N = L.N;
    if (!out) { out = c.allocf((size_t)M * L.N); ldc = L.N; coff = 0; }
    a.C = out; a.ldc = ldc; a.coff = coff; a.act = act; a.act_eps = eps; a.res = res; a.ldr = ldr;
    a.table = table; a.table_mod = mod;
    a.Wp = L.Wp; a.NT = L.NT; a.stream_b = L.stream_b;
    if (bf16 && linear_tc_supported(a)) op_linear_tc(c, a);
    else op_linear(c, a);
    return out;
  }
  float* norm(const float* x, const float* res, const LNW& n, int64_t M, int act = ACT_NONE, float* out = nullptr) {
    if (!out) out = c.allocf((size_t)M * n.D);
    op_layernorm(c, x, res, n.g, n.b, out, M, n.D, act);
    return out;
  }
  // x [B*N, D] -> self attention block output (before residual/norm): att_out = out_proj(attn(qkv)) + res
  float* self_attention(const float* x, const MHAW& m, int64_t B, int N, const float* res, bool kv_shift = false,
                        float* avgw = nullptr) {
    const int64_t M = B * N;
    const int D = m.D, Dp = m.Dp;                      // Dp > D: heads zero-padded through the weights (pack_mha)
    float* qkv = linear(x, D, m.in_proj, M);
    float* att = c.allocf((size_t)M * Dp);
    attention(qkv, 3 * Dp, qkv + Dp, 3 * Dp, qkv + 2 * Dp, 3 * Dp, att, Dp, B, N, N, m.H, m.hd, kv_shift, avgw);
    return linear(att, Dp, m.out_proj, M, ACT_NONE, res, D);
  }
  // post-norm encoder layer (nn.TransformerEncoderLayer / SelfAttentionBlock)
  float* encoder_layer(const float* x, const TELW& t, int64_t B, int N, int act, float* avgw = nullptr) {
    const int64_t M = B * N;
    static const bool unfused = std::getenv("PAUT_LRL_UNFUSED") != nullptr;          // A/B switch
    const bool fuse = bf16 && !unfused && t.attn.D == 128 && t.n1.D == 128 && t.n2.D == 128 && t.attn.out_proj.Wrow && t.l2.Wrow &&
                      lin_res_ln_supported(t.attn.out_proj.N, t.attn.out_proj.K, t.attn.Dp) &&
                      lin_res_ln_supported(t.l2.N, t.l2.K, t.l1.N);
    if (fuse) {
      // out-projection + residual + LayerNorm and FFN layer 2 + residual + LayerNorm as one mma.sync kernel each
      const MHAW& m = t.attn;
      float* qkv = linear(x, m.D, m.in_proj, M);
      float* att = c.allocf((size_t)M * m.Dp);
      attention(qkv, 3 * m.Dp, qkv + m.Dp, 3 * m.Dp, qkv + 2 * m.Dp, 3 * m.Dp, att, m.Dp, B, N, N, m.H, m.hd, false, avgw);
      float* x1 = c.allocf((size_t)M * 128);
      op_lin_res_ln(c, att, m.Dp, m.out_proj.Wrow, m.out_proj.K, m.out_proj.b, x, t.n1.g, t.n1.b, x1, M);
      float* h = linear(x1, m.D, t.l1, M, act);
      float* out = c.allocf((size_t)M * 128);
      op_lin_res_ln(c, h, t.l1.N, t.l2.Wrow, t.l2.K, t.l2.b, x1, t.n2.g, t.n2.b, out, M);
      return out;
    }
    float* y = self_attention(x, t.attn, B, N, x, false, avgw);
    float* x1 = norm(y, nullptr, t.n1, M);
    float* h = linear(x1, t.attn.D, t.l1, M, act);
    float* y2 = linear(h, t.l1.N, t.l2, M, ACT_NONE, x1, t.attn.D);
    return norm(y2, nullptr, t.n2, M);
  }
  float* conv(const float* in, int64_t A, int Lin_, const ConvW& w, int dil, int stride, int pad, bool relu,
              const float* res, float* out, int ldc, int coff, float* pool, int ldp, int poff, int* Lout_ = nullptr) {
    ConvArgs a;
    a.in = in; a.A = A; a.Lin = Lin_; a.Cin = w.Cin; a.w = w.w; a.shift = w.shift; a.Cout = w.Cout; a.taps = w.taps;
    a.dil = dil; a.stride = stride; a.pad = pad;
    a.Lout = (Lin_ + 2 * pad - dil * (w.taps - 1) - 1) / stride + 1;
    a.relu = relu; a.res = res; a.ldr = w.Cout; a.out = out; a.ldc = ldc; a.coff = coff;
    a.pool = pool; a.ldp = ldp; a.poff = poff;
    if (Lout_) *Lout_ = a.Lout;
    op_conv(c, a);
    return out;
  }
  // ---- bf16 mode: tcgen05 conv on flat rows
  bool tc_convs(int S) const { return bf16 && S % 8 == 0 && S + CONV_HALO >= 128; }
  __nv_bfloat16* alloc_flat(int64_t A, int L, int C) {
    return static_cast<__nv_bfloat16*>(c.alloc(flat_rows(A, L, CONV_HALO) * (size_t)C * sizeof(__nv_bfloat16)));
  }
  // the stem reads the caller's dtype directly (no fp32 expansion of a bf16 volume)
  void stem_flat(const XIn& x, int64_t A, int S, const ConvW& w, __nv_bfloat16* out, int ldc, int coff) {
    op_stem_flat(c, x.p, x.dtype, A, S, w.w, w.shift, w.taps, w.Cout, true, out, ldc, coff, CONV_HALO);
  }
  void stem_flat(const float* x, int64_t A, int S, const ConvW& w, __nv_bfloat16* out, int ldc, int coff) {
    op_stem_flat(c, x, PAUT_F32, A, S, w.w, w.shift, w.taps, w.Cout, true, out, ldc, coff, CONV_HALO);
  }
  void convtc(const __nv_bfloat16* in, int64_t A, int L, const ConvW& w, int dil, bool relu, const __nv_bfloat16* res,
              int ldr, __nv_bfloat16* out, int ldc, int coff, float* pool_out, int ldp, int poff) {
    ConvTcLaunch a;
    a.in = in; a.A = A; a.L = L; a.Cin = w.Cin; a.Cout = w.Cout; a.Wp = w.Wp; a.shift = w.shift; a.taps = w.taps;
    a.dil = dil; a.pad = w.taps / 2; a.relu = relu; a.res = res; a.ldr = ldr; a.out = out; a.ldc = ldc; a.coff = coff;
    a.NT = w.NT; a.CB = w.CB;
    if (pool_out) {
      const size_t tiles = (flat_rows(A, L, CONV_HALO) + 127) / 128;
      a.pool_partial = c.allocf(tiles * 3 * (size_t)w.Cout);
      a.pool_out = pool_out; a.ldp = ldp; a.poff = poff;
    }
    PAUT_CHECK(w.Wp != nullptr, PAUT_ERR_STATE, "conv_tc: weights were not packed for the tensor-core path");
    op_conv_tc(c, a);
  }
  // Conv1d k3 stride 2 pad 1 (+BN shift, ReLU) on flat rows row(a,l) = H0 + a*Lp + l with L, Lp, H0 even: reads the
  // input as [R/2, 2*Cin] row pairs and produces flat rows with (L/2, Lp/2, H0/2).  out may be null (pool only).
  void convtc_s2(const __nv_bfloat16* in, int64_t A, int L, int Lp, int H0, const ConvW& w, __nv_bfloat16* out,
                 float* pool_out, int ldp, int poff) {
    PAUT_CHECK(w.Wp_s2 != nullptr && L % 2 == 0 && Lp % 2 == 0 && H0 % 2 == 0, PAUT_ERR_STATE,
               "conv_tc stride 2: weights not packed or odd row geometry");
    ConvTcLaunch a;
    a.in = in; a.A = A; a.L = L / 2; a.Lp = Lp / 2; a.H0 = H0 / 2; a.Cin = 2 * w.Cin; a.Cout = w.Cout; a.Wp = w.Wp_s2;
    a.NT = w.NT_s2; a.CB = w.CB_s2; a.skip_lo = true;
    a.shift = w.shift; a.taps = 2; a.dil = 1; a.pad = 1; a.relu = true; a.out = out; a.ldc = w.Cout; a.coff = 0;
    if (pool_out) {
      const size_t tiles = ((size_t)a.H0 + (size_t)A * a.Lp + 127) / 128;
      a.pool_partial = c.allocf(tiles * 3 * (size_t)w.Cout);
      a.pool_out = pool_out; a.ldp = ldp; a.poff = poff;
    }
    op_conv_tc(c, a);
  }
};
template <typename T>
T* slot_at(const paut_outputs& o, int i, int64_t elem_off) {
  return o.slot[i] ? static_cast<T*>(o.slot[i]) + elem_off : nullptr;
}
}  // namespace

// ------------------------------------------------------------------------------------------ MSC / MSC_N
void Model::fwd_msc(const void* xin, int x_dtype, int64_t B, int N, int S, const paut_outputs& out, int64_t b0) {
  Ctx& c = *ctx;
  G g{c, cfg.precision == PAUT_PRECISION_BF16};
  const int64_t A = B * N;
  const bool isn = kind == PAUT_MODEL_MSC_N;
  const Lin& l2 = lin["shared_layer.2"];
  const int D = l2.N;
  float* h = nullptr;
  if (enc_tc.ready) {
    // bf16 mode: conv1 -> conv2 -> channel mean -> MLP -> + position table in one tcgen05 kernel
    h = c.allocf((size_t)A * D);
    if (x_dtype != PAUT_BF16) {              // the fused encoder streams bf16 rows with cp.async
      void* xb = c.alloc((size_t)A * S * sizeof(__nv_bfloat16));
      op_to_bf16(c, static_cast<const float*>(xin), xb, A * S);
      xin = xb;
      x_dtype = PAUT_BF16;
    }
    op_msc_encoder_tc(c, xin, x_dtype, A, S, N, H("conv1d.0.weight").data.data(), H("conv1d.0.bias").data.data(), enc_tc.Bc, enc_tc.W1p,
                      lin["shared_layer.0"].b, enc_tc.W2p, l2.b, raw["position_encoding.encoding"], h);
  } else if (mscn.ready && isn && std::getenv("PAUT_MSCN_SIMT") == nullptr) {
    // bf16 mode, MSC_N: conv1 -> conv2 -> background subtraction -> channel mean in one tcgen05 kernel (TMA input),
    // then the two shared layers on the tcgen05 GEMM
    if (x_dtype != PAUT_BF16) {
      void* xb = c.alloc((size_t)A * S * sizeof(__nv_bfloat16));
      op_to_bf16(c, static_cast<const float*>(xin), xb, A * S);
      xin = xb;
    }
    float* f = c.allocf((size_t)A * S);
    op_mscn_front(c, xin, A, S, mscn.W, mscn.b2, mscn.f_const, f);
    float* h0 = g.linear(f, S, lin["shared_layer.0"], A, ACT_RELU);
    h = g.linear(h0, l2.K, l2, A, ACT_RELU, nullptr, 0, nullptr, 0, 0, 0.f, raw["position_encoding.encoding"], N);
  } else {
    const float* x = static_cast<const float*>(xin);
    if (x_dtype != PAUT_F32) {
      float* t = c.allocf((size_t)A * S);
      op_to_f32(c, xin, x_dtype, t, A * S);
      x = t;
    }
    float* f = c.allocf((size_t)A * S);
    op_msc_front(c, x, A, S, raw["conv1d.0.weight"], raw["conv1d.0.bias"], raw["conv1d.2.weight"],
                 raw["conv1d.2.bias"], isn ? raw["background_extractor.weight"] : nullptr,
                 isn ? raw["background_extractor.bias"] : nullptr, f);
    float* h0 = g.linear(f, S, lin["shared_layer.0"], A, ACT_RELU);
    h = g.linear(h0, l2.K, l2, A, ACT_RELU, nullptr, 0, nullptr, 0, 0, 0.f, raw["position_encoding.encoding"], N);
  }
  if (set_tc.ready && msc_set_tc_supported(N, D, cfg.num_heads, cfg.hidden_sizes[2])) {
    // bf16 mode: the whole per-set stage in three fused mma.sync kernels
    const std::string te = "transformer_encoder.";
    const MHAW& sa = mha["self"];
    float* x1 = c.allocf((size_t)A * D);
    // A/B switch between the two attention blocks: PAUT_ATTN=tc | mma (default below)
    static const char* attn_env = std::getenv("PAUT_ATTN");
    static const bool attn_want_tc = attn_env ? std::string(attn_env) == "tc" : false;
    const bool attn_tc = attn_want_tc && msc_attn_tc_supported(N, D, cfg.num_heads) && set_tc.tc_w[0][0] != nullptr;
    if (attn_tc)
      op_msc_attn_tc(c, h, set_tc.tc_w[0][0], set_tc.tc_w[0][1], set_tc.tc_w[0][2], sa.in_proj.b, sa.out_proj.b,
                     ln[te + "norm1"].g, ln[te + "norm1"].b, x1, B, N, false);
    else
      op_msc_attn_block(c, h, set_tc.Wqkv_self, sa.in_proj.b, set_tc.Wo_self, sa.out_proj.b, ln[te + "norm1"].g,
                        ln[te + "norm1"].b, x1, B, N, false);
    const float* pre = nullptr;
    const float* pre_g = nullptr;
    const float* pre_b = nullptr;
    float* x2 = x1;
    if (!isn) {
      const MHAW& ca = mha["cross"];
      x2 = c.allocf((size_t)A * D);
      if (attn_tc)
        op_msc_attn_tc(c, x1, set_tc.tc_w[1][0], set_tc.tc_w[1][1], set_tc.tc_w[1][2], ca.in_proj.b, ca.out_proj.b,
                       ln[te + "norm2"].g, ln[te + "norm2"].b, x2, B, N, true);  // NN_models.py:35-37
      else if (msc_attn_tail_supported(c, N)) {
        // the FFN / LayerNorm / classifier tail rides in the epilogue of the second attention block: no [B, N, 64] round trip
        MscTail tl{set_tc.W1, lin[te + "ffn.0"].b, set_tc.W2, lin[te + "ffn.2"].b, ln[te + "norm3"].g, ln[te + "norm3"].b, set_tc.Wc,
                   lin["classifier"].b, slot_at<float>(out, 0, b0 * N), slot_at<float>(out, 1, b0 * N), slot_at<float>(out, 2, b0 * N)};
        op_msc_attn_block(c, x1, set_tc.Wqkv_cross, ca.in_proj.b, set_tc.Wo_cross, ca.out_proj.b, ln[te + "norm2"].g,
                          ln[te + "norm2"].b, x2, B, N, true, &tl);
        return;
      } else
        op_msc_attn_block(c, x1, set_tc.Wqkv_cross, ca.in_proj.b, set_tc.Wo_cross, ca.out_proj.b, ln[te + "norm2"].g,
                          ln[te + "norm2"].b, x2, B, N, true);
    } else {
      float* loc = c.allocf((size_t)A * D);
      op_dwconv_seq(c, x1, raw[te + "local_attn.local_conv.weight"], raw[te + "local_attn.local_conv.bias"], loc, B, N, D, 5);
      pre = loc; pre_g = ln[te + "norm2"].g; pre_b = ln[te + "norm2"].b;        // x2 = LN2(x1 + local(x1))
    }
    op_msc_ffn_head(c, x2, pre, pre_g, pre_b, set_tc.W1, lin[te + "ffn.0"].b, set_tc.W2, lin[te + "ffn.2"].b,
                    ln[te + "norm3"].g, ln[te + "norm3"].b, set_tc.Wc, lin["classifier"].b,
                    slot_at<float>(out, 0, b0 * N), slot_at<float>(out, 1, b0 * N), slot_at<float>(out, 2, b0 * N), A);
    return;
  }
  float* y = g.self_attention(h, mha["self"], B, N, h);
  h = g.norm(y, nullptr, ln["transformer_encoder.norm1"], A);
  if (!isn) {
    y = g.self_attention(h, mha["cross"], B, N, h, /*kv_shift=*/true);        // NN_models.py:35-36
    h = g.norm(y, nullptr, ln["transformer_encoder.norm2"], A);
  } else {
    float* loc = c.allocf((size_t)A * D);
    op_dwconv_seq(c, h, raw["transformer_encoder.local_attn.local_conv.weight"],
                  raw["transformer_encoder.local_attn.local_conv.bias"], loc, B, N, D, 5);
    h = g.norm(h, loc, ln["transformer_encoder.norm2"], A);
  }
  float* f1 = g.linear(h, D, lin["transformer_encoder.ffn.0"], A, ACT_RELU);
  y = g.linear(f1, lin["transformer_encoder.ffn.0"].N, lin["transformer_encoder.ffn.2"], A, ACT_NONE, h, D);
  h = g.norm(y, nullptr, ln["transformer_encoder.norm3"], A);
  float* o = g.linear(h, D, lin["classifier"], A);
  op_msc_head(c, o, A, slot_at<float>(out, 0, b0 * N), slot_at<float>(out, 1, b0 * N), slot_at<float>(out, 2, b0 * N));
}

// ------------------------------------------------------------------------------------------ MSC Conv1D
void Model::fwd_conv1d_msc(const void* x, int x_dtype, int64_t B, int N, int S, const paut_outputs& out, int64_t b0) {
  Ctx& c = *ctx;
  G g{c, cfg.precision == PAUT_PRECISION_BF16};
  const int64_t A = B * N;
  float* xt = c.allocf((size_t)A * S);
  op_transpose_sn(c, x, x_dtype, xt, B, S, N);                                  // MSC_Conv1D_training.py:81
  const ConvW& c0 = conv["feature_extractor.0"];
  float* feat = c.allocf((size_t)A * 128);
  if (g.tc_convs(S)) {
    __nv_bfloat16* a0 = g.alloc_flat(A, S, 64);
    g.stem_flat(xt, A, S, c0, a0, 64, 0);
    __nv_bfloat16* a1 = g.alloc_flat(A, S, 128);
    g.convtc(a0, A, S, conv["feature_extractor.2"], 1, true, nullptr, 0, a1, 128, 0, nullptr, 0, 0);
    g.convtc(a1, A, S, conv["feature_extractor.4"], 1, true, nullptr, 0, nullptr, 0, 0, feat, 128, 0);
  } else {
    float* a0 = c.allocf((size_t)A * S * 64);
    op_stem_conv(c, xt, A, S, c0.w, c0.shift, c0.taps, c0.Cout, true, a0);
    float* a1 = c.allocf((size_t)A * S * 128);
    g.conv(a0, A, S, conv["feature_extractor.2"], 1, 1, 1, true, nullptr, a1, 128, 0, nullptr, 0, 0);
    g.conv(a1, A, S, conv["feature_extractor.4"], 1, 1, 0, true, nullptr, nullptr, 0, 0, feat, 128, 0);   // mean over L
  }
  float* h = feat;
  for (const TELW& t : tel) h = g.encoder_layer(h, t, B, N, ACT_RELU);
  float* h1 = g.linear(h, 128, lin["classifier.0"], A, ACT_RELU);
  if (out.slot[0]) g.linear(h1, 64, lin["classifier.2"], A, ACT_SIGMOID, nullptr, 0, slot_at<float>(out, 0, b0 * N), 1, 0);
}

// ------------------------------------------------------------------------------------------ SignalSequenceDetector
void Model::fwd_ssd(XIn& x, int64_t B, int N, int S, const paut_outputs& out, int64_t b0, int64_t) {
  Ctx& c = *ctx;
  G g{c, cfg.precision == PAUT_PRECISION_BF16};
  const int64_t A = B * N;
  const int d = cfg.d_model, C = cfg.num_classes;
  const ConvW& c1 = conv["signal_encoder.conv1"];
  float* feat = c.allocf((size_t)A * 256);
  if (g.tc_convs(S)) {
    __nv_bfloat16* a0 = g.alloc_flat(A, S, 64);
    g.stem_flat(x, A, S, c1, a0, 64, 0);
    __nv_bfloat16* a1 = g.alloc_flat(A, S, 128);
    g.convtc(a0, A, S, conv["signal_encoder.conv2"], 1, true, nullptr, 0, a1, 128, 0, nullptr, 0, 0);
    g.convtc(a1, A, S, conv["signal_encoder.conv3"], 1, true, nullptr, 0, nullptr, 0, 0, feat, 256, 0);
  } else {
    float* a0 = c.allocf((size_t)A * S * 64);
    op_stem_conv(c, x.as_f32(c), A, S, c1.w, c1.shift, c1.taps, c1.Cout, true, a0);
    float* a1 = c.allocf((size_t)A * S * 128);
    g.conv(a0, A, S, conv["signal_encoder.conv2"], 1, 1, 2, true, nullptr, a1, 128, 0, nullptr, 0, 0);
    g.conv(a1, A, S, conv["signal_encoder.conv3"], 1, 1, 1, true, nullptr, nullptr, 0, 0, feat, 256, 0);
  }
  float* seq = g.linear(feat, 256, lin["signal_encoder.fc"], A, ACT_NONE, nullptr, 0, nullptr, 0, 0, 0.f,
                        raw["sequence_transformer.pos_encoder.pe"], N);          // + pe[:, :N]
  for (const TELW& t : tel) seq = g.encoder_layer(seq, t, B, N, ACT_RELU);
  // ContextAggregator: 2-layer bidirectional GRU + projection (model.py:179-192)
  const float* rin = seq;
  int rin_ld = d;
  float* ro = nullptr;
  for (const RNNW& r : rnn) {
    float* gi = g.linear(rin, rin_ld, r.ih, A);
    ro = c.allocf((size_t)A * 2 * r.H);
    op_rnn_bidir(c, gi, r.whh_t, r.bhh, ro, B, N, r.H, r.G);
    rin = ro;
    rin_ld = 2 * r.H;
  }
  float* ctxf = g.linear(ro, d, lin["context_aggregator.projection"], A);
  // health features go straight into the right half of the concat buffer
  float* comb = c.allocf((size_t)A * 2 * d);
  float* hf = g.linear(seq, d, lin["health_extractor.0"], A, ACT_RELU);
  hf = g.linear(hf, d / 2, lin["health_extractor.2"], A, ACT_RELU);
  g.linear(hf, d / 4, lin["health_extractor.4"], A, ACT_NONE, nullptr, 0, comb, 2 * d, d);
  // attention over the sequence (model.py:313-314)
  float* at = g.linear(seq, d, lin["attention.0"], A, ACT_RELU);
  float* sc = g.linear(at, d / 4, lin["attention.2"], A);
  float* attw = slot_at<float>(out, 3, b0 * N);
  if (!attw) attw = c.allocf((size_t)A);
  op_softmax_seq(c, sc, attw, B, N);
  float* enh = c.allocf((size_t)A * d);
  op_rowscale_add(c, seq, attw, ctxf, enh, A, d);                               // model.py:317
  op_copy_cols(c, enh, d, comb, 2 * d, 0, A, d);
  float* an = g.linear(comb, 2 * d, lin["anomaly_detector.anomaly_net.0"], A, ACT_RELU);
  an = g.linear(an, 64, lin["anomaly_detector.anomaly_net.3"], A, ACT_RELU);
  float* anomaly = slot_at<float>(out, 2, b0 * N);
  if (!anomaly) anomaly = c.allocf((size_t)A);
  g.linear(an, 32, lin["anomaly_detector.anomaly_net.5"], A, ACT_SIGMOID, nullptr, 0, anomaly, 1, 0);
  if (out.slot[0]) {
    float* ch = g.linear(enh, d, lin["detection_head.class_head.0"], A, ACT_RELU);
    float* logits = slot_at<float>(out, 0, b0 * N * C);
    g.linear(ch, d / 2, lin["detection_head.class_head.3"], A, ACT_NONE, nullptr, 0, logits, C, 0);
    op_add_anomaly(c, logits, anomaly, A, C);                                   // model.py:327-334
  }
  if (out.slot[1]) {
    float* ph = g.linear(enh, d, lin["detection_head.position_head.0"], A, ACT_RELU);
    g.linear(ph, d / 2, lin["detection_head.position_head.3"], A, ACT_SIGMOID, nullptr, 0,
             slot_at<float>(out, 1, b0 * N * 2), 2, 0);
  }
}

// ------------------------------------------------------------------------------------------ TwoStageDefectDetector
void Model::fwd_two_stage(XIn& x, int64_t B, int N, int S, const paut_outputs& out, int64_t b0) {
  Ctx& c = *ctx;
  G g{c, cfg.precision == PAUT_PRECISION_BF16};
  const int64_t A = B * N;
  const int d = cfg.d_model, q = d / 4;
  float* feat = c.allocf((size_t)A * d);
  const char* names[4] = {"small", "medium", "large", "xlarge"};
  const bool tcc = g.tc_convs(S) && q % 16 == 0;
  static const bool no_fused = std::getenv("PAUT_TS_UNFUSED") != nullptr;     // A/B switch for the per-layer schedule
  if (g.bf16 && ts_enc.ready && ts_encoder_supported(S, d) && !no_fused) {
    // one persistent kernel: TMA input staging, stems on HFMA2, the four second convolutions on tcgen05, BN shift +
    // ReLU + mean over the signal length in the epilogue (two_stage_model.py:102-118); no activation touches HBM
    op_ts_encoder(c, x.as_bf16(c), A, S, ts_enc.Wst, ts_enc.W2, ts_enc.shift2, feat);
  } else if (tcc && ts_grouped.ready) {
    // all four stems write one [rows, 4q] buffer; the four second convolutions + BN + ReLU + mean run as one
    // grouped tcgen05 launch whose pooled output is the concatenated feature vector (two_stage_model.py:102-118)
    __nv_bfloat16* a0h = g.alloc_flat(A, S, d);
    for (int i = 0; i < 4; ++i)
      g.stem_flat(x, A, S, conv[std::string("signal_encoder.conv_") + names[i] + ".0"], a0h, d, q * i);
    ConvTcLaunch a;
    a.in = a0h; a.A = A; a.L = S; a.Cin = d; a.Cout = d; a.Wp = ts_grouped.Wp; a.shift = ts_grouped.shift;
    a.NT = d; a.CB = q; a.groups = 4;
    int tmax = 0;
    for (int i = 0; i < 4; ++i) { a.gtaps[i] = ts_grouped.taps[i]; a.goff[i] = ts_grouped.goff[i]; tmax = std::max(tmax, a.gtaps[i]); }
    a.taps = tmax; a.dil = 1; a.pad = tmax / 2; a.relu = true;
    const size_t tiles = (flat_rows(A, S, CONV_HALO) + 127) / 128;
    a.pool_partial = c.allocf(tiles * 3 * (size_t)d);
    a.pool_out = feat; a.ldp = d; a.poff = 0;
    op_conv_tc(c, a);
  } else {
  float* a0 = tcc ? nullptr : c.allocf((size_t)A * S * q);
  __nv_bfloat16* a0h = tcc ? g.alloc_flat(A, S, q) : nullptr;
  for (int i = 0; i < 4; ++i) {                                                  // two_stage_model.py:102-114
    const std::string p = std::string("signal_encoder.conv_") + names[i] + ".";
    const ConvW& s = conv[p + "0"];
    const ConvW& w = conv[p + "3"];
    if (tcc) {
      g.stem_flat(x, A, S, s, a0h, q, 0);
      g.convtc(a0h, A, S, w, 1, true, nullptr, 0, nullptr, 0, 0, feat, d, q * i);
    } else {
      op_stem_conv(c, x.as_f32(c), A, S, s.w, s.shift, s.taps, s.Cout, true, a0);
      g.conv(a0, A, S, w, 1, 1, w.taps / 2, true, nullptr, nullptr, 0, 0, feat, d, q * i);
    }
  }
  }
  float* seq = nullptr;
  const Lin& pj = lin["signal_encoder.projection.0"];
  static const bool lrl_unfused = std::getenv("PAUT_LRL_UNFUSED") != nullptr;
  if (g.bf16 && !lrl_unfused && pj.Wrow && d == 128 && lin_res_ln_supported(pj.N, pj.K, d)) {
    // projection Linear + LayerNorm + positional encoding (two_stage_model.py:86-89,133) in one kernel
    seq = c.allocf((size_t)A * d);
    const LNW& pn = ln["signal_encoder.projection.1"];
    op_lin_res_ln(c, feat, d, pj.Wrow, pj.K, pj.b, nullptr, pn.g, pn.b, seq, A, raw["sequence_transformer.pos_encoder.pe"], N);
  } else {
    float* pr = g.linear(feat, d, pj, A);
    seq = g.norm(pr, nullptr, ln["signal_encoder.projection.1"], A);
    // + pe[:, :N]: reuse the row-table epilogue through an identity-free path
    float* seq2 = c.allocf((size_t)A * d);
    op_add_table(c, seq, raw["sequence_transformer.pos_encoder.pe"], N, seq2, A, d);
    seq = seq2;
  }
  for (const TELW& t : tel) seq = g.encoder_layer(seq, t, B, N, ACT_RELU);
  static const bool heads_unfused = std::getenv("PAUT_TS_HEADS_UNFUSED") != nullptr;     // A/B switch
  if (g.bf16 && ts_heads.ready && !heads_unfused) {
    // final LayerNorm + the four heads in one kernel (two_stage_model.py:160-251): the rows are read once
    const char* mods[4] = {"defect_classifier.classifier", "defect_classifier.uncertainty",
                           "position_predictor.position_predictor", "position_predictor.uncertainty"};
    const int acts[4] = {ACT_NONE, ACT_SOFTPLUS, ACT_SIGMOID, ACT_SOFTPLUS};
    const float epss[4] = {0.f, 1e-6f, 0.f, 1e-6f};
    float* logits = slot_at<float>(out, 0, b0 * N * 2);
    if (!logits) logits = c.allocf((size_t)A * 2);
    float* outs[4] = {logits, slot_at<float>(out, 2, b0 * N * 2), slot_at<float>(out, 3, b0 * N * 2), slot_at<float>(out, 4, b0 * N * 2)};
    const float *b0s[4], *lgs[4], *lbs[4], *w4s[4], *b4s[4];
    for (int h = 0; h < 4; ++h) {
      const std::string m = mods[h];
      b0s[h] = lin[m + ".0"].b; lgs[h] = ln[m + ".1"].g; lbs[h] = ln[m + ".1"].b; w4s[h] = lin[m + ".4"].W; b4s[h] = lin[m + ".4"].b;
    }
    const LNW& fn = ln["sequence_transformer.norm"];
    op_ts_heads(c, seq, fn.g, fn.b, ts_heads.W0, b0s, lgs, lbs, w4s, b4s, acts, epss, outs, A);
    op_two_stage_final(c, logits, slot_at<float>(out, 1, b0 * N * 2), outs[2], A);
    return;
  }
  seq = g.norm(seq, nullptr, ln["sequence_transformer.norm"], A);
  auto head = [&](const std::string& m, int act, float eps, float* dst) {
    float* h = g.linear(seq, d, lin[m + ".0"], A);
    h = g.norm(h, nullptr, ln[m + ".1"], A, ACT_RELU);
    if (!dst) dst = c.allocf((size_t)A * 2);
    g.linear(h, 64, lin[m + ".4"], A, act, nullptr, 0, dst, 2, 0, eps);
    return dst;
  };
  float* logits = head("defect_classifier.classifier", ACT_NONE, 0.f, slot_at<float>(out, 0, b0 * N * 2));
  if (out.slot[2]) head("defect_classifier.uncertainty", ACT_SOFTPLUS, 1e-6f, slot_at<float>(out, 2, b0 * N * 2));
  float* pos = nullptr;
  if (out.slot[3]) pos = head("position_predictor.position_predictor", ACT_SIGMOID, 0.f, slot_at<float>(out, 3, b0 * N * 2));
  if (out.slot[4]) head("position_predictor.uncertainty", ACT_SOFTPLUS, 1e-6f, slot_at<float>(out, 4, b0 * N * 2));
  op_two_stage_final(c, logits, slot_at<float>(out, 1, b0 * N * 2), pos, A);
}

// ------------------------------------------------------------------------------------------ EnhancedSignalSequenceDetector
void Model::fwd_enhanced(XIn& x, int64_t B, int N, int S, const paut_outputs& out, int64_t b0, int64_t Btot) {
  Ctx& c = *ctx;
  G g{c, cfg.precision == PAUT_PRECISION_BF16};
  const int64_t A = B * N;
  const int d = cfg.d_model, C = cfg.num_classes;
  const std::string e = "signal_encoder.";
  // ---- EnhancedSignalEncoder (enhanced_model.py:135-175); three rotating [A,S,128] buffers
  const ConvW& ci = conv[e + "conv_init.0"];
  float* feat = c.allocf((size_t)A * 640);
  int L1 = 0, L2 = 0;
  if (g.tc_convs(S)) {
    // bf16 mode: every stride-1 conv of the encoder is a tcgen05 implicit GEMM on flat bf16 rows.
    // The conv stack walks the resident chunk in SUB-CHUNKS of 16 384 A-scans with its own, smaller set of flat-row
    // buffers (378 KB per A-scan): the buffers no longer dictate the size of the resident chunk, which grows 4x for
    // everything behind the encoder (recurrences 52 -> 27 ms, linears 88 -> 70 ms per 1 M A-scans: fuller grids).
    // Sub-chunks start on multiples of the layout period (64 A-scans), so the pooled sums are the same tiles in the same
    // order as in one launch over the whole chunk.  L2-sized sub-chunks (256-1024 A-scans: a 128-channel layer moves 96 KB
    // per 128-row tile against 12.6 MFLOP and is HBM-bound at 2.9x its tensor time when it streams from HBM) were the
    // idea and are SLOWER, 1164 / 950 / 869 ms at 256 / 512 / 1024 against 816: every k_conv_tc launch stages its resident
    // weights and sets up its rings again, ~10 us per launch (profiles/r02/3c_enhanced_subchunk_sweep.log); on-chip reuse
    // needs the layers fused into one persistent kernel, not smaller launches.
    static const int64_t sub_env = [] { const char* e2 = std::getenv("PAUT_CONV_SUBCHUNK"); return e2 ? (int64_t)atoll(e2) : (int64_t)16384; }();
    const int64_t Asub = sub_env > 0 ? std::min<int64_t>(A, (sub_env + 63) / 64 * 64) : A;
    __nv_bfloat16* s0 = g.alloc_flat(Asub, S, 64);
    __nv_bfloat16* bufA = g.alloc_flat(Asub, S, 128);
    __nv_bfloat16* bufB = g.alloc_flat(Asub, S, 128);
    __nv_bfloat16* bufC = g.alloc_flat(Asub, S, 128);
    const int Lp0 = S + CONV_HALO;
    const ConvW& p1 = conv[e + "pyramid_1"];
    const ConvW& p2 = conv[e + "pyramid_2"];
    const bool tc_pyramid = p1.Wp_s2 && p2.Wp_s2 && S % 4 == 0 && Lp0 % 4 == 0 && Lp0 / 4 >= 64;
    __nv_bfloat16* x1 = nullptr;
    float* hd = nullptr;
    float* x1f = nullptr;
    if (tc_pyramid) {
      x1 = static_cast<__nv_bfloat16*>(c.alloc(((size_t)CONV_HALO / 2 + (size_t)Asub * (Lp0 / 2)) * 256 * sizeof(__nv_bfloat16)));
    } else {
      hd = c.allocf((size_t)Asub * S * 128);
      x1f = c.allocf((size_t)Asub * ((S + 1) / 2) * 256);
    }
    const size_t esz = x.dtype == PAUT_BF16 ? 2 : 4;
    const size_t ws_mark = c.ws_off;                       // per-sub-chunk scratch (pooled partial sums) is released every pass
    for (int64_t a0 = 0; a0 < A; a0 += Asub) {
      const int64_t na = std::min<int64_t>(Asub, A - a0);
      c.ws_off = ws_mark;
      XIn xs;
      xs.p = static_cast<const char*>(x.p) + (size_t)a0 * S * esz; xs.dtype = x.dtype; xs.n = na * S;
      float* feat_s = feat + (size_t)a0 * 640;
      g.stem_flat(xs, na, S, ci, s0, 64, 0);
      if (enh_branches.ready) {
        // the four dilated branches read the same 64 channels: one launch, one staged window (dilation 8 wide),
        // one 128-column accumulator = the concatenated output (enhanced_model.py:82-89)
        ConvTcLaunch a;
        a.in = s0; a.A = na; a.L = S; a.Cin = 64; a.Cout = 128; a.Wp = enh_branches.Wp; a.shift = enh_branches.shift;
        a.NT = 128; a.CB = 64; a.groups = 4; a.shared_input = true;
        for (int b = 0; b < 4; ++b) { a.gtaps[b] = 3; a.goff[b] = enh_branches.goff[b]; a.gdil[b] = 1 << b; }
        a.taps = 3; a.dil = 8; a.pad = 1; a.relu = false; a.out = bufA; a.ldc = 128; a.coff = 0;
        op_conv_tc(c, a);
      } else {
        for (int b = 0; b < 4; ++b)
          g.convtc(s0, na, S, conv[e + "multi_scale.branch" + istr(b + 1)], 1 << b, false, nullptr, 0, bufA, 128, 32 * b,
                   nullptr, 0, 0);
      }
      g.convtc(bufA, na, S, conv[e + "multi_scale.combine.0"], 1, true, nullptr, 0, bufB, 128, 0, nullptr, 0, 0);
      __nv_bfloat16* h = bufB;
      __nv_bfloat16* spare = bufC;
      for (int r = 0; r < 3; ++r) {
        const int dil = 1 << r;
        const std::string qn = e + "res_blocks." + istr(r) + ".conv_block.";
        g.convtc(h, na, S, conv[qn + "0"], dil, true, nullptr, 0, bufA, 128, 0, nullptr, 0, 0);
        g.convtc(bufA, na, S, conv[qn + "3"], dil, true, h, 128, spare, 128, 0, r == 2 ? feat_s : nullptr, 640, 0);
        std::swap(h, spare);
      }
      if (tc_pyramid) {
        // stride-2 pyramid on the tensor cores through the space-to-depth view of the flat rows
        L1 = S / 2; L2 = S / 4;
        g.convtc_s2(h, na, S, Lp0, CONV_HALO, p1, x1, feat_s, 640, 128);
        g.convtc_s2(x1, na, L1, Lp0 / 2, CONV_HALO / 2, p2, nullptr, feat_s, 640, 384);
      } else {
        // short signals: the two stride-2 pyramid convs run on the fp32 CUDA-core kernel
        op_unflatten(c, h, na, S, CONV_HALO, 128, hd);
        g.conv(hd, na, S, p1, 1, 2, 1, true, nullptr, x1f, 256, 0, feat_s, 640, 128, &L1);
        g.conv(x1f, na, L1, p2, 1, 2, 1, true, nullptr, nullptr, 0, 0, feat_s, 640, 384, &L2);
      }
    }
  } else {
    float* s0 = c.allocf((size_t)A * S * 64);
    op_stem_conv(c, x.as_f32(c), A, S, ci.w, ci.shift, ci.taps, ci.Cout, true, s0);
    float* bufA = c.allocf((size_t)A * S * 128);
    float* bufB = c.allocf((size_t)A * S * 128);
    float* bufC = c.allocf((size_t)A * S * 128);
    for (int b = 0; b < 4; ++b)
      g.conv(s0, A, S, conv[e + "multi_scale.branch" + istr(b + 1)], 1 << b, 1, 1 << b, false, nullptr, bufA, 128, 32 * b,
             nullptr, 0, 0);
    g.conv(bufA, A, S, conv[e + "multi_scale.combine.0"], 1, 1, 0, true, nullptr, bufB, 128, 0, nullptr, 0, 0);
    float* h = bufB;
    float* spare = bufC;
    for (int r = 0; r < 3; ++r) {
      const int dil = 1 << r;
      const std::string qn = e + "res_blocks." + istr(r) + ".conv_block.";
      g.conv(h, A, S, conv[qn + "0"], dil, 1, dil, true, nullptr, bufA, 128, 0, nullptr, 0, 0);
      g.conv(bufA, A, S, conv[qn + "3"], dil, 1, dil, true, h, spare, 128, 0, r == 2 ? feat : nullptr, 640, 0);
      std::swap(h, spare);
    }
    g.conv(h, A, S, conv[e + "pyramid_1"], 1, 2, 1, true, nullptr, bufA, 256, 0, feat, 640, 128, &L1);
    g.conv(bufA, A, L1, conv[e + "pyramid_2"], 1, 2, 1, true, nullptr, nullptr, 0, 0, feat, 640, 384, &L2);
  }
  float* fc = g.linear(feat, 640, lin[e + "fc.0"], A);
  float* sf = g.norm(fc, nullptr, ln[e + "fc.1"], A, ACT_RELU);
  // ---- EnhancedSequenceTransformer (:230-251)
  float* seq = c.allocf((size_t)A * d);
  op_add_table(c, sf, raw["sequence_transformer.pos_encoder.pe"], N, seq, A, d);
  for (size_t l = 0; l < tel.size(); ++l) {
    float* aw = out.slot[6] ? static_cast<float*>(out.slot[6]) + ((int64_t)l * Btot + b0) * N * N : nullptr;
    seq = g.encoder_layer(seq, tel[l], B, N, ACT_GELU, aw);
  }
  seq = g.norm(seq, nullptr, ln["sequence_transformer.norm"], A);
  // ---- EnhancedContextAggregator (:283-313)
  const float* rin = seq;
  int rin_ld = d;
  float* lo = nullptr;
  for (const RNNW& r : rnn) {
    float* gi = g.linear(rin, rin_ld, r.ih, A);
    lo = c.allocf((size_t)A * 2 * r.H);
    op_rnn_bidir(c, gi, r.whh_t, r.bhh, lo, B, N, r.H, r.G);
    rin = lo;
    rin_ld = 2 * r.H;
  }
  float* keys = g.linear(lo, d, lin["context_aggregator.attention_keys"], A);
  float* sc = c.allocf((size_t)A);
  op_rowdot(c, keys, raw["context_aggregator.attention_query"], sc, A, d);
  float* cw = slot_at<float>(out, 7, b0 * N);
  if (!cw) cw = c.allocf((size_t)A);
  op_softmax_seq(c, sc, cw, B, N);
  float* cat = c.allocf((size_t)A * 2 * d);
  float* vals = g.linear(lo, d, lin["context_aggregator.attention_values"], A);
  op_rowscale_add(c, vals, cw, nullptr, vals, A, d);
  op_copy_cols(c, lo, d, cat, 2 * d, 0, A, d);
  op_copy_cols(c, vals, d, cat, 2 * d, d, A, d);
  float* cp = g.linear(cat, 2 * d, lin["context_aggregator.projection.0"], A);
  float* ctxf = g.norm(cp, nullptr, ln["context_aggregator.projection.1"], A);
  // ---- cross attention (:525-528): Q = encoder features, K = V = context; always 8 heads (:491)
  float* qb = g.linear(sf, d, lin["cross.q"], A);
  float* kv = g.linear(ctxf, d, lin["cross.kv"], A);
  float* att = c.allocf((size_t)A * d);
  g.attention(qb, d, kv, 2 * d, kv + d, 2 * d, att, d, B, N, N, 8, d / 8, false,
               slot_at<float>(out, 8, b0 * N * N));
  float* co = g.linear(att, d, lin["cross_attention.out_proj"], A, ACT_NONE, sf, d);
  float* cf = g.norm(co, nullptr, ln["cross_norm"], A);
  // ---- sequence integration (:531-533)
  float* cat2 = c.allocf((size_t)A * 2 * d);
  op_copy_cols(c, cf, d, cat2, 2 * d, 0, A, d);
  op_copy_cols(c, seq, d, cat2, 2 * d, d, A, d);
  float* ig = g.linear(cat2, 2 * d, lin["sequence_integration.0"], A);
  float* integ = g.norm(ig, nullptr, ln["sequence_integration.1"], A, ACT_GELU);
  // ---- EnhancedAnomalyDetector (:358-380)
  const std::string a = "anomaly_detector.";
  float* comb = c.allocf((size_t)A * 2 * d);
  op_copy_cols(c, integ, d, comb, 2 * d, 0, A, d);
  float* hl = g.norm(g.linear(seq, d, lin[a + "health_extractor.0"], A), nullptr, ln[a + "health_extractor.1"], A, ACT_GELU);
  hl = g.norm(g.linear(hl, 128, lin[a + "health_extractor.4"], A), nullptr, ln[a + "health_extractor.5"], A, ACT_GELU);
  g.linear(hl, 64, lin[a + "health_extractor.7"], A, ACT_NONE, nullptr, 0, comb, 2 * d, d);
  float* an = g.norm(g.linear(comb, 2 * d, lin[a + "anomaly_net.0"], A), nullptr, ln[a + "anomaly_net.1"], A, ACT_GELU);
  an = g.norm(g.linear(an, 128, lin[a + "anomaly_net.4"], A), nullptr, ln[a + "anomaly_net.5"], A, ACT_GELU);
  float* anomaly = slot_at<float>(out, 4, b0 * N);
  if (!anomaly) anomaly = c.allocf((size_t)A);
  g.linear(an, 64, lin[a + "anomaly_net.7"], A, ACT_SIGMOID, nullptr, 0, anomaly, 1, 0);
  if (out.slot[5]) {
    float* un = g.norm(g.linear(comb, 2 * d, lin[a + "uncertainty_net.0"], A), nullptr, ln[a + "uncertainty_net.1"], A, ACT_GELU);
    g.linear(un, 128, lin[a + "uncertainty_net.4"], A, ACT_SOFTPLUS, nullptr, 0, slot_at<float>(out, 5, b0 * N), 1, 0);
  }
  // ---- EnhancedDefectDetectionHead (:433-448)
  const std::string hd = "detection_head.";
  auto deep = [&](const std::string& n, int nout, int act, float* dst) {
    float* t = g.norm(g.linear(integ, d, lin[hd + n + ".0"], A), nullptr, ln[hd + n + ".1"], A, ACT_GELU);
    t = g.norm(g.linear(t, d / 2, lin[hd + n + ".4"], A), nullptr, ln[hd + n + ".5"], A, ACT_GELU);
    g.linear(t, d / 4, lin[hd + n + ".7"], A, act, nullptr, 0, dst, nout, 0);
  };
  auto shallow = [&](const std::string& n, int nout, float* dst) {
    float* t = g.norm(g.linear(integ, d, lin[hd + n + ".0"], A), nullptr, ln[hd + n + ".1"], A, ACT_GELU);
    g.linear(t, d / 4, lin[hd + n + ".3"], A, ACT_SOFTPLUS, nullptr, 0, dst, nout, 0);
  };
  if (out.slot[0]) {
    float* logits = slot_at<float>(out, 0, b0 * N * C);
    deep("class_head", C, ACT_NONE, logits);
    op_add_anomaly(c, logits, anomaly, A, C);                                   // :545-552
  }
  if (out.slot[1]) shallow("class_uncertainty", C, slot_at<float>(out, 1, b0 * N * C));
  if (out.slot[2]) deep("position_head", 2, ACT_SIGMOID, slot_at<float>(out, 2, b0 * N * 2));
  if (out.slot[3]) shallow("position_uncertainty", 2, slot_at<float>(out, 3, b0 * N * 2));
}

// ------------------------------------------------------------------------------------------ section 8 "next" rows
// f3: the legacy no-conv MultiSignalClassifier (resaveModelOnnx.py:24-33): MLP per A-scan, one self-attention over
// the set WITHOUT residual or norm, MLP head + sigmoid.
void Model::fwd_msc_legacy(XIn& xin, int64_t B, int N, int S, const paut_outputs& out, int64_t b0) {
  Ctx& c = *ctx;
  G g{c, cfg.precision == PAUT_PRECISION_BF16};
  const int64_t A = B * N;
  const Lin& l0 = lin["shared_layer.0"];
  const Lin& l2 = lin["shared_layer.2"];
  const float* x = xin.as_f32(c);
  float* h0 = g.linear(x, S, l0, A, ACT_RELU);
  float* h = g.linear(h0, l0.N, l2, A, ACT_RELU);
  float* att = g.self_attention(h, mha["attention"], B, N, nullptr);
  float* h1 = g.linear(att, l2.N, lin["classifier.0"], A, ACT_RELU);
  if (out.slot[0])
    g.linear(h1, lin["classifier.0"].N, lin["classifier.2"], A, ACT_SIGMOID, nullptr, 0, slot_at<float>(out, 0, b0 * N), 1, 0);
}

namespace {
// Three-layer conv stack (1 -> C0 -> C1 -> C2, BN folded, ReLU) followed by channel mean + resampling to 128 values
// per A-scan (hybrid_binary.py:139-145, complex_detection_model.py:68-75).  bf16 mode: stem into flat rows and the two
// channel-mixing convolutions on the tcgen05 implicit-GEMM kernel when their weights were packed for it.
void conv_stack_pooled(Model& m, G& g, XIn& x, int64_t A, int S, const char* n0, const char* n1, const char* n2,
                       int mode, int pk, float* pooled) {
  Ctx& c = g.c;
  const ConvW& c0 = m.conv[n0];
  const ConvW& c1 = m.conv[n1];
  const ConvW& c2 = m.conv[n2];
  if (g.tc_convs(S) && c1.Wp && c2.Wp && c1.taps / 2 <= CONV_HALO && c2.taps / 2 <= CONV_HALO) {
    __nv_bfloat16* a0 = g.alloc_flat(A, S, c0.Cout);
    g.stem_flat(x, A, S, c0, a0, c0.Cout, 0);
    __nv_bfloat16* a1 = g.alloc_flat(A, S, c1.Cout);
    g.convtc(a0, A, S, c1, 1, true, nullptr, 0, a1, c1.Cout, 0, nullptr, 0, 0);
    __nv_bfloat16* a2 = g.alloc_flat(A, S, c2.Cout);
    g.convtc(a1, A, S, c2, 1, true, nullptr, 0, a2, c2.Cout, 0, nullptr, 0, 0);
    op_chanmean_resample(c, a2, PAUT_BF16, A, S, c2.Cout, S + CONV_HALO, CONV_HALO, mode, pk, 128, pooled);
  } else {
    float* a0 = c.allocf((size_t)A * S * c0.Cout);
    op_stem_conv(c, x.as_f32(c), A, S, c0.w, c0.shift, c0.taps, c0.Cout, true, a0);
    float* a1 = c.allocf((size_t)A * S * c1.Cout);
    g.conv(a0, A, S, c1, 1, 1, c1.taps / 2, true, nullptr, a1, c1.Cout, 0, nullptr, 0, 0);
    float* a2 = c.allocf((size_t)A * S * c2.Cout);
    g.conv(a1, A, S, c2, 1, 1, c2.taps / 2, true, nullptr, a2, c2.Cout, 0, nullptr, 0, 0);
    op_chanmean_resample(c, a2, PAUT_F32, A, S, c2.Cout, S, 0, mode, pk, 128, pooled);
  }
}
}  // namespace

// f2: ImprovedMultiSignalClassifier (improved_model.py:123-157) and HybridBinaryModel (hybrid_binary.py:136-168):
// a per-A-scan conv front end, the shared MLP, a learned position table and num_layers encoder layers of the
// MSC_N type (self-attention -> LN -> depthwise local convolution(s) over the set axis -> LN -> FFN -> LN).
void Model::fwd_improved(XIn& x, int64_t B, int N, int S, const paut_outputs& out, int64_t b0) {
  Ctx& c = *ctx;
  G g{c, cfg.precision == PAUT_PRECISION_BF16};
  const int64_t A = B * N;
  const bool hyb = kind == PAUT_MODEL_HYBRID;
  const int D = cfg.hidden_sizes[1];
  float* feat = nullptr;      // input of shared_layer.0: [A, S] (improved) or [A, 256] (hybrid)
  int featK = 0;
  if (!hyb) {
    const ConvW& c0 = conv["conv1d.0"];
    const ConvW& c1 = conv["conv1d.3"];
    feat = c.allocf((size_t)A * S);
    const float* wbg = raw["background_extractor.weight"];
    const float* bbg = raw["background_extractor.bias"];
    if (g.tc_convs(S) && c1.Wp) {
      // bf16 mode: stem into flat rows, the 16 -> 32 convolution on the tcgen05 implicit-GEMM kernel
      __nv_bfloat16* a0 = g.alloc_flat(A, S, c0.Cout);
      g.stem_flat(x, A, S, c0, a0, c0.Cout, 0);
      __nv_bfloat16* a1 = g.alloc_flat(A, S, c1.Cout);
      g.convtc(a0, A, S, c1, 1, true, nullptr, 0, a1, c1.Cout, 0, nullptr, 0, 0);
      op_bgsub_chanmean(c, a1, PAUT_BF16, A, S, c1.Cout, S + CONV_HALO, CONV_HALO, 15, wbg, bbg, feat);
    } else {
      float* a0 = c.allocf((size_t)A * S * c0.Cout);
      op_stem_conv(c, x.as_f32(c), A, S, c0.w, c0.shift, c0.taps, c0.Cout, true, a0);
      float* a1 = c.allocf((size_t)A * S * c1.Cout);
      g.conv(a0, A, S, c1, 1, 1, 1, true, nullptr, a1, c1.Cout, 0, nullptr, 0, 0);
      op_bgsub_chanmean(c, a1, PAUT_F32, A, S, c1.Cout, S, 0, 15, wbg, bbg, feat);
    }
    featK = S;
  } else {
    float* pooled = c.allocf((size_t)A * 128);
    const int pk = std::max(1, S / 128);                                       // hybrid_binary.py:108-111
    conv_stack_pooled(*this, g, x, A, S, "conv_layers.0", "conv_layers.3", "conv_layers.6", /*mode=*/0, pk, pooled);
    feat = c.allocf((size_t)A * 256);
    op_seqmean_concat(c, pooled, B, N, 128, feat);
    featK = 256;
  }
  float* h0 = g.linear(feat, featK, lin["shared_layer.0"], A, ACT_RELU);
  const Lin& l3 = lin["shared_layer.3"];
  float* h = g.linear(h0, l3.K, l3, A, ACT_RELU, nullptr, 0, nullptr, 0, 0, 0.f, raw["position_encoding.encoding"], N);
  for (int i = 0; i < cfg.num_layers; ++i) {
    const std::string t = "transformer_layers." + istr(i) + ".";
    float* y = g.self_attention(h, mha[t + "self_attn"], B, N, h);
    h = g.norm(y, nullptr, ln[t + "norm1"], A);
    float* loc = c.allocf((size_t)A * D);
    op_dwconv_seq(c, h, raw[t + "local_attn.local_conv.weight"], raw[t + "local_attn.local_conv.bias"], loc, B, N, D,
                  hyb ? 11 : 9);
    if (hyb) {
      float* loc2 = c.allocf((size_t)A * D);
      op_dwconv_seq(c, loc, raw[t + "local_attn.local_conv2.weight"], raw[t + "local_attn.local_conv2.bias"], loc2, B,
                    N, D, 5);
      loc = loc2;
    }
    h = g.norm(h, loc, ln[t + "norm2"], A);
    float* f1 = g.linear(h, D, lin[t + "ffn.0"], A, ACT_RELU);
    y = g.linear(f1, lin[t + "ffn.0"].N, lin[t + "ffn.3"], A, ACT_NONE, h, D);
    h = g.norm(y, nullptr, ln[t + "norm3"], A);
  }
  if (hyb) {
    if (out.slot[0]) g.linear(h, D, lin["classifier"], A, ACT_SIGMOID, nullptr, 0, slot_at<float>(out, 0, b0 * N), 1, 0);
  } else {
    float* o = g.linear(h, D, lin["classifier"], A);
    op_improved_head(c, o, A, slot_at<float>(out, 0, b0 * N), slot_at<float>(out, 1, b0 * N), slot_at<float>(out, 2, b0 * N));
  }
}

// f2: ComplexDetectionModel (complex_detection_model.py:63-96)
void Model::fwd_complex(XIn& x, int64_t B, int N, int S, const paut_outputs& out, int64_t b0) {
  Ctx& c = *ctx;
  G g{c, cfg.precision == PAUT_PRECISION_BF16};
  const int64_t A = B * N;
  const int d = cfg.d_model;
  float* pooled = c.allocf((size_t)A * 128);
  conv_stack_pooled(*this, g, x, A, S, "conv_layers.0", "conv_layers.3", "conv_layers.6", /*mode=*/1, 1, pooled);
  float* h = g.linear(pooled, 128, lin["feature_projection.0"], A, ACT_RELU, nullptr, 0, nullptr, 0, 0, 0.f,
                      raw["positional_encoding"], N);
  for (const TELW& t : tel) h = g.encoder_layer(h, t, B, N, ACT_RELU);
  float* h1 = g.linear(h, d, lin["detection_head.0"], A, ACT_RELU);
  if (out.slot[0])
    g.linear(h1, d / 2, lin["detection_head.3"], A, ACT_SIGMOID, nullptr, 0, slot_at<float>(out, 0, b0 * N), 1, 0);
}

// ------------------------------------------------------------------------------------------ driver
void Model::forward(const void* x, int x_dtype, int64_t B, int64_t N, int64_t S, const paut_outputs& out) {
  PAUT_CHECK(finalized, PAUT_ERR_STATE, "forward before paut_model_finalize");
  PAUT_CHECK(x != nullptr, PAUT_ERR_INVALID, "forward: x is null");
  PAUT_CHECK(B > 0 && N > 0 && S > 0, PAUT_ERR_INVALID, "forward: B, N, S must be positive");
  PAUT_CHECK(x_dtype == PAUT_F32 || x_dtype == PAUT_BF16, PAUT_ERR_INVALID, "forward: x dtype must be F32 or BF16");
  PAUT_CHECK(N <= 4096 && S <= 8192, PAUT_ERR_UNSUPPORTED, "forward: N or S too large");
  if (kind == PAUT_MODEL_MSC || kind == PAUT_MODEL_MSC_N) {
    PAUT_CHECK(S == cfg.signal_length, PAUT_ERR_INVALID,
               "forward: signal length differs from the model's signal_length (shared_layer.0 in_features)");
    PAUT_CHECK(N <= 300, PAUT_ERR_INVALID, "forward: more than 300 signals per set (position table has 300 rows)");
  } else if (kind == PAUT_MODEL_MSC_LEGACY) {
    PAUT_CHECK(S == cfg.signal_length, PAUT_ERR_INVALID,
               "forward: signal length differs from the model's signal_length (shared_layer.0 in_features)");
  } else if (kind == PAUT_MODEL_IMPROVED || kind == PAUT_MODEL_HYBRID || kind == PAUT_MODEL_COMPLEX) {
    PAUT_CHECK(kind != PAUT_MODEL_IMPROVED || S == cfg.signal_length, PAUT_ERR_INVALID,
               "forward: signal length differs from the model's signal_length (shared_layer.0 in_features)");
    PAUT_CHECK(kind == PAUT_MODEL_IMPROVED || S >= 128, PAUT_ERR_UNSUPPORTED,
               "forward: signals shorter than 128 samples are not supported for this model");
    PAUT_CHECK(N <= (kind == PAUT_MODEL_HYBRID ? 1200 : 300), PAUT_ERR_INVALID,
               "forward: more signals per set than rows in the position table");
  } else if (kind != PAUT_MODEL_CONV1D_MSC) {
    PAUT_CHECK(N <= 5000, PAUT_ERR_INVALID, "forward: sequence longer than the positional-encoding table");
  }
  PAUT_CHECK(S % 4 == 0, PAUT_ERR_UNSUPPORTED, "forward: signal length must be a multiple of 4");
  PAUT_CUDA(cudaSetDevice(ctx->device));
  Ctx& c = *ctx;
  const size_t esz = x_dtype == PAUT_BF16 ? 2 : 4;
  auto run = [&](int64_t b0, int64_t nb) {
    const char* xc = static_cast<const char*>(x) + (size_t)b0 * N * S * esz;
    if (kind == PAUT_MODEL_CONV1D_MSC) {
      fwd_conv1d_msc(xc, x_dtype, nb, (int)N, (int)S, out, b0);
      return;
    }
    if (kind == PAUT_MODEL_MSC || kind == PAUT_MODEL_MSC_N) {
      fwd_msc(xc, x_dtype, nb, (int)N, (int)S, out, b0);
      return;
    }
    XIn xf;                                   // fp32 / bf16 copies are made lazily by the paths that need them
    xf.p = xc; xf.dtype = x_dtype; xf.n = nb * N * S;
    switch (kind) {
      case PAUT_MODEL_SSD: fwd_ssd(xf, nb, (int)N, (int)S, out, b0, B); break;
      case PAUT_MODEL_ENHANCED: fwd_enhanced(xf, nb, (int)N, (int)S, out, b0, B); break;
      case PAUT_MODEL_TWO_STAGE: fwd_two_stage(xf, nb, (int)N, (int)S, out, b0); break;
      case PAUT_MODEL_MSC_LEGACY: fwd_msc_legacy(xf, nb, (int)N, (int)S, out, b0); break;
      case PAUT_MODEL_IMPROVED:
      case PAUT_MODEL_HYBRID: fwd_improved(xf, nb, (int)N, (int)S, out, b0); break;
      case PAUT_MODEL_COMPLEX: fwd_complex(xf, nb, (int)N, (int)S, out, b0); break;
      default: throw Error(PAUT_ERR_INVALID, "unknown model kind");
    }
  };
  // size the chunk with dry runs (allocation only): the exact footprint of a candidate chunk is measured, and the
  // chunk shrinks proportionally until it fits the workspace limit (the footprint is monotonic in the set count)
  auto footprint = [&](int64_t nb) {
    c.dry = true;
    c.reset();
    run(0, nb);
    const size_t s = c.ws_off;
    c.dry = false;
    c.reset();
    return s;
  };
  const size_t slack = size_t(1) << 20;
  int64_t chunk = B;
  size_t need = footprint(chunk);
  while (need + slack > c.ws_limit && chunk > 1) {
    int64_t next = (int64_t)((double)chunk * (double)c.ws_limit / (double)(need + slack) * 0.97);
    if (next >= chunk) next = chunk - 1;
    if (next < 1) next = 1;
    chunk = next;
    need = footprint(chunk);
  }
  // bf16 conv models: the pooled-mean epilogue sums per 128-row tile of the flat layout, so which rows share a
  // tile depends on the A-scan's index modulo the layout period (16 / 32 / 64 A-scans for the row periods 328 /
  // 164 / 82).  Chunks that start on a multiple of 64 A-scans reproduce the unchunked summation order exactly.
  if (cfg.precision == PAUT_PRECISION_BF16 && kind != PAUT_MODEL_MSC && kind != PAUT_MODEL_MSC_N && chunk < B) {
    int64_t a = 64, b = N;
    while (b) { const int64_t t = a % b; a = b; b = t; }         // a = gcd(64, N)
    const int64_t align = 64 / a;                                  // sets per alignment period
    if (chunk >= align && chunk % align != 0) {
      chunk -= chunk % align;
      need = footprint(chunk);
    }
  }
  c.reserve(need + slack);
  for (int64_t b0 = 0; b0 < B; b0 += chunk) {
    c.reset();
    run(b0, std::min<int64_t>(chunk, B - b0));
  }
}

// Intermediate tensors of the fused kernels (paut_debug_stage): one resident chunk, no chunking logic
void Model::debug_stage(int stage, const void* x, int x_dtype, int64_t B, int64_t N, int64_t S, float* out_dev) {
  PAUT_CHECK(finalized, PAUT_ERR_STATE, "debug_stage before paut_model_finalize");
  PAUT_CHECK(x_dtype == PAUT_F32 || x_dtype == PAUT_BF16, PAUT_ERR_INVALID, "debug_stage: x dtype must be F32 or BF16");
  PAUT_CUDA(cudaSetDevice(ctx->device));
  Ctx& c = *ctx;
  const int64_t A = B * N;
  c.reset();
  c.reserve((size_t)A * S * 4 + (size_t(1) << 20));
  XIn xin;
  xin.p = x; xin.dtype = x_dtype; xin.n = A * S;
  if (stage == 1) {
    PAUT_CHECK(kind == PAUT_MODEL_TWO_STAGE && ts_enc.ready && ts_encoder_supported((int)S, cfg.d_model), PAUT_ERR_UNSUPPORTED,
               "debug_stage 1: the fused two-stage encoder is not available for this model / precision / length");
    op_ts_encoder(c, xin.as_bf16(c), A, (int)S, ts_enc.Wst, ts_enc.W2, ts_enc.shift2, out_dev);
  } else if (stage == 6) {
    PAUT_CHECK(kind == PAUT_MODEL_MSC_N && mscn.ready && mscn_front_supported((int)S), PAUT_ERR_UNSUPPORTED,
               "debug_stage 6: the fused MSC_N front end is not available for this model / precision / length");
    op_mscn_front(c, xin.as_bf16(c), A, (int)S, mscn.W, mscn.b2, mscn.f_const, out_dev);
  } else if (stage >= 2 && stage <= 5) {
    // MSC attention block on an fp32 [B, N, 64] input: stage 2 / 3 = tcgen05 kernel (self / shifted keys and values),
    // stage 4 / 5 = the mma.sync kernel
    PAUT_CHECK(kind == PAUT_MODEL_MSC && set_tc.ready && x_dtype == PAUT_F32 && S == 64, PAUT_ERR_UNSUPPORTED,
               "debug_stage 2-5: MSC model in bf16 mode and an fp32 [B, N, 64] input");
    const bool cross = (stage & 1) != 0;
    const std::string te = "transformer_encoder.";
    const MHAW& a = mha[cross ? "cross" : "self"];
    const LNW& n = ln[te + (cross ? "norm2" : "norm1")];
    if (stage <= 3) {
      PAUT_CHECK(msc_attn_tc_supported((int)N, 64, cfg.num_heads), PAUT_ERR_UNSUPPORTED, "debug_stage: set too long");
      op_msc_attn_tc(c, static_cast<const float*>(x), set_tc.tc_w[cross][0], set_tc.tc_w[cross][1], set_tc.tc_w[cross][2],
                     a.in_proj.b, a.out_proj.b, n.g, n.b, out_dev, B, (int)N, cross);
    } else {
      op_msc_attn_block(c, static_cast<const float*>(x), cross ? set_tc.Wqkv_cross : set_tc.Wqkv_self, a.in_proj.b,
                        cross ? set_tc.Wo_cross : set_tc.Wo_self, a.out_proj.b, n.g, n.b, out_dev, B, (int)N, cross);
    }
  } else {
    throw Error(PAUT_ERR_INVALID, "debug_stage: unknown stage");
  }
}

void Model::postprocess(const paut_outputs& o, int64_t B, int64_t N, int64_t S, double thr, paut_detection* det,
                        int32_t* count_dev) {
  PAUT_CHECK(det && count_dev, PAUT_ERR_INVALID, "postprocess: null output");
  PAUT_CUDA(cudaSetDevice(ctx->device));
  PostArgs a;
  a.kind = kind; a.B = B; a.N = (int)N; a.S = (int)S; a.threshold = thr; a.C = cfg.num_classes;
  auto need = [&](int i) {
    PAUT_CHECK(o.slot[i] != nullptr, PAUT_ERR_INVALID, "postprocess: required forward output slot is null");
    return static_cast<const float*>(o.slot[i]);
  };
  switch (kind) {
    case PAUT_MODEL_MSC:
    case PAUT_MODEL_MSC_N: a.score_src = need(0); a.start = need(1); a.end = need(2); break;
    case PAUT_MODEL_CONV1D_MSC: a.score_src = need(0); break;
    case PAUT_MODEL_MSC_LEGACY: a.score_src = need(0); a.cmp = 1; break;
    case PAUT_MODEL_IMPROVED: a.score_src = need(0); a.start = need(1); a.end = need(2); a.cmp = 1; break;
    case PAUT_MODEL_HYBRID: a.score_src = need(0); a.cmp = 2; break;
    case PAUT_MODEL_COMPLEX: a.score_src = need(0); a.cmp = 3; break;
    case PAUT_MODEL_SSD: a.score_src = need(0); a.pos = need(1); a.anomaly = need(2); break;
    case PAUT_MODEL_ENHANCED: a.score_src = need(0); a.unc = need(1); a.pos = need(2); a.anomaly = need(4); break;
    case PAUT_MODEL_TWO_STAGE: a.score_src = need(1); a.unc = need(2); a.pos = need(3); break;
    default: throw Error(PAUT_ERR_INVALID, "unknown model kind");
  }
  ctx->reset();
  ctx->reserve(sizeof(int32_t) * 2 * (size_t)B + 1024);
  op_postprocess(*ctx, a, det, count_dev);
}

}  // namespace paut
