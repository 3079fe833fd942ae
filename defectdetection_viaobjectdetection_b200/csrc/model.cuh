// Model object of libpaut.so: state_dict contract, packed weights, forward graphs.
#pragma once
#include "common.cuh"

namespace paut {

struct KeySpec {
  std::string key;
  std::vector<int64_t> shape;
};

struct HostTensor {
  std::vector<float> data;
  std::vector<int64_t> shape;
};

// Packed parameter views (device pointers owned by Model::dev_allocs)
struct Lin {
  const float* Wt = nullptr;  // [K][N]
  const float* W = nullptr;   // [N][K]
  const float* b = nullptr;
  const void* Wp = nullptr;   // bf16 packed for tcgen05 (bf16 mode only)
  const void* Wrow = nullptr; // plain bf16 [N][K] for the mma.sync Linear + residual + LayerNorm kernel (N = 128, bf16 mode)
  int NT = 0;
  int stream_b = 0;              // tcgen05 GEMM mode chosen with NT (weights resident or streamed)
  int K = 0, N = 0;
};
struct ConvW {
  const float* w = nullptr;      // [taps][Cin][Cout] (BN scale folded); Cin == 1 -> [taps][Cout]
  const float* shift = nullptr;  // folded bias
  const void* Wp = nullptr;      // bf16 blocks for the tcgen05 conv (bf16 mode, Cin and Cout multiples of 16)
  int NT = 0, CB = 0;            // tile plan of Wp
  const void* Wp_s2 = nullptr;   // stride-2 k3 conv as a 2-tap conv over the space-to-depth view (2*Cin channels)
  int NT_s2 = 0, CB_s2 = 0;
  int Cin = 0, Cout = 0, taps = 0;
};
struct LNW {
  const float* g = nullptr;
  const float* b = nullptr;
  int D = 0;
};
struct MHAW {
  Lin in_proj, out_proj;
  int D = 0, H = 0;
  int hd = 0, Dp = 0;            // head_dim / width the attention kernel sees (padded in bf16 mode when head_dim = 8)
};
struct TELW {  // post-norm transformer encoder layer
  MHAW attn;
  Lin l1, l2;
  LNW n1, n2;
};
struct RNNW {  // one bidirectional layer
  Lin ih;                        // N = 2*G*H (forward rows first), bias = b_ih
  const float* whh_t = nullptr;  // [2][H][G*H]
  const float* bhh = nullptr;    // [2][G*H]
  int H = 0, G = 0;
};

// Input of one resident chunk in the caller's dtype; the fp32 / bf16 copies are made lazily, only by the paths
// that need them (the bf16-mode stems and the fused encoders read bf16 directly)
struct XIn {
  const void* p = nullptr;
  int dtype = PAUT_F32;
  int64_t n = 0;
  const float* f32 = nullptr;
  const void* bf16 = nullptr;
  const float* as_f32(Ctx& c) {
    if (dtype == PAUT_F32) return static_cast<const float*>(p);
    if (!f32) { float* t = c.allocf((size_t)n); op_to_f32(c, p, dtype, t, n); f32 = t; }
    return f32;
  }
  const void* as_bf16(Ctx& c) {
    if (dtype == PAUT_BF16) return p;
    if (!bf16) { void* t = c.alloc((size_t)n * 2); op_to_bf16(c, static_cast<const float*>(p), t, n); bf16 = t; }
    return bf16;
  }
};

struct Model {
  Ctx* ctx = nullptr;
  int kind = 0;
  paut_model_cfg cfg{};
  std::vector<KeySpec> spec;
  std::map<std::string, HostTensor> host;
  bool finalized = false;
  std::vector<void*> dev_allocs;

  // ---- packed parameters (filled by finalize) ----
  std::map<std::string, Lin> lin;
  std::map<std::string, ConvW> conv;
  std::map<std::string, LNW> ln;
  std::map<std::string, MHAW> mha;
  std::vector<TELW> tel;
  std::vector<RNNW> rnn;
  std::map<std::string, const float*> raw;   // tensors uploaded as they are
  struct EncTc {                               // operands of the fused tcgen05 MSC encoder (bf16 mode)
    const void* Bc = nullptr;
    const void* W1p = nullptr;
    const void* W2p = nullptr;
    bool ready = false;
  } enc_tc;
  struct GroupedConv {                         // the two-stage encoder's four second convolutions as one launch
    const void* Wp = nullptr;
    const float* shift = nullptr;              // [4 * q]
    int taps[4] = {0, 0, 0, 0}, goff[4] = {0, 0, 0, 0};
    bool ready = false;
  } ts_grouped;
  struct MscnFront {                           // operands of the fused MSC_N front end (ops_mscn_front.cu, bf16 mode)
    const void* W = nullptr;
    float b2[16] = {0};
    float f_const = 0.f;
    bool ready = false;
  } mscn;
  struct TsEnc {                               // operands of the fused two-stage encoder (ops_ts_enc.cu, bf16 mode)
    const void* Wst = nullptr;                 // the four stems + folded BN shift as one [128 x 16] bf16 operand
    const void* W2 = nullptr;                  // second convolutions, fp16, packed
    const float* shift2 = nullptr;             // [128]
    bool ready = false;
  } ts_enc;
  struct TsHeads {                             // first Linear of the four two-stage heads as plain bf16 [64][128] (ops_set_tc.cu)
    const void* W0[4] = {nullptr, nullptr, nullptr, nullptr};
    bool ready = false;
  } ts_heads;
  struct BranchConv {                          // the enhanced encoder's four dilated branches as one launch
    const void* Wp = nullptr;
    const float* shift = nullptr;              // [128] (the conv biases)
    int goff[4] = {0, 0, 0, 0};
    bool ready = false;
  } enh_branches;
  struct SetTc {                               // plain bf16 [N][K] weights of the fused set-stage kernels
    const void* Wqkv_self = nullptr; const void* Wo_self = nullptr;
    const void* Wqkv_cross = nullptr; const void* Wo_cross = nullptr;
    const void* W1 = nullptr; const void* W2 = nullptr; const void* Wc = nullptr;
    // K-major operand packing of the tcgen05 attention block: [self | cross][Wqk, Wv, Wo]
    const void* tc_w[2][3] = {{nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}};
    bool ready = false;
  } set_tc;

  ~Model();
  void build_spec();
  void set_tensor(const char* key, const void* ptr, int dtype, const int64_t* shape, int ndim);
  void finalize();
  void forward(const void* x, int x_dtype, int64_t B, int64_t N, int64_t S, const paut_outputs& out);
  void debug_stage(int stage, const void* x, int x_dtype, int64_t B, int64_t N, int64_t S, float* out_dev);
  void postprocess(const paut_outputs& outs, int64_t B, int64_t N, int64_t S, double thr, paut_detection* det,
                   int32_t* count_dev);

  // helpers used by finalize
  const HostTensor& H(const std::string& key) const;
  const float* upload(const std::vector<float>& v);
  Lin pack_lin(const std::string& name);
  Lin make_lin(const std::vector<float>& W, const std::vector<float>& bias, int rows, int K);
  void pack_tc(Lin& l, const std::vector<float>& W);
  Lin pack_lin_rows(const std::string& wkey, const std::string& bkey, int row0, int rows);
  ConvW pack_conv(const std::string& conv_name, const std::string& bn_name, int max_dil = 1, bool stride2 = false);
  LNW pack_ln(const std::string& name);
  MHAW pack_mha(const std::string& name, int heads);
  TELW pack_tel(const std::string& name, int heads);
  RNNW pack_rnn(const std::string& name, int layer, int G, int Hh);

  // forward graphs (one chunk of whole sets)
  void fwd_msc(const void* x, int x_dtype, int64_t B, int N, int S, const paut_outputs& out, int64_t b0);
  void fwd_conv1d_msc(const void* x, int x_dtype, int64_t B, int N, int S, const paut_outputs& out, int64_t b0);
  void fwd_ssd(XIn& x, int64_t B, int N, int S, const paut_outputs& out, int64_t b0, int64_t Btot);
  void fwd_enhanced(XIn& x, int64_t B, int N, int S, const paut_outputs& out, int64_t b0, int64_t Btot);
  void fwd_two_stage(XIn& x, int64_t B, int N, int S, const paut_outputs& out, int64_t b0);
  // SURVEY section 8 "next" rows f2 / f3
  void fwd_msc_legacy(XIn& x, int64_t B, int N, int S, const paut_outputs& out, int64_t b0);
  void fwd_improved(XIn& x, int64_t B, int N, int S, const paut_outputs& out, int64_t b0);
  void fwd_complex(XIn& x, int64_t B, int N, int S, const paut_outputs& out, int64_t b0);
};

}  // namespace paut
