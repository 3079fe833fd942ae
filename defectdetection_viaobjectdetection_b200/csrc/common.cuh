// Internal declarations shared by the translation units of libpaut.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#include <map>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/paut.h"

namespace paut {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define PAUT_CUDA(expr)                                                                            \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess)                                                                         \
      throw ::paut::Error(PAUT_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));      \
  } while (0)

#define PAUT_CHECK(cond, code, msg)                                                                \
  do {                                                                                             \
    if (!(cond)) throw ::paut::Error(code, std::string(msg));                                      \
  } while (0)

// One context = one GPU = one stream.  Owns the activation workspace (bump arena).
struct Ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  std::string last_error;
  char* ws = nullptr;
  size_t ws_cap = 0, ws_off = 0;
  size_t ws_limit = size_t(2) << 30;
  int64_t launches = 0;
  int num_sms = 148;
  int smem_optin = 0;
  bool profiling = false;                // per-kernel CUDA-event timing (paut_ctx_profile_*)
  std::vector<std::pair<std::string, cudaEvent_t>> prof_events;
  bool dry = false;                      // allocation-only pass used to size chunks: ops do nothing

  void reserve(size_t bytes);            // grow workspace (synchronises only when growing)
  void* alloc(size_t bytes);             // bump allocation, 256-byte aligned
  float* allocf(size_t n) { return static_cast<float*>(alloc(n * sizeof(float))); }
  void reset() { ws_off = 0; }
  void launched(const char* what);       // counts the launch and checks cudaGetLastError
};

// cudaFuncAttributeMaxDynamicSharedMemorySize is a property of (device, kernel), not of a context: every kernel that
// needs more than 48 KB is opted in ONCE per device to everything the device allows beside the kernel's static
// shared memory (process-wide table behind a mutex, api.cu), so that contexts with different shapes on different
// streams can never lower each other's limit.
void smem_optin_once(Ctx& c, const void* kernel);
template <typename K>
inline void smem_optin(Ctx& c, K kernel) { smem_optin_once(c, reinterpret_cast<const void*>(kernel)); }

enum Act { ACT_NONE = 0, ACT_RELU = 1, ACT_GELU = 2, ACT_SIGMOID = 3, ACT_SOFTPLUS = 4, ACT_TANH_HALF = 5 };

// ---------------------------------------------------------------------------------------------
// fp32 CUDA-core operators (ops_f32.cu).  Activations are channels-last: [rows, C], C contiguous.
// ---------------------------------------------------------------------------------------------
// [B,S,N] -> [B,N,S] (DefectDetectionModel's permute, MSC_Conv1D_training.py:81), any input dtype -> fp32
void op_transpose_sn(Ctx& c, const void* x, int x_dtype, float* out, int64_t B, int S, int N);
// any dtype -> fp32 copy
void op_to_f32(Ctx& c, const void* x, int x_dtype, float* out, int64_t n);
void op_to_bf16(Ctx& c, const float* x, void* out, int64_t n);

// Conv1d with C_in = 1 (+ folded BN, optional ReLU): x [A,S] -> out [A,S,Cout].  w is [k][Cout].
void op_stem_conv(Ctx& c, const float* x, int64_t A, int S, const float* w, const float* shift, int k,
                  int Cout, bool relu, float* out);

struct ConvArgs {
  const float* in = nullptr;   // [A, Lin, Cin]
  int64_t A = 0;
  int Lin = 0, Cin = 0;
  const float* w = nullptr;    // [taps][Cin][Cout], BN scale folded
  const float* shift = nullptr;
  int Cout = 0, taps = 1, dil = 1, stride = 1, pad = 0, Lout = 0;
  bool relu = false;
  const float* res = nullptr;  // [A, Lout, ldr], added before the ReLU
  int ldr = 0;
  float* out = nullptr;        // [A, Lout, ldc] written at column offset coff (nullable)
  int ldc = 0, coff = 0;
  float* pool = nullptr;       // [A, ldp]: mean over Lout written at column offset poff (nullable)
  int ldp = 0, poff = 0;
};
void op_conv(Ctx& c, const ConvArgs& a);

struct LinArgs {
  const float* A = nullptr;    // [M, lda]
  int lda = 0;
  const float* Wt = nullptr;   // [K][N]  (transposed nn.Linear weight)
  const float* W = nullptr;    // [N][K]  (original layout, used when N is tiny)
  const void* Wp = nullptr;    // bf16, packed for tcgen05 (ops_tc.cu); null in fp32 mode
  int NT = 0;                  // N tile of the packed weight
  int stream_b = 0;            // the packed weight is streamed through the ring (tc_pick_ntile)
  const float* bias = nullptr;
  int64_t M = 0;
  int K = 0, N = 0;
  float* C = nullptr;          // [M, ldc] written at column offset coff
  int ldc = 0, coff = 0;
  int act = ACT_NONE;
  float act_eps = 0.f;         // added after the activation (two-stage softplus + 1e-6)
  const float* res = nullptr;  // [M, ldr] added after the activation
  int ldr = 0;
  const float* table = nullptr;  // [table_mod, N] row (m % table_mod) added after the activation
  int table_mod = 1;
};
void op_linear(Ctx& c, const LinArgs& a);
// tcgen05 path (ops_tc.cu): bf16 operands, fp32 accumulation in TMEM
int tc_pick_ntile(int N, int K, int* stream = nullptr);
void tc_pack_weight(const float* W, int N, int K, int NT, std::vector<uint16_t>& out, int* Kp_out);
bool linear_tc_supported(const LinArgs& a);
void op_linear_tc(Ctx& c, const LinArgs& a);

// out = act(LN(x + res)) row-wise; res nullable; in-place allowed.
void op_layernorm(Ctx& c, const float* x, const float* res, const float* gamma, const float* beta, float* out,
                  int64_t M, int D, int act);

// Multi-head attention core (softmax(QK^T/sqrt(hd)) V) for B sets.  q/k/v may be column slices of a
// packed buffer (leading dimensions ldq/ldk/ldv).  kv_shift: key/value row j reads row min(j+1, Nk-1)
// (the MSC cross-attention, NN_models.py:35).  avgw (nullable) [B,Nq,Nk] gets the head-averaged weights.
void op_attention(Ctx& c, const float* q, int ldq, const float* k, int ldk, const float* v, int ldv, float* out,
                  int ldo, int64_t B, int Nq, int Nk, int H, int hd, bool kv_shift, float* avgw);

// bf16-mode attention (ops_attn.cu): same contract, bf16 operands on the tensor cores (mma.sync), fp32 softmax.
bool attention_bf16_supported(int Nq, int Nk, int hd, bool want_weights);
void op_attention_bf16(Ctx& c, const float* q, int ldq, const float* k, int ldk, const float* v, int ldv, float* out,
                       int ldo, int64_t B, int Nq, int Nk, int H, int hd, bool kv_shift, float* avgw);

// depthwise conv along the set axis (LocalAttention_N, NN_models.py:159-167): x [B,N,D], w [D][k]
void op_dwconv_seq(Ctx& c, const float* x, const float* w, const float* bias, float* out, int64_t B, int N,
                   int D, int k);

// Bidirectional recurrent layer.  gi [B,T,2*G*H] holds x W_ih^T + b_ih for both directions (forward
// first); whh_t [2][H][G*H]; bhh [2][G*H]; out [B,T,2H].  G = 3 (GRU, gates r|z|n) or 4 (LSTM, i|f|g|o).
void op_rnn_bidir(Ctx& c, const float* gi, const float* whh_t, const float* bhh, float* out, int64_t B, int T,
                  int H, int G);

// softmax over the sequence axis: x [B,N] (row stride N) -> out [B,N]
void op_softmax_seq(Ctx& c, const float* x, float* out, int64_t B, int N);
// out[m,:] = a[m,:] * s[m] + b[m,:]   (b nullable)
void op_rowscale_add(Ctx& c, const float* a, const float* s, const float* b, float* out, int64_t M, int D);
// out[m] = sum_c x[m,c] * q[c]
void op_rowdot(Ctx& c, const float* x, const float* q, float* out, int64_t M, int D);
// out[m,:] = x[m,:] + table[m % mod, :]   (positional encodings)
void op_add_table(Ctx& c, const float* x, const float* table, int mod, float* out, int64_t M, int D);
// dst[m, coff:coff+D] = src[m, :]
void op_copy_cols(Ctx& c, const float* src, int lds, float* dst, int ldd, int coff, int64_t M, int D);

// MSC front end (NN_models.py:111-115 / :227-234): per A-scan conv 1->8 k3 ReLU, 8->16 k3 ReLU,
// optional depthwise k11 background subtraction, mean over channels -> f [A,S].
void op_msc_front(Ctx& c, const float* x, int64_t A, int S, const float* w1, const float* b1, const float* w2,
                  const float* b2, const float* wbg, const float* bbg, float* f);
// Fused tcgen05 encoder of MultiSignalClassifier (ops_msc_tc.cu): x -> h [A,64] (bf16 mode)
bool msc_encoder_tc_supported(int S, int h0, int h1);
void msc_pack_conv2(const float* w2, const float* b2, std::vector<uint16_t>& out);
// w1_host / b1_host: conv1d.0 weight [8][3] and bias [8] in HOST memory (they travel as kernel parameters)
void op_msc_encoder_tc(Ctx& c, const void* x, int x_dtype, int64_t A, int S, int Nset, const float* w1_host, const float* b1_host,
                       const void* Bc, const void* W1p, const float* bl1, const void* W2p, const float* bl2,
                       const float* pos, float* h);
// ---- tcgen05 implicit-GEMM Conv1d on "flat rows" (ops_conv_tc.cu), bf16 mode
constexpr int CONV_HALO = 8;                       // zero rows between A-scans; covers pad*dilation <= 8
struct ConvTcLaunch {
  const void* in = nullptr;      // bf16 flat rows [R, Cin]
  int64_t A = 0;
  int L = 0, halo = CONV_HALO, Cin = 0, Cout = 0;
  int Lp = 0, H0 = 0;            // explicit row geometry row(a,l) = H0 + a*Lp + l (0: the (L, halo) flat layout)
  int NT = 0, CB = 0;            // tile plan the weights were packed with (conv_tc_plan)
  bool skip_lo = false;          // space-to-depth stride-2 view: tap 0 only multiplies the upper half of Cin
  int groups = 0;                // > 0: grouped mode (conv_tc_pack_grouped); taps = the largest tap count
  int gtaps[4] = {0, 0, 0, 0}, goff[4] = {0, 0, 0, 0};
  bool shared_input = false;     // groups read the same Cin (= CB) channels with dilations gdil[g]; dil = the largest
  int gdil[4] = {1, 1, 1, 1};
  const void* Wp = nullptr;      // conv_tc_pack() layout
  const float* shift = nullptr;
  int taps = 1, dil = 1, pad = 0;  // pad in taps (PyTorch padding = pad * dil)
  bool relu = false;
  const void* res = nullptr;     // bf16 flat rows [R, ldr], added before the ReLU
  int ldr = 0;
  void* out = nullptr;           // bf16 flat rows [R, ldc] at column offset coff (nullable)
  int ldc = 0, coff = 0;
  float* pool_partial = nullptr; // scratch [ceil(R/128)][3][Cout]
  float* pool_out = nullptr;     // [A, ldp] mean over L at column offset poff
  int ldp = 0, poff = 0;
};
bool conv_tc_plan(int Cin, int taps, int Cout, int max_dil, int* NT, int* CB);
void conv_tc_pack(const float* w, int taps, int Cin, int Cout, int NT, int CB, std::vector<uint16_t>& out);
void conv_tc_pack_grouped(const float* const* w, const int* taps, int groups, int CB, int GN, std::vector<uint16_t>& out,
                          int* goff16);
size_t flat_rows(int64_t A, int L, int halo);
void op_conv_tc(Ctx& c, const ConvTcLaunch& a);
void op_stem_flat(Ctx& c, const void* x, int x_dtype, int64_t A, int S, const float* w, const float* shift, int k,
                  int Cout, bool relu, void* out, int ldc, int coff, int halo);
// bf16 flat rows [R, C] -> dense fp32 [A, L, C]
void op_unflatten(Ctx& c, const void* flat, int64_t A, int L, int halo, int C, float* out);
// Fused front end of MultiSignalClassifier_N (ops_mscn_front.cu, bf16 mode): x [A,S] bf16 -> f [A,S] fp32 (conv1, conv2 and
// the background-subtraction + channel-mean stencil as three chained tcgen05 stages over two shared-memory rings)
bool mscn_front_supported(int S);
void mscn_front_pack(const float* w1, const float* b1, const float* w2, const float* wbg, std::vector<uint16_t>& W);
void op_mscn_front(Ctx& c, const void* x_bf16, int64_t A, int S, const void* W, const float* b2_host, float f_const, float* f);
// Fused encoder of TwoStageDefectDetector (ops_ts_enc.cu, bf16 mode): x [A,S] bf16 -> feat [A,128] (mean over the signal
// length of the four conv branches), TMA input staging, stem and second convolutions on tcgen05, pooled epilogue
bool ts_encoder_supported(int S, int d_model);
void ts_encoder_pack(const float* const* w1, const float* const* sh1, const float* const* w2, std::vector<uint16_t>& Wst,
                     std::vector<uint16_t>& W2);
void op_ts_encoder(Ctx& c, const void* x_bf16, int64_t A, int S, const void* Wst, const void* W2, const float* shift2, float* feat);
// Fused per-set stage of MSC / MSC_N (ops_set_tc.cu, bf16 mode): attention block and FFN + head
bool msc_set_tc_supported(int N, int d, int heads, int ff);
// optional tail of the second attention block of MultiSignalClassifier: FFN + LayerNorm + classifier in its epilogue
struct MscTail {
  const void* W1; const float* b1; const void* W2; const float* b2; const float* ln_g; const float* ln_b; const void* Wc; const float* bc;
  float* prob; float* start; float* end;
};
bool msc_attn_tail_supported(const Ctx& c, int N);
void op_msc_attn_block(Ctx& c, const float* x, const void* Wqkv, const float* bqkv, const void* Wo, const float* bo,
                       const float* ln_g, const float* ln_b, float* out, int64_t B, int N, bool kv_shift, const MscTail* tail = nullptr);
// the same attention block with every product on tcgen05 (ops_attn_tc.cu): S and P in tensor memory
bool msc_attn_tc_supported(int N, int d, int heads);
void msc_attn_tc_pack(const float* W, int row0, int rows, std::vector<uint16_t>& out);
void op_msc_attn_tc(Ctx& c, const float* x, const void* Wqk, const void* Wv, const void* Wo, const float* bqkv, const float* bo,
                    const float* ln_g, const float* ln_b, float* out, int64_t B, int N, bool kv_shift);
bool lin_res_ln_supported(int N, int K, int lda);
void op_lin_res_ln(Ctx& c, const float* x, int lda, const void* Wrow, int K, const float* bias, const float* res, const float* g,
                   const float* b, float* out, int64_t M, const float* table = nullptr, int table_mod = 0);
bool ts_heads_supported(int d, int hidden);
void op_ts_heads(Ctx& c, const float* x, const float* ng, const float* nb, const void* const* W0, const float* const* b0,
                 const float* const* lg, const float* const* lb, const float* const* W4, const float* const* b4, const int* act,
                 const float* eps, float* const* out, int64_t M);
void op_msc_ffn_head(Ctx& c, const float* x, const float* pre, const float* pre_g, const float* pre_b, const void* W1,
                     const float* b1, const void* W2, const float* b2, const float* ln_g, const float* ln_b,
                     const void* Wc, const float* bc, float* prob, float* start, float* end, int64_t M);
void op_debug_mma(Ctx& c, int mode, int N, int reps, int lbo, int alt, float* out_dev);
// MSC head (NN_models.py:123-127): o [M,3] -> sigmoid / tanh*0.5+0.5 into three arrays
void op_msc_head(Ctx& c, const float* o, int64_t M, float* prob, float* start, float* end);
// logits[:, 1:] += anomaly (model.py:332, enhanced_model.py:550)
void op_add_anomaly(Ctx& c, float* logits, const float* anomaly, int64_t M, int C);
// two-stage finalisation (two_stage_model.py:294-301): probs = softmax(logits); pos *= probs[:,1]
void op_two_stage_final(Ctx& c, const float* logits, float* probs, float* pos, int64_t M);

// ---------------------------------------------------------------------------------------------
// post-processing and windowing (ops_post.cu)
// ---------------------------------------------------------------------------------------------
struct PostArgs {
  int kind = 0;
  const float* score_src = nullptr;   // class logits [M,C] (SSD/ENHANCED), defect_probs [M,2], or prob [M]
  const float* unc = nullptr;         // class_uncertainty [M,C] / defect_uncertainty [M,2]
  const float* pos = nullptr;         // position_preds [M,2] (or nullptr for MSC kinds)
  const float* start = nullptr;       // MSC: defect_start [M]
  const float* end = nullptr;         // MSC: defect_end [M]
  const float* anomaly = nullptr;     // [M]
  int C = 2;
  int64_t B = 0;
  int N = 0, S = 0;
  double threshold = 0.5;
  int cmp = 0;                        // MSC-family keep rule: 0 fp64 '>', 1 fp64 '>=', 2 fp32 '>=', 3 fp32 '>'
};
void op_postprocess(Ctx& c, const PostArgs& a, paut_detection* det, int32_t* count_dev);
void op_window_gather(Ctx& c, const void* volume, int src_dtype, int64_t G, int64_t n, int S,
                      const int32_t* table, int64_t W, int L, void* sets, int dst_dtype);


// ---------------------------------------------------------------------------------------------
// SURVEY section 8 "next" rows (ops_next.cu)
// ---------------------------------------------------------------------------------------------
// improved_model.py:126-133: f[a,l] = mean_c(x - depthwise_conv_k(x)); rows row(a,l) = H0 + a*Lrow + l of a
// [rows, C] fp32 or bf16 buffer, w [C][k]
void op_bgsub_chanmean(Ctx& c, const void* in, int in_dtype, int64_t A, int L, int C, int Lrow, int H0, int k,
                       const float* w, const float* bias, float* f);
// channel mean + (mode 0: AvgPool1d(pk) -> linear interpolate to P | mode 1: AdaptiveAvgPool1d(P)); rows
// row(a,l) = H0 + a*Lp + l of a [rows, C] fp32 or bf16 buffer -> out [A, P]
void op_chanmean_resample(Ctx& c, const void* in, int in_dtype, int64_t A, int L, int C, int Lp, int H0, int mode,
                          int pk, int P, float* out);
// hybrid_binary.py:147-149: out [B,N,2D] = cat(seq, seq - mean over N)
void op_seqmean_concat(Ctx& c, const float* seq, int64_t B, int N, int D, float* out);
// improved_model.py:147-156: sigmoid / clamp / clamp of o [M,3]
void op_improved_head(Ctx& c, const float* o, int64_t M, float* prob, float* start, float* end);
// dataset_preparation.py:205: flags[g] = any(volume[g] != 0)
void op_group_nonzero(Ctx& c, const void* vol, int dtype, int64_t G, int64_t per_group, int32_t* flags);
// teststtt.py:54-69
void op_difference_matrix(Ctx& c, const void* x, int x_dtype, const float* prob, int64_t B, int N, int S, double thr,
                          float* ref, float* diff, int32_t* healthy);
// two_stage_train.py:284-375 (rule 0), train.py:279-361 (rule 1), acc_metrics_hybrid_binary_dynamic_.py:73-94
void op_metrics_match(Ctx& c, int rule, const paut_detection* det, const int32_t* count_dev, int64_t B, int N,
                      const int32_t* tlabel, const float* tpos, double thr, paut_metrics* out);
void op_metrics_confusion(Ctx& c, const float* prob, const float* label, int64_t M, double thr, int ge,
                          paut_metrics* out);

}  // namespace paut
