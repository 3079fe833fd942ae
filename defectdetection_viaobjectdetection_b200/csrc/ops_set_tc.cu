// Fused per-set stage of MultiSignalClassifier / _N for the bf16 mode (d_model 64, 4 heads of 16, ffn 32):
//
//   k_msc_attn_block : x[N,64] -> QKV projection -> softmax(QK^T/4) V per head -> out-projection -> + x -> LayerNorm
//                      (TransformerEncoder.forward NN_models.py:31-37; with kv_shift the keys/values come from the
//                      sequence shifted left by one with the last row repeated, :35)
//   k_msc_ffn_head   : x -> Linear 64->32 ReLU -> Linear 32->64 -> + x -> LayerNorm -> Linear 64->3 ->
//                      sigmoid / tanh*0.5+0.5                                 (NN_models.py:39-41, :123-127)
//
// One CTA of the attention block owns one set: all tokens' K and V^T (bf16) and the two weight matrices live in
// shared memory, every other intermediate (Q, scores, probabilities, attention output, projection, LayerNorm
// statistics) stays in registers as mma.sync fragments.  The products are warp-level mma.sync m16n8k16
// (bf16 -> fp32): the stage is bound by the exp of the softmax (MUFU) and by shared-memory fragment loads,
// not by the tensor pipe, so the accumulators are better off in registers than in TMEM.
#include "common.cuh"
#include "mma_common.cuh"

namespace paut {

using namespace mma;

namespace {

constexpr int DM = 64, HD = 16, NH = 4, FF = 32;
constexpr int WS = DM + 8;            // bf16 row stride of K rows and of weight rows with K = 64
constexpr int AB_WARPS = 20;
constexpr int KBLK = 64;              // keys per softmax block

struct AttnBlockArgs {
  const float* x;                     // [B, N, 64]
  const __nv_bfloat16* Wqkv;          // [192][64] (q | k | v rows)
  const float* bqkv;                  // [192]
  const __nv_bfloat16* Wo;            // [64][64]
  const float* bo;
  const float* ln_g;
  const float* ln_b;
  float* out;                         // [B, N, 64]
  int N;
  int kv_shift;
};

// A fragments (4 k-steps) of the 16 x 64 fp32 tile starting at row r0 of one set; rows >= N are zero
__device__ __forceinline__ void load_x_frags(const float* __restrict__ xs, int r0, int N, int g, int t,
                                             uint32_t (&xa)[4][4]) {
  const int row_lo = r0 + g, row_hi = r0 + g + 8;
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    float2 a = make_float2(0.f, 0.f), b = a, c = a, d = a;
    if (row_lo < N) {
      const float* p = xs + (size_t)row_lo * DM + ks * 16 + 2 * t;
      a = __ldg(reinterpret_cast<const float2*>(p));
      c = __ldg(reinterpret_cast<const float2*>(p + 8));
    }
    if (row_hi < N) {
      const float* p = xs + (size_t)row_hi * DM + ks * 16 + 2 * t;
      b = __ldg(reinterpret_cast<const float2*>(p));
      d = __ldg(reinterpret_cast<const float2*>(p + 8));
    }
    xa[ks][0] = pack_bf16(a.x, a.y);
    xa[ks][1] = pack_bf16(b.x, b.y);
    xa[ks][2] = pack_bf16(c.x, c.y);
    xa[ks][3] = pack_bf16(d.x, d.y);
  }
}

// one 8-column output tile of  X[16 x 64] * W[rows n0..n0+7][64]^T + bias, W rows in shared memory (stride WS)
__device__ __forceinline__ void proj_tile(const uint32_t (&xa)[4][4], const __nv_bfloat16* __restrict__ W, int n0,
                                          const float* __restrict__ bias, int g, int t, float (&c)[4]) {
  const float b0 = bias ? __ldg(bias + n0 + 2 * t) : 0.f, b1 = bias ? __ldg(bias + n0 + 2 * t + 1) : 0.f;
  c[0] = b0; c[1] = b1; c[2] = b0; c[3] = b1;
  const __nv_bfloat16* wr = W + (size_t)(n0 + g) * WS + 2 * t;
#pragma unroll
  for (int ks = 0; ks < 4; ++ks)
    mma_bf16_16816(c, xa[ks], *reinterpret_cast<const uint32_t*>(wr + ks * 16),
                   *reinterpret_cast<const uint32_t*>(wr + ks * 16 + 8));
}

// LayerNorm over the 64 columns of a 16-row tile held as 8 accumulator tiles; rows g and g+8 of the quad
__device__ __forceinline__ void layer_norm_tile(float (&y)[8][4], const float* __restrict__ gam,
                                                const float* __restrict__ bet, int t) {
  float s0 = 0.f, s1 = 0.f;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) { s0 += y[nt][0] + y[nt][1]; s1 += y[nt][2] + y[nt][3]; }
  s0 += __shfl_xor_sync(0xffffffffu, s0, 1); s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
  s1 += __shfl_xor_sync(0xffffffffu, s1, 1); s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
  const float m0 = s0 * (1.f / DM), m1 = s1 * (1.f / DM);
  float q0 = 0.f, q1 = 0.f;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    const float a = y[nt][0] - m0, b = y[nt][1] - m0, c = y[nt][2] - m1, d = y[nt][3] - m1;
    q0 += a * a + b * b;
    q1 += c * c + d * d;
  }
  q0 += __shfl_xor_sync(0xffffffffu, q0, 1); q0 += __shfl_xor_sync(0xffffffffu, q0, 2);
  q1 += __shfl_xor_sync(0xffffffffu, q1, 1); q1 += __shfl_xor_sync(0xffffffffu, q1, 2);
  const float r0 = rsqrtf(q0 * (1.f / DM) + 1e-5f), r1 = rsqrtf(q1 * (1.f / DM) + 1e-5f);
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    const int col = nt * 8 + 2 * t;
    const float g0 = __ldg(gam + col), g1 = __ldg(gam + col + 1), b0 = __ldg(bet + col), b1 = __ldg(bet + col + 1);
    y[nt][0] = (y[nt][0] - m0) * r0 * g0 + b0;
    y[nt][1] = (y[nt][1] - m0) * r0 * g1 + b1;
    y[nt][2] = (y[nt][2] - m1) * r1 * g0 + b0;
    y[nt][3] = (y[nt][3] - m1) * r1 * g1 + b1;
  }
}

__global__ void __launch_bounds__(AB_WARPS * 32, 1) k_msc_attn_block(AttnBlockArgs p) {
  extern __shared__ __align__(16) unsigned char smraw[];
  const int N = p.N;
  const int Np = (N + KBLK - 1) / KBLK * KBLK;
  const int vts = Np + 8;
  __nv_bfloat16* Kb = reinterpret_cast<__nv_bfloat16*>(smraw);        // [Np][WS]
  __half* Vt = reinterpret_cast<__half*>(Kb + (size_t)Np * WS);           // [64][Np + 8], fp16: the P V product runs in fp16
  __nv_bfloat16* Wq = reinterpret_cast<__nv_bfloat16*>(Vt + (size_t)DM * vts);   // [192][WS]
  __nv_bfloat16* Wo = Wq + (size_t)3 * DM * WS;                         // [64][WS]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const float* xs = p.x + (size_t)blockIdx.x * N * DM;
  float* outs = p.out + (size_t)blockIdx.x * N * DM;

  // ---- weights -> shared memory (rows padded to WS), zero the key padding
  for (int i = tid; i < 3 * DM * (DM / 8); i += AB_WARPS * 32) {
    const int r = i / (DM / 8), c8 = i - r * (DM / 8);
    *reinterpret_cast<uint4*>(Wq + (size_t)r * WS + c8 * 8) = __ldg(reinterpret_cast<const uint4*>(p.Wqkv + (size_t)r * DM + c8 * 8));
  }
  for (int i = tid; i < DM * (DM / 8); i += AB_WARPS * 32) {
    const int r = i / (DM / 8), c8 = i - r * (DM / 8);
    *reinterpret_cast<uint4*>(Wo + (size_t)r * WS + c8 * 8) = __ldg(reinterpret_cast<const uint4*>(p.Wo + (size_t)r * DM + c8 * 8));
  }
  for (int i = tid; i < (Np - N) * WS; i += AB_WARPS * 32) Kb[(size_t)N * WS + i] = __float2bfloat16_rn(0.f);
  for (int i = tid; i < DM * (vts - N); i += AB_WARPS * 32) {
    const int d = i / (vts - N), j = i - d * (vts - N);
    Vt[(size_t)d * vts + N + j] = __float2half_rn(0.f);
  }
  __syncthreads();

  const int tiles = (N + 15) / 16;
  // ---- phase 1: K and V of every token (rows 64..191 of the packed projection) -> shared memory
  for (int rt = warp; rt < tiles; rt += AB_WARPS) {
    const int r0 = rt * 16;
    uint32_t xa[4][4];
    load_x_frags(xs, r0, N, g, t, xa);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      float kc[4], vc[4];
      proj_tile(xa, Wq, DM + nt * 8, p.bqkv, g, t, kc);
      proj_tile(xa, Wq, 2 * DM + nt * 8, p.bqkv, g, t, vc);
      const int col = nt * 8 + 2 * t;
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int r = r0 + g + hh * 8;
        if (r >= N) continue;
        const uint32_t kp = pack_bf16(kc[2 * hh], kc[2 * hh + 1]);
        const __half v0 = __float2half_rn(vc[2 * hh]), v1 = __float2half_rn(vc[2 * hh + 1]);
        // key slot(s) of source row r: identity, or (shifted sequence) j = r-1 and the repeated last row
        int j0 = r, j1 = -1;
        if (p.kv_shift) { j0 = r - 1; j1 = (r == N - 1) ? N - 1 : -1; }
        if (j0 >= 0) {
          *reinterpret_cast<uint32_t*>(Kb + (size_t)j0 * WS + col) = kp;
          Vt[(size_t)col * vts + j0] = v0;
          Vt[(size_t)(col + 1) * vts + j0] = v1;
        }
        if (j1 >= 0) {
          *reinterpret_cast<uint32_t*>(Kb + (size_t)j1 * WS + col) = kp;
          Vt[(size_t)col * vts + j1] = v0;
          Vt[(size_t)(col + 1) * vts + j1] = v1;
        }
      }
    }
  }
  __syncthreads();

  // ---- phase 2: per 16-row tile: Q, attention over all keys per head, out-projection, residual, LayerNorm
  const float qscale = 0.25f * 1.4426950408889634f;          // 1/sqrt(16) * log2(e)
  for (int rt = warp; rt < tiles; rt += AB_WARPS) {
    const int r0 = rt * 16;
    const int row_lo = r0 + g, row_hi = r0 + g + 8;
    uint32_t xa[4][4];
    load_x_frags(xs, r0, N, g, t, xa);
    uint32_t oa[NH][4];
#pragma unroll
    for (int h = 0; h < NH; ++h) {
      // Q of this head (two 8-column tiles), pre-scaled, as an A fragment
      float q0[4], q1[4];
      proj_tile(xa, Wq, h * HD, p.bqkv, g, t, q0);
      proj_tile(xa, Wq, h * HD + 8, p.bqkv, g, t, q1);
#pragma unroll
      for (int i = 0; i < 4; ++i) { q0[i] *= qscale; q1[i] *= qscale; }
      uint32_t qa[4];
      c_to_a(q0, q1, qa);
      float m0 = -INFINITY, m1 = -INFINITY;
      float o[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
      // third output tile = P x [1 0 .. 0]: its column 0 is the row sum of the (rounded) probabilities, so the
      // softmax denominator costs four MMAs per key block instead of 32 adds; B fragment is a lane constant
      float osum[4] = {0.f, 0.f, 0.f, 0.f};
      const uint32_t ones_b = g == 0 ? 0x3C003C00u : 0u;
      for (int kb = 0; kb < Np; kb += KBLK) {
        float s[8][4];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
          const __nv_bfloat16* kr = Kb + (size_t)(kb + j * 8 + g) * WS + h * HD + 2 * t;
          mma_bf16_16816(s[j], qa, *reinterpret_cast<const uint32_t*>(kr), *reinterpret_cast<const uint32_t*>(kr + 8));
        }
        if (kb + KBLK > N) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int c0 = kb + j * 8 + 2 * t;
            if (c0 >= N) s[j][0] = s[j][2] = -INFINITY;
            if (c0 + 1 >= N) s[j][1] = s[j][3] = -INFINITY;
          }
        }
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          mx0 = fmaxf(mx0, fmaxf(s[j][0], s[j][1]));
          mx1 = fmaxf(mx1, fmaxf(s[j][2], s[j][3]));
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
        const float c0 = fast_exp2(m0 - mn0), c1 = fast_exp2(m1 - mn1);
        m0 = mn0; m1 = mn1;
#pragma unroll
        for (int d = 0; d < 2; ++d) { o[d][0] *= c0; o[d][1] *= c0; o[d][2] *= c1; o[d][3] *= c1; }
        osum[0] *= c0; osum[1] *= c0; osum[2] *= c1; osum[3] *= c1;
        // probabilities as packed fp16 pairs: exactly the registers of the A fragment of P (k-step ks = key
        // tiles 2ks, 2ks+1), one MUFU instruction per pair
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          uint32_t pa[4];
          pa[0] = exp2_pair_f16(s[2 * ks][0] - mn0, s[2 * ks][1] - mn0);
          pa[1] = exp2_pair_f16(s[2 * ks][2] - mn1, s[2 * ks][3] - mn1);
          pa[2] = exp2_pair_f16(s[2 * ks + 1][0] - mn0, s[2 * ks + 1][1] - mn0);
          pa[3] = exp2_pair_f16(s[2 * ks + 1][2] - mn1, s[2 * ks + 1][3] - mn1);
#pragma unroll
          for (int d = 0; d < 2; ++d) {
            const __half* vr = Vt + (size_t)(h * HD + d * 8 + g) * vts + kb + ks * 16 + 2 * t;
            mma_f16_16816(o[d], pa, *reinterpret_cast<const uint32_t*>(vr), *reinterpret_cast<const uint32_t*>(vr + 8));
          }
          mma_f16_16816(osum, pa, ones_b, ones_b);
        }
      }
      // column 0 of the sum tile lives in the lanes with t == 0: rows g (osum[0]) and g + 8 (osum[2])
      const float l0 = __shfl_sync(0xffffffffu, osum[0], lane & ~3), l1 = __shfl_sync(0xffffffffu, osum[2], lane & ~3);
      const float i0 = 1.f / l0, i1 = 1.f / l1;
#pragma unroll
      for (int d = 0; d < 2; ++d) { o[d][0] *= i0; o[d][1] *= i0; o[d][2] *= i1; o[d][3] *= i1; }
      c_to_a(o[0], o[1], oa[h]);       // head h = k-step h of the out-projection
    }
    // out-projection + bias + residual + LayerNorm
    float y[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float b0 = __ldg(p.bo + nt * 8 + 2 * t), b1 = __ldg(p.bo + nt * 8 + 2 * t + 1);
      y[nt][0] = b0; y[nt][1] = b1; y[nt][2] = b0; y[nt][3] = b1;
      const __nv_bfloat16* wr = Wo + (size_t)(nt * 8 + g) * WS + 2 * t;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
        mma_bf16_16816(y[nt], oa[ks], *reinterpret_cast<const uint32_t*>(wr + ks * 16),
                       *reinterpret_cast<const uint32_t*>(wr + ks * 16 + 8));
      const int col = nt * 8 + 2 * t;
      if (row_lo < N) {
        const float2 r = __ldg(reinterpret_cast<const float2*>(xs + (size_t)row_lo * DM + col));
        y[nt][0] += r.x; y[nt][1] += r.y;
      }
      if (row_hi < N) {
        const float2 r = __ldg(reinterpret_cast<const float2*>(xs + (size_t)row_hi * DM + col));
        y[nt][2] += r.x; y[nt][3] += r.y;
      }
    }
    layer_norm_tile(y, p.ln_g, p.ln_b, t);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int col = nt * 8 + 2 * t;
      if (row_lo < N) *reinterpret_cast<float2*>(outs + (size_t)row_lo * DM + col) = make_float2(y[nt][0], y[nt][1]);
      if (row_hi < N) *reinterpret_cast<float2*>(outs + (size_t)row_hi * DM + col) = make_float2(y[nt][2], y[nt][3]);
    }
  }
}

// ------------------------------------------------------------------------------------------ FFN + LayerNorm + head
struct FfnHeadArgs {
  const float* x;                      // [M, 64]
  const float* pre;                    // optional [M, 64] added to x before anything else (MSC_N local attention)
  const float* pre_g;                  // optional LayerNorm applied to x + pre (norm2 of the _N variant)
  const float* pre_b;
  const __nv_bfloat16* W1;             // [32][64]
  const float* b1;
  const __nv_bfloat16* W2;             // [64][32]
  const float* b2;
  const float* ln_g;
  const float* ln_b;
  const __nv_bfloat16* Wc;             // [8][64] (rows 3..7 zero)
  const float* bc;                     // [3]
  float* prob;
  float* start;
  float* end;
  int64_t M;
};

constexpr int W2S = FF + 8;

__global__ void __launch_bounds__(256) k_msc_ffn_head(FfnHeadArgs p) {
  __shared__ __align__(16) __nv_bfloat16 W1s[FF * WS];
  __shared__ __align__(16) __nv_bfloat16 W2s[DM * W2S];
  __shared__ __align__(16) __nv_bfloat16 Wcs[8 * WS];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  for (int i = tid; i < FF * (DM / 8); i += 256) {
    const int r = i / (DM / 8), c8 = i - r * (DM / 8);
    *reinterpret_cast<uint4*>(W1s + r * WS + c8 * 8) = __ldg(reinterpret_cast<const uint4*>(p.W1 + r * DM + c8 * 8));
  }
  for (int i = tid; i < DM * (FF / 8); i += 256) {
    const int r = i / (FF / 8), c8 = i - r * (FF / 8);
    *reinterpret_cast<uint4*>(W2s + r * W2S + c8 * 8) = __ldg(reinterpret_cast<const uint4*>(p.W2 + r * FF + c8 * 8));
  }
  for (int i = tid; i < 8 * (DM / 8); i += 256) {
    const int r = i / (DM / 8), c8 = i - r * (DM / 8);
    *reinterpret_cast<uint4*>(Wcs + r * WS + c8 * 8) = __ldg(reinterpret_cast<const uint4*>(p.Wc + r * DM + c8 * 8));
  }
  __syncthreads();

  const int64_t tiles = (p.M + 15) / 16;
  for (int64_t tile = (int64_t)blockIdx.x * 8 + warp; tile < tiles; tile += (int64_t)gridDim.x * 8) {
    const int64_t r0 = tile * 16;
    const int64_t row_lo = r0 + g, row_hi = r0 + g + 8;
    // x tile as fp32 accumulator-layout registers (residual) and as bf16 A fragments
    float x[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int col = nt * 8 + 2 * t;
      float2 a = make_float2(0.f, 0.f), b = a;
      if (row_lo < p.M) a = __ldg(reinterpret_cast<const float2*>(p.x + row_lo * DM + col));
      if (row_hi < p.M) b = __ldg(reinterpret_cast<const float2*>(p.x + row_hi * DM + col));
      if (p.pre) {
        if (row_lo < p.M) { const float2 q = __ldg(reinterpret_cast<const float2*>(p.pre + row_lo * DM + col)); a.x += q.x; a.y += q.y; }
        if (row_hi < p.M) { const float2 q = __ldg(reinterpret_cast<const float2*>(p.pre + row_hi * DM + col)); b.x += q.x; b.y += q.y; }
      }
      x[nt][0] = a.x; x[nt][1] = a.y; x[nt][2] = b.x; x[nt][3] = b.y;
    }
    if (p.pre_g) layer_norm_tile(x, p.pre_g, p.pre_b, t);
    uint32_t xa[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) c_to_a(x[2 * ks], x[2 * ks + 1], xa[ks]);
    // hidden = relu(x W1^T + b1)  [16 x 32]
    float hc[4][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      proj_tile(xa, W1s, nt * 8, p.b1, g, t, hc[nt]);
#pragma unroll
      for (int i = 0; i < 4; ++i) hc[nt][i] = fmaxf(hc[nt][i], 0.f);
    }
    uint32_t ha[2][4];
    c_to_a(hc[0], hc[1], ha[0]);
    c_to_a(hc[2], hc[3], ha[1]);
    // y = hidden W2^T + b2 + x ; LayerNorm
    float y[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float b0 = __ldg(p.b2 + nt * 8 + 2 * t), b1 = __ldg(p.b2 + nt * 8 + 2 * t + 1);
      y[nt][0] = b0 + x[nt][0]; y[nt][1] = b1 + x[nt][1]; y[nt][2] = b0 + x[nt][2]; y[nt][3] = b1 + x[nt][3];
      const __nv_bfloat16* wr = W2s + (size_t)(nt * 8 + g) * W2S + 2 * t;
#pragma unroll
      for (int ks = 0; ks < 2; ++ks)
        mma_bf16_16816(y[nt], ha[ks], *reinterpret_cast<const uint32_t*>(wr + ks * 16),
                       *reinterpret_cast<const uint32_t*>(wr + ks * 16 + 8));
    }
    layer_norm_tile(y, p.ln_g, p.ln_b, t);
    // classifier 64 -> 3 (one 8-column tile, columns 3..7 are zero weights)
    uint32_t ya[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) c_to_a(y[2 * ks], y[2 * ks + 1], ya[ks]);
    float oc[4];
    proj_tile(ya, Wcs, 0, nullptr, g, t, oc);
    // thread t = 0 holds columns 0,1 (probability, start); t = 1 holds column 2 (end)
    if (t == 0) {
      const float bc0 = __ldg(p.bc), bc1 = __ldg(p.bc + 1);
      if (row_lo < p.M) {
        if (p.prob) p.prob[row_lo] = 1.f / (1.f + expf(-(oc[0] + bc0)));
        if (p.start) p.start[row_lo] = tanhf(oc[1] + bc1) * 0.5f + 0.5f;
      }
      if (row_hi < p.M) {
        if (p.prob) p.prob[row_hi] = 1.f / (1.f + expf(-(oc[2] + bc0)));
        if (p.start) p.start[row_hi] = tanhf(oc[3] + bc1) * 0.5f + 0.5f;
      }
    } else if (t == 1) {
      const float bc2 = __ldg(p.bc + 2);
      if (row_lo < p.M && p.end) p.end[row_lo] = tanhf(oc[0] + bc2) * 0.5f + 0.5f;
      if (row_hi < p.M && p.end) p.end[row_hi] = tanhf(oc[2] + bc2) * 0.5f + 0.5f;
    }
  }
}

}  // namespace

bool msc_set_tc_supported(int N, int d, int heads, int ff) { return d == DM && heads == NH && ff == FF && N >= 1 && N <= 320; }

void op_msc_attn_block(Ctx& c, const float* x, const void* Wqkv, const float* bqkv, const void* Wo, const float* bo,
                       const float* ln_g, const float* ln_b, float* out, int64_t B, int N, bool kv_shift) {
  if (c.dry) return;
  AttnBlockArgs p;
  p.x = x; p.Wqkv = static_cast<const __nv_bfloat16*>(Wqkv); p.bqkv = bqkv; p.Wo = static_cast<const __nv_bfloat16*>(Wo);
  p.bo = bo; p.ln_g = ln_g; p.ln_b = ln_b; p.out = out; p.N = N; p.kv_shift = kv_shift ? 1 : 0;
  const int Np = (N + KBLK - 1) / KBLK * KBLK;
  const size_t smem = sizeof(__nv_bfloat16) * ((size_t)Np * WS + (size_t)DM * (Np + 8) + (size_t)4 * DM * WS);
  PAUT_CHECK((int)smem <= c.smem_optin, PAUT_ERR_UNSUPPORTED, "attn block: set too long for shared memory");
  smem_optin(c, k_msc_attn_block);
  PAUT_CHECK(B < (int64_t(1) << 31), PAUT_ERR_INVALID, "attn block: too many sets");
  k_msc_attn_block<<<(unsigned)B, AB_WARPS * 32, smem, c.stream>>>(p);
  c.launched("msc_attn_block");
}

void op_msc_ffn_head(Ctx& c, const float* x, const float* pre, const float* pre_g, const float* pre_b, const void* W1,
                     const float* b1, const void* W2, const float* b2, const float* ln_g, const float* ln_b,
                     const void* Wc, const float* bc, float* prob, float* start, float* end, int64_t M) {
  if (c.dry) return;
  FfnHeadArgs p;
  p.x = x; p.pre = pre; p.pre_g = pre_g; p.pre_b = pre_b; p.W1 = static_cast<const __nv_bfloat16*>(W1); p.b1 = b1;
  p.W2 = static_cast<const __nv_bfloat16*>(W2); p.b2 = b2; p.ln_g = ln_g; p.ln_b = ln_b;
  p.Wc = static_cast<const __nv_bfloat16*>(Wc); p.bc = bc; p.prob = prob; p.start = start; p.end = end; p.M = M;
  const int64_t tiles = (M + 15) / 16;
  int64_t blocks = (tiles + 7) / 8;
  if (blocks > c.num_sms * 8) blocks = c.num_sms * 8;
  k_msc_ffn_head<<<(unsigned)blocks, 256, 0, c.stream>>>(p);
  c.launched("msc_ffn_head");
}

}  // namespace paut
