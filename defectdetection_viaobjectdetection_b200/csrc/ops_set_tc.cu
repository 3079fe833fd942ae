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
#include <algorithm>
#include <cstdlib>
#include <cstring>

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
  // optional fused tail (persistent kernel only): FFN 64 -> 32 -> 64 + residual + LayerNorm + classifier 64 -> 3 on the
  // block's output rows while they are still in registers (k_msc_ffn_head's arithmetic); `out` is then not written
  const __nv_bfloat16* f_W1 = nullptr;   // [32][64]
  const float* f_b1 = nullptr;
  const __nv_bfloat16* f_W2 = nullptr;   // [64][32]
  const float* f_b2 = nullptr;
  const float* f_g = nullptr;
  const float* f_b = nullptr;
  const __nv_bfloat16* f_Wc = nullptr;   // [8][64] (rows 3..7 zero)
  const float* f_bc = nullptr;
  float* f_prob = nullptr;
  float* f_start = nullptr;
  float* f_end = nullptr;
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

// same with the affine parameters loaded as pairs
__device__ __forceinline__ void layer_norm_tile2(float (&y)[8][4], const float* __restrict__ gam,
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
    const float2 gg = __ldg(reinterpret_cast<const float2*>(gam + nt * 8 + 2 * t));
    const float2 bb = __ldg(reinterpret_cast<const float2*>(bet + nt * 8 + 2 * t));
    y[nt][0] = (y[nt][0] - m0) * r0 * gg.x + bb.x;
    y[nt][1] = (y[nt][1] - m0) * r0 * gg.y + bb.y;
    y[nt][2] = (y[nt][2] - m1) * r1 * gg.x + bb.x;
    y[nt][3] = (y[nt][3] - m1) * r1 * gg.y + bb.y;
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

// ------------------------------------------------------------------------------------------ persistent attention block
// Second generation of the attention block (the first, above, stays as the fallback for sets whose K / V do not fit).
// ncu of the first at 300 A-scans per set (profiles/r02/r_ncu_msc_n300_summary.txt): 60 k cycles per set, issue slots
// 56 % busy, special-function pipe 44 %, legacy tensor pipe 33 %; timing variants of a first persistent version
// (profiles/r02/t_attn_variants.log) showed that HALF of the time is outside the softmax loop -- the K / V projection
// with its per-element slot arithmetic for the shifted sequence (1250 SASS instructions per row tile), the
// out-projection / LayerNorm epilogue (800) and their global-load latencies, which all 20 warps of a CTA go through in
// lock step.  This version:
//   * persistent CTA per SM, projection weights staged once; the 20 warps form TWO TEAMS of ten that work on different
//     sets with their own K / V buffer and their own named barrier, so that one team's latency-bound phases (x loads,
//     projections, LayerNorm, stores) overlap the other team's softmax loops;
//   * shifted keys / values (cross block) without any slot arithmetic: softmax attention is invariant under a
//     permutation of the keys, and the shifted sequence [x_1 .. x_{N-1}, x_{N-1}] is the set's own rows with row 0
//     replaced by a second copy of row N-1 -- rows go to their own slot, the owner of row 0 skips its store and the
//     owner of row N-1 also writes slot 0;
//   * V stays row-major like K (plain 32-bit stores of the accumulator fragments) and is read through ldmatrix.trans;
//   * 221 -> ~100 instructions per (head, 64 keys): B fragments through ldmatrix.x4, the running maximum enters the
//     QK^T product as its C operand (no subtraction), fp32 ex2 + one packing conversion per pair (ex2.approx.f16x2
//     compiles to two MUFU and a PRMT), and the reference maximum is only raised when a score exceeds it by more than
//     2^8 (warp vote; probabilities stay below 2^8 in fp16, the denominators are sums of the same rounded values),
//     which takes the rescaling of the partial outputs out of the common path; 32-key blocks keep the scores in 16
//     registers; the last block runs only the 16-key steps that hold real keys.
// 1.52 -> 1.29 ms per 1 M A-scans for the two calls (profiles/r02/u_bench_msc_attn_p.log).  Tried on top of it and dropped,
// all within +-3 % or slower (profiles/r02/t_attn_variants.log): the next block's Q K^T issued before this block's
// exponentials (ping-pong score registers), two heads interleaved per warp, 16 / 48 / 64-key blocks, and a quarter or
// three eighths of the exponentials as a cubic on the FMA pipe -- which made it SLOWER, so the special-function unit
// (50 % busy) is not the limiter, and neither is any other pipe: issue slots 40 %, legacy tensor pipe 41 %.  Removing the
// exponentials, the P V products, the Q K^T products or the vote saves 10 / 7 / 21 / 9 %, and removing the whole softmax
// loop still leaves 45 % of the time in the projections, LayerNorm and barriers at five warps per scheduler.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ldsm4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr)
               : "memory");
}
__device__ __forceinline__ void ldsm4t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr)
               : "memory");
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
// D = A * B + C with C in its own registers
__device__ __forceinline__ void mma_bf16_16816_c(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1,
                                                 const float (&c)[4]) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%11,%12,%13};\n"
      : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "f"(c[0]), "f"(c[1]), "f"(c[2]), "f"(c[3]));
}
__device__ __forceinline__ uint32_t pack_f16(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
// one 8-column output tile of X[16 x 64] * W[rows n0..n0+7]^T + bias; w_addr = this lane's ldmatrix row address of the tile
__device__ __forceinline__ void proj_tile_lm(const uint32_t (&xa)[4][4], uint32_t w_addr, const float* __restrict__ bias,
                                             int n0, int t, float (&c)[4]) {
  const float2 b = __ldg(reinterpret_cast<const float2*>(bias + n0 + 2 * t));
  c[0] = b.x; c[1] = b.y; c[2] = b.x; c[3] = b.y;
  uint32_t w[8];
  ldsm4(w_addr, w[0], w[1], w[2], w[3]);
  ldsm4(w_addr + 64, w[4], w[5], w[6], w[7]);
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) mma_bf16_16816(c, xa[ks], w[2 * ks], w[2 * ks + 1]);
}

constexpr int TEAMS = 2, TW = AB_WARPS / TEAMS;   // two teams of ten warps
constexpr int PB = 2;                  // 16-key steps per softmax block of the persistent kernel
constexpr float LAZY_MAX = 8.f;        // the reference maximum of a row is raised when a score exceeds it by more than this

// VAR != 0: timing experiments only (wrong results): 1 no exponentials, 2 no P V / row-sum products, 3 no Q K^T products,
// 4 no K / V projection, 5 no maximum check after the first block, 6 no softmax loop at all
template <int VAR>
__global__ void __launch_bounds__(AB_WARPS * 32, 1) k_msc_attn_block_p(AttnBlockArgs p, int nsets, int Np, unsigned skew_ns) {
  extern __shared__ __align__(16) unsigned char smraw[];
  const int N = p.N;
  const uint32_t voff = (uint32_t)(Np * WS * 2);                          // V [Np][WS] fp16 follows K [Np][WS] bf16
  const uint32_t kv_bytes = 2 * voff;
  __nv_bfloat16* Wq = reinterpret_cast<__nv_bfloat16*>(smraw);           // [192][WS]
  __nv_bfloat16* Wo = Wq + (size_t)3 * DM * WS;                          // [64][WS]
  __nv_bfloat16* W1f = Wo + (size_t)DM * WS;                            // fused tail: [32][WS], [64][FF + 8], [8][WS]
  __nv_bfloat16* W2f = W1f + (size_t)FF * WS;
  __nv_bfloat16* Wcf = W2f + (size_t)DM * (FF + 8);
  const bool tail = p.f_W1 != nullptr;
  unsigned char* KV = reinterpret_cast<unsigned char*>(tail ? Wcf + (size_t)8 * WS : W1f);   // one K | V buffer per team

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;

  // ---- once per CTA: weights -> shared memory (rows padded to WS), K / V buffers zeroed (the padding keys stay zero)
  for (int i = tid; i < 3 * DM * (DM / 8); i += AB_WARPS * 32) {
    const int r = i / (DM / 8), c8 = i - r * (DM / 8);
    *reinterpret_cast<uint4*>(Wq + (size_t)r * WS + c8 * 8) = __ldg(reinterpret_cast<const uint4*>(p.Wqkv + (size_t)r * DM + c8 * 8));
  }
  for (int i = tid; i < DM * (DM / 8); i += AB_WARPS * 32) {
    const int r = i / (DM / 8), c8 = i - r * (DM / 8);
    *reinterpret_cast<uint4*>(Wo + (size_t)r * WS + c8 * 8) = __ldg(reinterpret_cast<const uint4*>(p.Wo + (size_t)r * DM + c8 * 8));
  }
  for (uint32_t i = tid; i < TEAMS * kv_bytes / 16; i += AB_WARPS * 32) reinterpret_cast<uint4*>(KV)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (tail) {
    for (int i = tid; i < FF * (DM / 8); i += AB_WARPS * 32) {
      const int r = i / (DM / 8), c8 = i - r * (DM / 8);
      *reinterpret_cast<uint4*>(W1f + r * WS + c8 * 8) = __ldg(reinterpret_cast<const uint4*>(p.f_W1 + r * DM + c8 * 8));
    }
    for (int i = tid; i < DM * (FF / 8); i += AB_WARPS * 32) {
      const int r = i / (FF / 8), c8 = i - r * (FF / 8);
      *reinterpret_cast<uint4*>(W2f + r * (FF + 8) + c8 * 8) = __ldg(reinterpret_cast<const uint4*>(p.f_W2 + r * FF + c8 * 8));
    }
    for (int i = tid; i < 8 * (DM / 8); i += AB_WARPS * 32) {
      const int r = i / (DM / 8), c8 = i - r * (DM / 8);
      *reinterpret_cast<uint4*>(Wcf + r * WS + c8 * 8) = __ldg(reinterpret_cast<const uint4*>(p.f_Wc + r * DM + c8 * 8));
    }
  }
  __syncthreads();

  const int team = warp / TW, wt = warp - team * TW;
  const int tiles = (N + 15) / 16;
  // ldmatrix row addresses of this lane: matrix i = lane / 8, row r = lane % 8
  const int lm_i = lane >> 3, lm_r = lane & 7;
  const uint32_t wq_lane = smem_u32(Wq) + (uint32_t)(lm_r * WS * 2 + (lm_i >> 1) * 32 + (lm_i & 1) * 16);
  const uint32_t wo_lane = smem_u32(Wo) + (uint32_t)(lm_r * WS * 2 + (lm_i >> 1) * 32 + (lm_i & 1) * 16);
  const uint32_t kbuf = smem_u32(KV) + (uint32_t)team * kv_bytes;
  // K: matrices (key tile j, dims lo), (j, dims hi), (j+1, lo), (j+1, hi)
  const uint32_t k_lane = kbuf + (uint32_t)(((lm_i >> 1) * 8 + lm_r) * WS * 2 + (lm_i & 1) * 16);
  // V through .trans: (keys lo, dims lo), (keys hi, dims lo), (keys lo, dims hi), (keys hi, dims hi)
  const uint32_t v_lane = kbuf + voff + (uint32_t)(((lm_i & 1) * 8 + lm_r) * WS * 2 + (lm_i >> 1) * 16);
  const float qscale = 0.25f * 1.4426950408889634f;          // 1/sqrt(16) * log2(e)
  const uint32_t ones_b = g == 0 ? 0x3C003C00u : 0u;
  const int bar_id = 1 + team;

  // the teams start half a set apart and, doing the same work at the same speed, stay apart
  if (team == 1 && skew_ns) __nanosleep(skew_ns);

  for (int set = (int)blockIdx.x * TEAMS + team; set < nsets; set += (int)gridDim.x * TEAMS) {
    const float* xs = p.x + (size_t)set * N * DM;
    float* outs = p.out + (size_t)set * N * DM;

    // ---- phase 1: K and V of the team's set (rows 64..191 of the packed projection) -> the team's buffer
    if (VAR != 4) {
      for (int rt = wt; rt < tiles; rt += TW) {
        const int r0 = rt * 16;
        uint32_t xa[4][4];
        load_x_frags(xs, r0, N, g, t, xa);
        const int row_lo = r0 + g, row_hi = row_lo + 8;
        // shifted sequence: row 0 is dropped and row N-1 counts twice (slot 0 takes the second copy)
        const bool p_lo = row_lo < N && !(p.kv_shift && row_lo == 0), p_hi = row_hi < N;
        const bool d_lo = p.kv_shift && row_lo == N - 1, d_hi = p.kv_shift && row_hi == N - 1;
        const bool has_dup = p.kv_shift && N - 1 >= r0 && N - 1 < r0 + 16;     // warp-uniform
        const uint32_t st = kbuf + (uint32_t)(row_lo * WS * 2 + 4 * t), st0 = kbuf + (uint32_t)(4 * t);
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          float kc[4], vc[4];
          proj_tile_lm(xa, wq_lane + (uint32_t)((DM + nt * 8) * WS * 2), p.bqkv, DM + nt * 8, t, kc);
          proj_tile_lm(xa, wq_lane + (uint32_t)((2 * DM + nt * 8) * WS * 2), p.bqkv, 2 * DM + nt * 8, t, vc);
          const uint32_t k_lo = pack_bf16(kc[0], kc[1]), k_hi = pack_bf16(kc[2], kc[3]);
          const uint32_t v_lo = pack_f16(vc[0], vc[1]), v_hi = pack_f16(vc[2], vc[3]);
          if (p_lo) { sts32(st + nt * 16, k_lo); sts32(st + voff + nt * 16, v_lo); }
          if (p_hi) { sts32(st + 8 * WS * 2 + nt * 16, k_hi); sts32(st + voff + 8 * WS * 2 + nt * 16, v_hi); }
          if (has_dup) {
            if (d_lo) { sts32(st0 + nt * 16, k_lo); sts32(st0 + voff + nt * 16, v_lo); }
            if (d_hi) { sts32(st0 + nt * 16, k_hi); sts32(st0 + voff + nt * 16, v_hi); }
          }
        }
      }
    }
    asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(TW * 32) : "memory");

    // ---- phase 2: per 16-row tile: Q, attention over all keys per head, out-projection, residual, LayerNorm
    for (int rt = wt; rt < tiles; rt += TW) {
      const int r0 = rt * 16;
      const int row_lo = r0 + g, row_hi = row_lo + 8;
      // Q of all four heads first (pre-scaled A fragments): x is dead afterwards, and head h's slot of `qo` is reused
      // for its normalised output (= k-step h of the out-projection), so the two never add up
      uint32_t qo[NH][4];
      {
        uint32_t xa[4][4];
        load_x_frags(xs, r0, N, g, t, xa);
#pragma unroll
        for (int h = 0; h < NH; ++h) {
          float q0[4], q1[4];
          proj_tile_lm(xa, wq_lane + (uint32_t)((h * HD) * WS * 2), p.bqkv, h * HD, t, q0);
          proj_tile_lm(xa, wq_lane + (uint32_t)((h * HD + 8) * WS * 2), p.bqkv, h * HD + 8, t, q1);
#pragma unroll
          for (int i = 0; i < 4; ++i) { q0[i] *= qscale; q1[i] *= qscale; }
          c_to_a(q0, q1, qo[h]);
        }
      }
#pragma unroll
      for (int h = 0; h < NH; ++h) {
        uint32_t qa[4] = {qo[h][0], qo[h][1], qo[h][2], qo[h][3]};
        float negm[4] = {0.f, 0.f, 0.f, 0.f};                 // minus the reference maximum of rows g (0, 1) and g + 8 (2, 3)
        float o[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
        float osum[4] = {0.f, 0.f, 0.f, 0.f};                 // P x [1 0 .. 0]: column 0 = row sums of the rounded probabilities
        const uint32_t ka = k_lane + (uint32_t)(h * HD * 2), va = v_lane + (uint32_t)(h * HD * 2);
        // scores of one block of `nks` <= PB 16-key steps starting at key kb (reference maximum = C operand)
        auto scores = [&](float (&s)[2 * PB][4], int kb, int nks, bool last) {
#pragma unroll
          for (int jp = 0; jp < PB; ++jp) {
            if (last && jp >= nks) break;
            if (VAR == 3) {
#pragma unroll
              for (int i = 0; i < 4; ++i) { s[2 * jp][i] = negm[i] + __uint_as_float(qa[i]) * 1e-30f; s[2 * jp + 1][i] = negm[i]; }
              continue;
            }
            uint32_t k0, k1, k2, k3;
            ldsm4(ka + (uint32_t)((kb + jp * 16) * WS * 2), k0, k1, k2, k3);
            mma_bf16_16816_c(s[2 * jp], qa, k0, k1, negm);
            mma_bf16_16816_c(s[2 * jp + 1], qa, k2, k3, negm);
          }
          if (last) {
#pragma unroll
            for (int j = 0; j < 2 * PB; ++j) {
              if (j >= 2 * nks) break;
              const int c0 = kb + j * 8 + 2 * t;
              if (c0 >= N) s[j][0] = s[j][2] = -INFINITY;
              if (c0 + 1 >= N) s[j][1] = s[j][3] = -INFINITY;
            }
          }
        };
        // softmax numerators and P V of a block whose scores are in s.  Blocks are short (32 keys): without a rescale per
        // block their only overhead is the vote, and 16 score registers keep the loop free of spills
        auto consume = [&](float (&s)[2 * PB][4], int kb, int nks, bool last) {
          float lm = -INFINITY;
#pragma unroll
          for (int j = 0; j < 2 * PB; ++j) {
            if (last && j >= 2 * nks) break;
            lm = fmaxf(lm, fmaxf(fmaxf(s[j][0], s[j][1]), fmaxf(s[j][2], s[j][3])));
          }
          // does any score exceed its row's reference by more than 2^LAZY_MAX (always true for the first block, whose
          // reference is still zero)?
          const bool first = kb == 0;
          if (first || (VAR != 5 && __any_sync(0xffffffffu, lm > LAZY_MAX))) {
            float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
            for (int j = 0; j < 2 * PB; ++j) {
              if (last && j >= 2 * nks) break;
              mx0 = fmaxf(mx0, fmaxf(s[j][0], s[j][1]));
              mx1 = fmaxf(mx1, fmaxf(s[j][2], s[j][3]));
            }
            mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
            mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
            mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
            mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
            // first block: the reference becomes the block maximum (nothing accumulated yet); later: raise, never lower
            const float d0 = first ? mx0 : fmaxf(mx0, 0.f), d1 = first ? mx1 : fmaxf(mx1, 0.f);
            const float c0 = first ? 1.f : fast_exp2(-d0), c1 = first ? 1.f : fast_exp2(-d1);
            negm[0] -= d0; negm[1] -= d0; negm[2] -= d1; negm[3] -= d1;
#pragma unroll
            for (int d = 0; d < 2; ++d) { o[d][0] *= c0; o[d][1] *= c0; o[d][2] *= c1; o[d][3] *= c1; }
            osum[0] *= c0; osum[1] *= c0; osum[2] *= c1; osum[3] *= c1;
#pragma unroll
            for (int j = 0; j < 2 * PB; ++j) {
              if (last && j >= 2 * nks) break;
              s[j][0] -= d0; s[j][1] -= d0; s[j][2] -= d1; s[j][3] -= d1;
            }
          }
#pragma unroll
          for (int ks = 0; ks < PB; ++ks) {
            if (last && ks >= nks) break;
            uint32_t pa[4];
            if (VAR == 1) {
              pa[0] = pack_f16(s[2 * ks][0], s[2 * ks][1]);
              pa[1] = pack_f16(s[2 * ks][2], s[2 * ks][3]);
              pa[2] = pack_f16(s[2 * ks + 1][0], s[2 * ks + 1][1]);
              pa[3] = pack_f16(s[2 * ks + 1][2], s[2 * ks + 1][3]);
            } else {
              pa[0] = pack_f16(fast_exp2(s[2 * ks][0]), fast_exp2(s[2 * ks][1]));
              pa[1] = pack_f16(fast_exp2(s[2 * ks][2]), fast_exp2(s[2 * ks][3]));
              pa[2] = pack_f16(fast_exp2(s[2 * ks + 1][0]), fast_exp2(s[2 * ks + 1][1]));
              pa[3] = pack_f16(fast_exp2(s[2 * ks + 1][2]), fast_exp2(s[2 * ks + 1][3]));
            }
            if (VAR == 2) {
#pragma unroll
              for (int i = 0; i < 4; ++i) o[0][i] += __uint_as_float(pa[i]);
              continue;
            }
            uint32_t v0, v1, v2, v3;
            ldsm4t(va + (uint32_t)((kb + ks * 16) * WS * 2), v0, v1, v2, v3);
            mma_f16_16816(o[0], pa, v0, v1);
            mma_f16_16816(o[1], pa, v2, v3);
            mma_f16_16816(osum, pa, ones_b, ones_b);
          }
        };
        if (VAR != 6) {
          const int nfull = N / (PB * 16), tail_ks = (N - nfull * PB * 16 + 15) >> 4;      // full blocks, 16-key steps of the tail
          float sa[2 * PB][4];
          for (int b = 0; b < nfull; ++b) {
            scores(sa, b * PB * 16, PB, false);
            consume(sa, b * PB * 16, PB, false);
          }
          if (tail_ks) {
            scores(sa, nfull * PB * 16, tail_ks, true);
            consume(sa, nfull * PB * 16, tail_ks, true);
          }
        } else {
          o[0][0] = __uint_as_float(qa[0]); osum[0] = osum[2] = 1.f;
        }
        // column 0 of the sum tile lives in the lanes with t == 0: rows g (osum[0]) and g + 8 (osum[2])
        const float l0 = __shfl_sync(0xffffffffu, osum[0], lane & ~3), l1 = __shfl_sync(0xffffffffu, osum[2], lane & ~3);
        const float i0 = 1.f / l0, i1 = 1.f / l1;
#pragma unroll
        for (int d = 0; d < 2; ++d) { o[d][0] *= i0; o[d][1] *= i0; o[d][2] *= i1; o[d][3] *= i1; }
        c_to_a(o[0], o[1], qo[h]);       // head h = k-step h of the out-projection
      }
      // out-projection + bias + residual + LayerNorm
      const bool p_lo = row_lo < N, p_hi = row_hi < N;
      const float* x_lo = xs + (size_t)(p_lo ? row_lo : 0) * DM + 2 * t;     // rows beyond the set read row 0 and are not stored
      const float* x_hi = xs + (size_t)(p_hi ? row_hi : 0) * DM + 2 * t;
      float y[8][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const float2 bo = __ldg(reinterpret_cast<const float2*>(p.bo + nt * 8 + 2 * t));
        const float2 rl = __ldg(reinterpret_cast<const float2*>(x_lo + nt * 8));
        const float2 rh = __ldg(reinterpret_cast<const float2*>(x_hi + nt * 8));
        y[nt][0] = bo.x + rl.x; y[nt][1] = bo.y + rl.y; y[nt][2] = bo.x + rh.x; y[nt][3] = bo.y + rh.y;
        uint32_t w[8];
        ldsm4(wo_lane + (uint32_t)(nt * 8 * WS * 2), w[0], w[1], w[2], w[3]);
        ldsm4(wo_lane + (uint32_t)(nt * 8 * WS * 2) + 64, w[4], w[5], w[6], w[7]);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) mma_bf16_16816(y[nt], qo[ks], w[2 * ks], w[2 * ks + 1]);
      }
      layer_norm_tile2(y, p.ln_g, p.ln_b, t);
      if (tail) {
        // FFN 64 -> 32 -> 64, residual, LayerNorm, classifier 64 -> 3, sigmoid / tanh (NN_models.py:39-41, :123-127) on the
        // rows in registers: what k_msc_ffn_head does in a launch of its own, without the [B, N, 64] round trip
        uint32_t xa[4][4];
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) c_to_a(y[2 * ks], y[2 * ks + 1], xa[ks]);
        float hc[4][4];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          proj_tile(xa, W1f, nt * 8, p.f_b1, g, t, hc[nt]);
#pragma unroll
          for (int i = 0; i < 4; ++i) hc[nt][i] = fmaxf(hc[nt][i], 0.f);
        }
        uint32_t ha[2][4];
        c_to_a(hc[0], hc[1], ha[0]);
        c_to_a(hc[2], hc[3], ha[1]);
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          const float b0 = __ldg(p.f_b2 + nt * 8 + 2 * t), b1 = __ldg(p.f_b2 + nt * 8 + 2 * t + 1);
          y[nt][0] += b0; y[nt][1] += b1; y[nt][2] += b0; y[nt][3] += b1;           // residual = the block's output itself
          const __nv_bfloat16* wr = W2f + (size_t)(nt * 8 + g) * (FF + 8) + 2 * t;
#pragma unroll
          for (int ks = 0; ks < 2; ++ks)
            mma_bf16_16816(y[nt], ha[ks], *reinterpret_cast<const uint32_t*>(wr + ks * 16),
                           *reinterpret_cast<const uint32_t*>(wr + ks * 16 + 8));
        }
        layer_norm_tile2(y, p.f_g, p.f_b, t);
        uint32_t ya[4][4];
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) c_to_a(y[2 * ks], y[2 * ks + 1], ya[ks]);
        float oc[4];
        proj_tile(ya, Wcf, 0, nullptr, g, t, oc);
        const size_t base = (size_t)set * N;
        // thread t = 0 holds columns 0, 1 (probability, start); t = 1 holds column 2 (end)
        if (t == 0) {
          const float bc0 = __ldg(p.f_bc), bc1 = __ldg(p.f_bc + 1);
          if (p_lo) {
            if (p.f_prob) p.f_prob[base + row_lo] = 1.f / (1.f + expf(-(oc[0] + bc0)));
            if (p.f_start) p.f_start[base + row_lo] = tanhf(oc[1] + bc1) * 0.5f + 0.5f;
          }
          if (p_hi) {
            if (p.f_prob) p.f_prob[base + row_hi] = 1.f / (1.f + expf(-(oc[2] + bc0)));
            if (p.f_start) p.f_start[base + row_hi] = tanhf(oc[3] + bc1) * 0.5f + 0.5f;
          }
        } else if (t == 1) {
          const float bc2 = __ldg(p.f_bc + 2);
          if (p_lo && p.f_end) p.f_end[base + row_lo] = tanhf(oc[0] + bc2) * 0.5f + 0.5f;
          if (p_hi && p.f_end) p.f_end[base + row_hi] = tanhf(oc[2] + bc2) * 0.5f + 0.5f;
        }
      } else {
        float* o_lo = outs + (size_t)row_lo * DM + 2 * t;
        float* o_hi = outs + (size_t)row_hi * DM + 2 * t;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          if (p_lo) *reinterpret_cast<float2*>(o_lo + nt * 8) = make_float2(y[nt][0], y[nt][1]);
          if (p_hi) *reinterpret_cast<float2*>(o_hi + nt * 8) = make_float2(y[nt][2], y[nt][3]);
        }
      }
    }
    // the team's next set overwrites the buffer: every warp of the team must be out of its softmax loops
    asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(TW * 32) : "memory");
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

// ------------------------------------------------------------------------------------------ two-stage heads
// The four heads of TwoStageDefectDetector (defect classifier, its uncertainty, position predictor, its uncertainty:
// two_stage_model.py:160-251) are each  Linear 128 -> 64, LayerNorm(64), ReLU, Linear 64 -> 2, activation  on the
// transformer output after its final LayerNorm(128).  As separate launches that is 13 kernels and ~4.5 ms per 1 M A-scans
// (four GEMMs that each re-read the 512-byte rows, four LayerNorms, four 64 -> 2 row kernels, the final norm); here a
// warp takes 16 rows, normalises them in registers and runs the four heads on mma.sync fragments: the rows are read
// once and 8 floats per head row are written.
constexpr int HW = 128, HH = 64, HS = HW + 8;      // head input width, hidden width, bf16 row stride of the staged weights

struct TsHeadsArgs {
  const float* x;                      // [M, 128] transformer output before sequence_transformer.norm
  const float* ng;                     // final norm
  const float* nb;
  const __nv_bfloat16* W0[4];          // [64][128]
  const float* b0[4];
  const float* lg[4];                  // LayerNorm(64)
  const float* lb[4];
  const float* W4[4];                  // [2][64] fp32
  const float* b4[4];
  int act[4];
  float eps[4];
  float* out[4];                       // [M, 2] (null: head not wanted)
  int64_t M;
};

__device__ __forceinline__ float head_act(float v, int act) {
  if (act == ACT_SIGMOID) return 1.f / (1.f + expf(-v));
  if (act == ACT_SOFTPLUS) return v > 20.f ? v : log1pf(expf(v));
  return v;
}

__global__ void __launch_bounds__(256) k_ts_heads(TsHeadsArgs p) {
  extern __shared__ __align__(16) unsigned char hsm[];
  __nv_bfloat16* Ws = reinterpret_cast<__nv_bfloat16*>(hsm);            // [4][64][HS]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  for (int i = tid; i < 4 * HH * (HW / 8); i += 256) {
    const int h = i / (HH * (HW / 8)), r = (i / (HW / 8)) % HH, c8 = i % (HW / 8);
    if (p.out[h])
      *reinterpret_cast<uint4*>(Ws + ((size_t)h * HH + r) * HS + c8 * 8) = __ldg(reinterpret_cast<const uint4*>(p.W0[h] + (size_t)r * HW + c8 * 8));
  }
  __syncthreads();
  const int lm_i = lane >> 3, lm_r = lane & 7;
  const uint32_t w_lane = smem_u32(Ws) + (uint32_t)(lm_r * HS * 2 + (lm_i >> 1) * 32 + (lm_i & 1) * 16);

  const int64_t tiles = (p.M + 15) / 16;
  for (int64_t tile = (int64_t)blockIdx.x * 8 + warp; tile < tiles; tile += (int64_t)gridDim.x * 8) {
    const int64_t row_lo = tile * 16 + g, row_hi = row_lo + 8;
    const bool p_lo = row_lo < p.M, p_hi = row_hi < p.M;
    const float* x_lo = p.x + (p_lo ? row_lo : 0) * HW + 2 * t;
    const float* x_hi = p.x + (p_hi ? row_hi : 0) * HW + 2 * t;
    // ---- the 16 x 128 tile in accumulator layout, final LayerNorm(128) in registers
    float x[16][4];
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < 16; ++nt) {
      const float2 a = __ldg(reinterpret_cast<const float2*>(x_lo + nt * 8));
      const float2 b = __ldg(reinterpret_cast<const float2*>(x_hi + nt * 8));
      x[nt][0] = a.x; x[nt][1] = a.y; x[nt][2] = b.x; x[nt][3] = b.y;
      s0 += a.x + a.y; s1 += b.x + b.y;
    }
    s0 += __shfl_xor_sync(0xffffffffu, s0, 1); s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
    s1 += __shfl_xor_sync(0xffffffffu, s1, 1); s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
    const float m0 = s0 * (1.f / HW), m1 = s1 * (1.f / HW);
    float q0 = 0.f, q1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < 16; ++nt) {
      const float a = x[nt][0] - m0, b = x[nt][1] - m0, c = x[nt][2] - m1, d = x[nt][3] - m1;
      q0 += a * a + b * b; q1 += c * c + d * d;
    }
    q0 += __shfl_xor_sync(0xffffffffu, q0, 1); q0 += __shfl_xor_sync(0xffffffffu, q0, 2);
    q1 += __shfl_xor_sync(0xffffffffu, q1, 1); q1 += __shfl_xor_sync(0xffffffffu, q1, 2);
    const float r0 = rsqrtf(q0 * (1.f / HW) + 1e-5f), r1 = rsqrtf(q1 * (1.f / HW) + 1e-5f);
    uint32_t xa[8][4];                                     // A fragments of the normalised tile, 8 k-steps
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
      float v[2][4];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int nt = 2 * ks + j;
        const float2 gg = __ldg(reinterpret_cast<const float2*>(p.ng + nt * 8 + 2 * t));
        const float2 bb = __ldg(reinterpret_cast<const float2*>(p.nb + nt * 8 + 2 * t));
        v[j][0] = (x[nt][0] - m0) * r0 * gg.x + bb.x; v[j][1] = (x[nt][1] - m0) * r0 * gg.y + bb.y;
        v[j][2] = (x[nt][2] - m1) * r1 * gg.x + bb.x; v[j][3] = (x[nt][3] - m1) * r1 * gg.y + bb.y;
      }
      c_to_a(v[0], v[1], xa[ks]);
    }
    // ---- the heads
#pragma unroll 1
    for (int h = 0; h < 4; ++h) {
      if (!p.out[h]) continue;
      float hc[8][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const float2 b = __ldg(reinterpret_cast<const float2*>(p.b0[h] + nt * 8 + 2 * t));
        hc[nt][0] = b.x; hc[nt][1] = b.y; hc[nt][2] = b.x; hc[nt][3] = b.y;
        const uint32_t wa = w_lane + (uint32_t)((h * HH + nt * 8) * HS * 2);
#pragma unroll
        for (int kq = 0; kq < 4; ++kq) {                   // four k-steps pairs: ldmatrix.x4 = two k-steps of one n-tile
          uint32_t w0, w1, w2, w3;
          ldsm4(wa + kq * 64, w0, w1, w2, w3);
          mma_bf16_16816(hc[nt], xa[2 * kq], w0, w1);
          mma_bf16_16816(hc[nt], xa[2 * kq + 1], w2, w3);
        }
      }
      layer_norm_tile2(hc, p.lg[h], p.lb[h], t);
      // ReLU, then Linear 64 -> 2 as per-lane partial dot products reduced over the quad
      float o00 = 0.f, o01 = 0.f, o10 = 0.f, o11 = 0.f;     // [row lo / hi][output 0 / 1]
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const float2 wA = __ldg(reinterpret_cast<const float2*>(p.W4[h] + nt * 8 + 2 * t));
        const float2 wB = __ldg(reinterpret_cast<const float2*>(p.W4[h] + HH + nt * 8 + 2 * t));
        const float a = fmaxf(hc[nt][0], 0.f), b = fmaxf(hc[nt][1], 0.f), c = fmaxf(hc[nt][2], 0.f), d = fmaxf(hc[nt][3], 0.f);
        o00 += a * wA.x + b * wA.y; o01 += a * wB.x + b * wB.y;
        o10 += c * wA.x + d * wA.y; o11 += c * wB.x + d * wB.y;
      }
      o00 += __shfl_xor_sync(0xffffffffu, o00, 1); o00 += __shfl_xor_sync(0xffffffffu, o00, 2);
      o01 += __shfl_xor_sync(0xffffffffu, o01, 1); o01 += __shfl_xor_sync(0xffffffffu, o01, 2);
      o10 += __shfl_xor_sync(0xffffffffu, o10, 1); o10 += __shfl_xor_sync(0xffffffffu, o10, 2);
      o11 += __shfl_xor_sync(0xffffffffu, o11, 1); o11 += __shfl_xor_sync(0xffffffffu, o11, 2);
      if (t == 0) {
        const float bA = __ldg(p.b4[h]), bB = __ldg(p.b4[h] + 1);
        const int act = p.act[h];
        const float eps = p.eps[h];
        if (p_lo) *reinterpret_cast<float2*>(p.out[h] + row_lo * 2) = make_float2(head_act(o00 + bA, act) + eps, head_act(o01 + bB, act) + eps);
        if (p_hi) *reinterpret_cast<float2*>(p.out[h] + row_hi * 2) = make_float2(head_act(o10 + bA, act) + eps, head_act(o11 + bB, act) + eps);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------ Linear + residual + LayerNorm
// out = LayerNorm(x W^T + b + res) for a 128-wide output: the attention out-projection and the second FFN layer of a
// post-norm encoder layer with d_model 128 (two-stage model, SignalSequenceDetector).  As a tcgen05 GEMM followed by a
// LayerNorm launch the sum made a round trip through HBM (512 B written + read per row) and the GEMM's epilogue was its
// slowest part; here a warp owns 16 rows: A fragments straight from the fp32 activations, W (bf16, rows padded by 8)
// resident in shared memory and read through ldmatrix, the 16 x 128 sum and its LayerNorm in registers.
constexpr int LRL_N = 128;

struct LinResLnArgs {
  const float* x;                      // [M, lda] fp32
  int lda, K;                          // K % 32 == 0
  const __nv_bfloat16* W;              // [128][K]
  const float* bias;
  const float* res;                    // [M, 128] (may be null)
  const float* g;
  const float* b;
  const float* table;                  // optional [table_mod, 128] added AFTER the LayerNorm to row (m % table_mod): positional encoding
  int table_mod;
  float* out;                          // [M, 128]
  int64_t M;
};

// THREADS = 256 with two CTAs per SM (weights <= ~100 KB) or 512 with one: 16 warps per SM either way at 128 registers
template <int THREADS>
__global__ void __launch_bounds__(THREADS, THREADS == 256 ? 2 : 1) k_lin_res_ln(LinResLnArgs p) {
  extern __shared__ __align__(16) unsigned char lsm[];
  __nv_bfloat16* Ws = reinterpret_cast<__nv_bfloat16*>(lsm);            // [128][K + 8]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int K = p.K, KS = K + 8;
  for (int i = tid; i < LRL_N * (K / 8); i += THREADS) {
    const int r = i / (K / 8), c8 = i - r * (K / 8);
    *reinterpret_cast<uint4*>(Ws + (size_t)r * KS + c8 * 8) = __ldg(reinterpret_cast<const uint4*>(p.W + (size_t)r * K + c8 * 8));
  }
  __syncthreads();
  const int lm_i = lane >> 3, lm_r = lane & 7;
  const uint32_t w_lane = smem_u32(Ws) + (uint32_t)(lm_r * KS * 2 + (lm_i >> 1) * 32 + (lm_i & 1) * 16);
  const uint32_t nt_stride = (uint32_t)(8 * KS * 2);

  const int64_t tiles = (p.M + 15) / 16;
  for (int64_t tile = (int64_t)blockIdx.x * (THREADS / 32) + warp; tile < tiles; tile += (int64_t)gridDim.x * (THREADS / 32)) {
    const int64_t row_lo = tile * 16 + g, row_hi = row_lo + 8;
    const bool p_lo = row_lo < p.M, p_hi = row_hi < p.M;
    const float* x_lo = p.x + (p_lo ? row_lo : 0) * p.lda + 2 * t;
    const float* x_hi = p.x + (p_hi ? row_hi : 0) * p.lda + 2 * t;
    float y[16][4];
    if (p.res) {
      const float* r_lo = p.res + (p_lo ? row_lo : 0) * LRL_N + 2 * t;
      const float* r_hi = p.res + (p_hi ? row_hi : 0) * LRL_N + 2 * t;
#pragma unroll
      for (int nt = 0; nt < 16; ++nt) {
        const float2 bb = __ldg(reinterpret_cast<const float2*>(p.bias + nt * 8 + 2 * t));
        const float2 a = __ldg(reinterpret_cast<const float2*>(r_lo + nt * 8));
        const float2 b = __ldg(reinterpret_cast<const float2*>(r_hi + nt * 8));
        y[nt][0] = bb.x + a.x; y[nt][1] = bb.y + a.y; y[nt][2] = bb.x + b.x; y[nt][3] = bb.y + b.y;
      }
    } else {
#pragma unroll
      for (int nt = 0; nt < 16; ++nt) {
        const float2 bb = __ldg(reinterpret_cast<const float2*>(p.bias + nt * 8 + 2 * t));
        y[nt][0] = bb.x; y[nt][1] = bb.y; y[nt][2] = bb.x; y[nt][3] = bb.y;
      }
    }
    // K in steps of 32 (two MMA k-steps = one ldmatrix.x4 per 8 output columns); the activations of the next step are
    // requested before this step's products
    auto load_a = [&](int kq, float2 (&v)[8]) {
      const int k = kq * 32;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        v[4 * j + 0] = __ldg(reinterpret_cast<const float2*>(x_lo + k + 16 * j));
        v[4 * j + 1] = __ldg(reinterpret_cast<const float2*>(x_hi + k + 16 * j));
        v[4 * j + 2] = __ldg(reinterpret_cast<const float2*>(x_lo + k + 16 * j + 8));
        v[4 * j + 3] = __ldg(reinterpret_cast<const float2*>(x_hi + k + 16 * j + 8));
      }
    };
    const int nkq = K / 32;
    float2 cur[8], nxt[8];
    load_a(0, cur);
    for (int kq = 0; kq < nkq; ++kq) {
      if (kq + 1 < nkq) load_a(kq + 1, nxt);
      uint32_t a0[4], a1[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a0[i] = pack_bf16(cur[i].x, cur[i].y); a1[i] = pack_bf16(cur[4 + i].x, cur[4 + i].y); }
      const uint32_t wk = w_lane + (uint32_t)kq * 64;
#pragma unroll
      for (int nt = 0; nt < 16; ++nt) {
        uint32_t w0, w1, w2, w3;
        ldsm4(wk + (uint32_t)nt * nt_stride, w0, w1, w2, w3);
        mma_bf16_16816(y[nt], a0, w0, w1);
        mma_bf16_16816(y[nt], a1, w2, w3);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) cur[i] = nxt[i];
    }
    // LayerNorm(128) on the accumulator fragments
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < 16; ++nt) { s0 += y[nt][0] + y[nt][1]; s1 += y[nt][2] + y[nt][3]; }
    s0 += __shfl_xor_sync(0xffffffffu, s0, 1); s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
    s1 += __shfl_xor_sync(0xffffffffu, s1, 1); s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
    const float m0 = s0 * (1.f / LRL_N), m1 = s1 * (1.f / LRL_N);
    float q0 = 0.f, q1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < 16; ++nt) {
      const float a = y[nt][0] - m0, b = y[nt][1] - m0, c = y[nt][2] - m1, d = y[nt][3] - m1;
      q0 += a * a + b * b; q1 += c * c + d * d;
    }
    q0 += __shfl_xor_sync(0xffffffffu, q0, 1); q0 += __shfl_xor_sync(0xffffffffu, q0, 2);
    q1 += __shfl_xor_sync(0xffffffffu, q1, 1); q1 += __shfl_xor_sync(0xffffffffu, q1, 2);
    const float r0 = rsqrtf(q0 * (1.f / LRL_N) + 1e-5f), r1 = rsqrtf(q1 * (1.f / LRL_N) + 1e-5f);
    float* o_lo = p.out + row_lo * LRL_N + 2 * t;
    float* o_hi = p.out + row_hi * LRL_N + 2 * t;
    const float* t_lo = p.table ? p.table + ((p_lo ? row_lo : 0) % p.table_mod) * LRL_N + 2 * t : nullptr;
    const float* t_hi = p.table ? p.table + ((p_hi ? row_hi : 0) % p.table_mod) * LRL_N + 2 * t : nullptr;
#pragma unroll
    for (int nt = 0; nt < 16; ++nt) {
      const float2 gg = __ldg(reinterpret_cast<const float2*>(p.g + nt * 8 + 2 * t));
      const float2 bb = __ldg(reinterpret_cast<const float2*>(p.b + nt * 8 + 2 * t));
      float2 ta = make_float2(0.f, 0.f), tb = ta;
      if (p.table) { ta = __ldg(reinterpret_cast<const float2*>(t_lo + nt * 8)); tb = __ldg(reinterpret_cast<const float2*>(t_hi + nt * 8)); }
      if (p_lo) *reinterpret_cast<float2*>(o_lo + nt * 8) = make_float2((y[nt][0] - m0) * r0 * gg.x + bb.x + ta.x, (y[nt][1] - m0) * r0 * gg.y + bb.y + ta.y);
      if (p_hi) *reinterpret_cast<float2*>(o_hi + nt * 8) = make_float2((y[nt][2] - m1) * r1 * gg.x + bb.x + tb.x, (y[nt][3] - m1) * r1 * gg.y + bb.y + tb.y);
    }
  }
}

}  // namespace

bool msc_set_tc_supported(int N, int d, int heads, int ff) { return d == DM && heads == NH && ff == FF && N >= 1 && N <= 320; }

constexpr size_t TAIL_BYTES = sizeof(__nv_bfloat16) * ((size_t)FF * WS + (size_t)DM * (FF + 8) + (size_t)8 * WS);

// can the FFN + head tail ride in the attention block's epilogue for sets of N tokens (persistent kernel, shared memory)?
bool msc_attn_tail_supported(const Ctx& c, int N) {
  static const bool off = [] { const char* e = std::getenv("PAUT_ATTN"); return e && (std::strcmp(e, "v1") == 0 || std::strcmp(e, "notail") == 0); }();
  const int Np16 = (N + 15) / 16 * 16;
  const size_t smem_p = sizeof(__nv_bfloat16) * ((size_t)TEAMS * 2 * Np16 * WS + (size_t)4 * DM * WS) + TAIL_BYTES;
  return !off && (int)smem_p <= c.smem_optin - 64 && (N + 15) / 16 <= 2 * TW;
}

void op_msc_attn_block(Ctx& c, const float* x, const void* Wqkv, const float* bqkv, const void* Wo, const float* bo,
                       const float* ln_g, const float* ln_b, float* out, int64_t B, int N, bool kv_shift, const MscTail* tail) {
  if (c.dry) return;
  AttnBlockArgs p;
  p.x = x; p.Wqkv = static_cast<const __nv_bfloat16*>(Wqkv); p.bqkv = bqkv; p.Wo = static_cast<const __nv_bfloat16*>(Wo);
  p.bo = bo; p.ln_g = ln_g; p.ln_b = ln_b; p.out = out; p.N = N; p.kv_shift = kv_shift ? 1 : 0;
  if (tail) {
    PAUT_CHECK(msc_attn_tail_supported(c, N), PAUT_ERR_UNSUPPORTED, "attn block: the fused tail does not fit for this set length");
    p.f_W1 = static_cast<const __nv_bfloat16*>(tail->W1); p.f_b1 = tail->b1; p.f_W2 = static_cast<const __nv_bfloat16*>(tail->W2);
    p.f_b2 = tail->b2; p.f_g = tail->ln_g; p.f_b = tail->ln_b; p.f_Wc = static_cast<const __nv_bfloat16*>(tail->Wc); p.f_bc = tail->bc;
    p.f_prob = tail->prob; p.f_start = tail->start; p.f_end = tail->end;
  }
  const int Np = (N + KBLK - 1) / KBLK * KBLK;
  const size_t smem = sizeof(__nv_bfloat16) * ((size_t)Np * WS + (size_t)DM * (Np + 8) + (size_t)4 * DM * WS);
  PAUT_CHECK((int)smem <= c.smem_optin, PAUT_ERR_UNSUPPORTED, "attn block: set too long for shared memory");
  PAUT_CHECK(B < (int64_t(1) << 31), PAUT_ERR_INVALID, "attn block: too many sets");
  // persistent kernel: one CTA per SM, two teams of ten warps with a K / V buffer each (PAUT_ATTN=v1 keeps the
  // one-CTA-per-set kernel for A/B runs)
  static const bool force_v1 = [] { const char* e = std::getenv("PAUT_ATTN"); return e && std::strcmp(e, "v1") == 0; }();
  const int Np16 = (N + 15) / 16 * 16;
  const size_t smem_p = sizeof(__nv_bfloat16) * ((size_t)TEAMS * 2 * Np16 * WS + (size_t)4 * DM * WS) + (tail ? TAIL_BYTES : 0);
  if ((tail || !force_v1) && (int)smem_p <= c.smem_optin && (N + 15) / 16 <= 2 * TW) {
    const unsigned grid = (unsigned)std::min<int64_t>((B + TEAMS - 1) / TEAMS, c.num_sms);
    static const int var = [] { const char* e = std::getenv("PAUT_ATTN_VARIANT"); return e ? std::atoi(e) : 0; }();
    static const unsigned skew = [] { const char* e = std::getenv("PAUT_ATTN_SKEW_NS"); return e ? (unsigned)std::atoi(e) : 12000u; }();
    auto launch = [&](auto kern) {
      smem_optin(c, kern);
      kern<<<grid, AB_WARPS * 32, smem_p, c.stream>>>(p, (int)B, Np16, skew);
    };
    switch (var) {                                   // timing experiments (tools/r2t.sh); 0 = the product kernel
      case 1: launch(k_msc_attn_block_p<1>); break;
      case 2: launch(k_msc_attn_block_p<2>); break;
      case 3: launch(k_msc_attn_block_p<3>); break;
      case 4: launch(k_msc_attn_block_p<4>); break;
      case 5: launch(k_msc_attn_block_p<5>); break;
      case 6: launch(k_msc_attn_block_p<6>); break;
      default: launch(k_msc_attn_block_p<0>); break;
    }
    c.launched("msc_attn_block");
    return;
  }
  smem_optin(c, k_msc_attn_block);
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

bool ts_heads_supported(int d, int hidden) { return d == HW && hidden == HH; }

void op_ts_heads(Ctx& c, const float* x, const float* ng, const float* nb, const void* const* W0, const float* const* b0,
                 const float* const* lg, const float* const* lb, const float* const* W4, const float* const* b4, const int* act,
                 const float* eps, float* const* out, int64_t M) {
  if (c.dry) return;
  TsHeadsArgs p;
  p.x = x; p.ng = ng; p.nb = nb; p.M = M;
  for (int h = 0; h < 4; ++h) {
    p.W0[h] = static_cast<const __nv_bfloat16*>(W0[h]); p.b0[h] = b0[h]; p.lg[h] = lg[h]; p.lb[h] = lb[h];
    p.W4[h] = W4[h]; p.b4[h] = b4[h]; p.act[h] = act[h]; p.eps[h] = eps[h]; p.out[h] = out[h];
  }
  const size_t smem = (size_t)4 * HH * HS * sizeof(__nv_bfloat16);
  smem_optin(c, k_ts_heads);
  const int64_t tiles = (M + 15) / 16;
  int64_t blocks = (tiles + 7) / 8;
  if (blocks > (int64_t)c.num_sms * 3) blocks = (int64_t)c.num_sms * 3;   // 70 KB of staged weights: three CTAs per SM
  k_ts_heads<<<(unsigned)blocks, 256, smem, c.stream>>>(p);
  c.launched("ts_heads");
}

bool lin_res_ln_supported(int N, int K, int lda) { return N == LRL_N && K % 32 == 0 && K >= 32 && K <= 512 && lda % 2 == 0; }

void op_lin_res_ln(Ctx& c, const float* x, int lda, const void* Wrow, int K, const float* bias, const float* res, const float* g,
                   const float* b, float* out, int64_t M, const float* table, int table_mod) {
  if (c.dry) return;
  PAUT_CHECK(lin_res_ln_supported(LRL_N, K, lda), PAUT_ERR_UNSUPPORTED, "lin_res_ln: unsupported shape");
  LinResLnArgs p;
  p.x = x; p.lda = lda; p.K = K; p.W = static_cast<const __nv_bfloat16*>(Wrow); p.bias = bias; p.res = res; p.g = g; p.b = b;
  p.out = out; p.M = M; p.table = table; p.table_mod = table_mod > 0 ? table_mod : 1;
  const size_t smem = (size_t)LRL_N * (K + 8) * sizeof(__nv_bfloat16);
  const int64_t tiles = (M + 15) / 16;
  if (smem > 100 * 1024) {                                 // one CTA of 16 warps per SM
    smem_optin(c, k_lin_res_ln<512>);
    int64_t blocks = (tiles + 15) / 16;
    if (blocks > (int64_t)c.num_sms) blocks = c.num_sms;
    k_lin_res_ln<512><<<(unsigned)blocks, 512, smem, c.stream>>>(p);
  } else {                                                 // two CTAs of 8 warps per SM
    smem_optin(c, k_lin_res_ln<256>);
    int64_t blocks = (tiles + 7) / 8;
    if (blocks > (int64_t)c.num_sms * 2) blocks = (int64_t)c.num_sms * 2;
    k_lin_res_ln<256><<<(unsigned)blocks, 256, smem, c.stream>>>(p);
  }
  c.launched("lin_res_ln");
}

}  // namespace paut
