// Attention block of MultiSignalClassifier / _N on tcgen05 (bf16 mode, d_model 64, 4 heads of 16):
//   x[N,64] -> QKV projection -> softmax(Q K^T / 4) V per head -> out-projection -> + x -> LayerNorm
//   (TransformerEncoder.forward, NN_models.py:31-37; kv_shift: keys / values from the sequence shifted left by one
//   with the last row repeated, :35).
// One persistent CTA per SM walks the sets; every product is a tcgen05.mma with the accumulator in tensor memory:
//
//   phase A (once per set)   V^T[d, key] = Wv x^T   (M = d rows (64 used), N = keys, K = 64)   -> fp16 V^T operand in SMEM
//                            K[key, d]   = x Wk^T   per 128-row tile                            -> bf16 K operand in SMEM
//   phase B (per 128 query rows)
//                            Q = x Wq^T, scaled by log2(e)/4                                    -> bf16 Q operand in SMEM
//     per head h:            S = Q_h K_h^T  (128 x Np x 16, ONE K step, accumulator = 128 lanes x Np columns)
//                            softmax: two warps per lane quarter split the key columns: row max (pass 1, exchanged
//                            through SMEM), P = exp2(S - max) as packed fp16 pairs written straight back into TENSOR
//                            MEMORY (tcgen05.st) in the A-operand layout of the next product
//                            O_h = P V_h  (TS form: A = P in TMEM, B = V^T rows 16h..16h+15; a second N = 16 block
//                            against a row of ones yields the row sums, so the denominator costs no thread work)
//                            O_h / sum -> bf16 -> the SMEM chunks Q_h occupied (dead after S)
//                            out = LayerNorm(x + O Wo^T + bo)
// The S accumulator of head h+1 is issued as soon as the P of head h is complete (P has its own TMEM columns), so the
// tensor pipe works under the threads' softmax; the exp (MUFU, one ex2.f16x2 per pair) is the binding unit.
#include <cuda_fp16.h>

#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "mma_common.cuh"
#include "tc_common.cuh"

namespace paut {

using namespace tc;

namespace {

constexpr int AT_COMPUTE = 256;                   // 8 compute warps: warps w and w + 4 share TMEM lane quarter w
constexpr int AT_THREADS = AT_COMPUTE + 32;       // + MMA issuer warp
constexpr int AT_MMA_WARP = 8;
constexpr int AT_DM = 64, AT_HD = 16, AT_NH = 4;
constexpr int AT_NMAX = 320;                      // keys per set (padded to 16)
constexpr int AT_XROWS = 392;                     // 3 tiles of 128 rows + the shifted view's extra row, multiple of 8
constexpr int AT_LBO_X = AT_XROWS * 16;           // chunk strides (bytes) of the K-major operand tiles
constexpr int AT_LBO_WQK = 128 * 16, AT_LBO_W64 = 64 * 16, AT_LBO_Q = 128 * 16, AT_LBO_K = AT_NMAX * 16;
constexpr int AT_VROWS = 80, AT_LBO_V = AT_VROWS * 16;   // V^T rows: 64 d + a row of ones + 15 zero rows
constexpr int AT_XB = 8 * AT_LBO_X, AT_WQK = 8 * AT_LBO_WQK, AT_W64 = 8 * AT_LBO_W64, AT_QB = 8 * AT_LBO_Q,
              AT_KB = 8 * AT_LBO_K, AT_VT = (AT_NMAX / 8) * AT_LBO_V;
constexpr int AT_SMEM = AT_XB + AT_WQK + 2 * AT_W64 + AT_QB + AT_KB + AT_VT;   // 191,488 B
// tensor memory columns
constexpr int TC_S = 0;                           // S accumulator (phase A: V^T accumulator)
constexpr int TC_P = 320;                         // P operand, 160 columns (phase A: K accumulators; also Q and Y accumulators)
constexpr int TC_O = 480;                         // O_h (16) + ones block (16)

struct AttnTcArgs {
  const float* x;                     // [B, N, 64]
  float* out;                         // [B, N, 64]
  const __nv_bfloat16* Wqk;           // in_proj rows 0..127 (q | k), K-major chunks [8][128][8]
  const __nv_bfloat16* Wv;            // in_proj rows 128..191, [8][64][8]
  const __nv_bfloat16* Wo;            // out_proj, [8][64][8]
  const float* bqkv;                  // [192]
  const float* bo;
  const float* ln_g;
  const float* ln_b;
  long long B;
  int N, kv_shift;
};

__device__ __forceinline__ void named_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Parity wait that does not burn issue slots (a failed probe puts the warp to sleep before the next one: a spinning
// warp takes the issue slots of the working warps on its scheduler), with a dead-lock guard: a protocol error traps
// instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait_g(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  for (uint32_t spin = 0;; ++spin) {
    uint32_t done;
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
        "selp.u32 %0, 1, 0, P1;\n"
        "}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) break;
    asm volatile("nanosleep.u32 %0;" ::"r"(40));
    if (spin > (1u << 22)) __trap();
  }
}
// one arrival per warp: every lane has executed its fences, lane 0 arrives for the warp
__device__ __forceinline__ void warp_arrive(uint64_t* bar, int lane) {
  __syncwarp();
  if (lane == 0) mbar_arrive(bar);
}

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
      "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

__device__ __forceinline__ uint32_t pack_f16(float lo, float hi) {
  const __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&h);
}

__global__ void __launch_bounds__(AT_THREADS, 1) k_msc_attn_tc(const AttnTcArgs p) {
  extern __shared__ __align__(128) unsigned char smem[];
  // MMA -> threads (count 1, tcgen05.commit):  bar_a, bar_q, s_full, pv_done, bar_y
  // threads -> MMA (one arrival per warp):     bar_x, bar_free, bar_qs, p_full, bar_os
  __shared__ __align__(8) uint64_t bar_x, bar_a, bar_free, bar_q, bar_qs, s_full, p_full, pv_done, bar_os, bar_y;
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float bias_s[192], bo_s[64], g_s[64], b_s[64];
  __shared__ float smax[2][128];

  unsigned char* XB = smem;                        // x of the set, bf16 [8 chunks][392 rows][16 B] (rows >= N zero)
  unsigned char* WQK = XB + AT_XB;                 // [8][128 rows: q | k][16 B]
  unsigned char* WV = WQK + AT_WQK;                // [8][64][16 B]  (A operand: rows 64..127 of the M tile are don't-care)
  unsigned char* WO = WV + AT_W64;                 // [8][64][16 B]
  unsigned char* QB = WO + AT_W64;                 // [8][128][16 B]: Q of the tile; chunks 2h, 2h+1 later hold O_h
  unsigned char* KB = QB + AT_QB;                  // [8][320 key rows][16 B]
  unsigned char* VT = KB + AT_KB;                  // [40 key chunks][80 rows][16 B], fp16

  const int tid = threadIdx.x, lane = tid & 31, warp = uniform_warp_id();
  const int N = p.N;
  const int Np = (N + 15) & ~15;                   // keys padded to the MMA's N granularity
  const int NA = Np < 160 ? Np : 160, NB = Np - NA;   // key columns of warp group 0 / 1 (= the two S MMAs)
  const int tiles = (N + 127) / 128;

  // ---- one-time setup
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  if (tid == 0) {
    mbar_init(&bar_a, 1); mbar_init(&bar_q, 1); mbar_init(&s_full, 1); mbar_init(&pv_done, 1); mbar_init(&bar_y, 1);
    mbar_init(&bar_x, AT_COMPUTE / 32); mbar_init(&bar_free, AT_COMPUTE / 32); mbar_init(&bar_qs, AT_COMPUTE / 32);
    mbar_init(&p_full, AT_COMPUTE / 32); mbar_init(&bar_os, AT_COMPUTE / 32);
    fence_mbar_init();
  }
  for (int i = tid; i < AT_WQK / 16; i += AT_THREADS) reinterpret_cast<uint4*>(WQK)[i] = __ldg(reinterpret_cast<const uint4*>(p.Wqk) + i);
  for (int i = tid; i < AT_W64 / 16; i += AT_THREADS) {
    reinterpret_cast<uint4*>(WV)[i] = __ldg(reinterpret_cast<const uint4*>(p.Wv) + i);
    reinterpret_cast<uint4*>(WO)[i] = __ldg(reinterpret_cast<const uint4*>(p.Wo) + i);
  }
  for (int i = tid; i < (AT_XB + 0) / 16; i += AT_THREADS) reinterpret_cast<uint4*>(XB)[i] = make_uint4(0u, 0u, 0u, 0u);
  for (int i = tid; i < (AT_QB + AT_KB + AT_VT) / 16; i += AT_THREADS) reinterpret_cast<uint4*>(QB)[i] = make_uint4(0u, 0u, 0u, 0u);
  for (int i = tid; i < 192; i += AT_THREADS) bias_s[i] = __ldg(p.bqkv + i);
  for (int i = tid; i < 64; i += AT_THREADS) { bo_s[i] = __ldg(p.bo + i); g_s[i] = __ldg(p.ln_g + i); b_s[i] = __ldg(p.ln_b + i); }
  __syncthreads();
  // the row of ones (V^T row 64): 1.0 for real keys, 0 for the padding -> the extra MMA block yields the row sums of P
  for (int j = tid; j < N; j += AT_THREADS)
    reinterpret_cast<__half*>(VT + (size_t)(j >> 3) * AT_LBO_V + 64 * 16)[j & 7] = __float2half_rn(1.f);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t xb_a = smem_u32(XB), qb_a = smem_u32(QB), kb_a = smem_u32(KB), vt_a = smem_u32(VT);
  const uint32_t sh = p.kv_shift ? 1u : 0u;        // keys / values read x one row further down

  if (warp == AT_MMA_WARP) {
    // ================= MMA issuer =================
    const bool leader = elect_one();
    const uint32_t hi = (uint32_t)(128 >> 4) | (1u << 14);          // descriptor high word: SBO = 128 B, version 1
    auto lo = [](uint32_t addr, uint32_t lbo) { return ((addr >> 4) & 0x3FFFu) | ((lbo >> 4) << 16); };
    const uint32_t wqk_a = smem_u32(WQK), wv_a = smem_u32(WV), wo_a = smem_u32(WO);
    uint32_t n_x = 0, n_free = 0, n_qs = 0, n_p = 0, n_os = 0;      // uses of the barriers this warp waits on
    for (long long set = blockIdx.x; set < p.B; set += gridDim.x) {
      // ---- phase A: V^T and K projections of the whole set
      if (set != (long long)blockIdx.x) mbar_wait_g(&bar_free, n_free++ & 1);   // the previous set's last Y is drained
      mbar_wait_g(&bar_x, n_x++ & 1);                               // x of this set is in shared memory
      if (leader) {
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint32_t a = lo(wv_a + ks * 2 * AT_LBO_W64, AT_LBO_W64);
          mma_bf16_ss2(tmem + TC_S, a, hi, lo(xb_a + sh * 16 + ks * 2 * AT_LBO_X, AT_LBO_X), hi, make_idesc_bf16(128, NA), ks ? 1u : 0u);
          if (NB > 0)
            mma_bf16_ss2(tmem + TC_S + NA, a, hi, lo(xb_a + (sh + NA) * 16 + ks * 2 * AT_LBO_X, AT_LBO_X), hi,
                         make_idesc_bf16(128, NB), ks ? 1u : 0u);
        }
        for (int t = 0; t < tiles; ++t)
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            mma_bf16_ss2(tmem + TC_P + 64 * t, lo(xb_a + (sh + 128 * t) * 16 + ks * 2 * AT_LBO_X, AT_LBO_X), hi,
                         lo(wqk_a + 64 * 16 + ks * 2 * AT_LBO_WQK, AT_LBO_WQK), hi, make_idesc_bf16(128, 64), ks ? 1u : 0u);
        mma_commit(&bar_a);
      }
      __syncwarp();
      for (int t = 0; t < tiles; ++t) {
        // ---- Q projection of the tile (its accumulator aliases K accumulator 0 / the previous tile's Y)
        mbar_wait_g(&bar_free, n_free++ & 1);
        if (leader) {
          tc_fence_after();
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            mma_bf16_ss2(tmem + TC_P, lo(xb_a + 128 * t * 16 + ks * 2 * AT_LBO_X, AT_LBO_X), hi,
                         lo(wqk_a + ks * 2 * AT_LBO_WQK, AT_LBO_WQK), hi, make_idesc_bf16(128, 64), ks ? 1u : 0u);
          mma_commit(&bar_q);
        }
        __syncwarp();
        mbar_wait_g(&bar_qs, n_qs++ & 1);                           // Q (and, for tile 0, K and V^T) are in shared memory
        auto issue_s = [&](int h) {                                 // S = Q_h K_h^T: one K step of 16
          const uint32_t a = lo(qb_a + 2 * h * AT_LBO_Q, AT_LBO_Q);
          mma_bf16_ss2(tmem + TC_S, a, hi, lo(kb_a + 2 * h * AT_LBO_K, AT_LBO_K), hi, make_idesc_bf16(128, NA), 0u);
          if (NB > 0)
            mma_bf16_ss2(tmem + TC_S + NA, a, hi, lo(kb_a + 2 * h * AT_LBO_K + NA * 16, AT_LBO_K), hi, make_idesc_bf16(128, NB), 0u);
          mma_commit(&s_full);
        };
        if (leader) {
          tc_fence_after();
          issue_s(0);
        }
        __syncwarp();
        for (int h = 0; h < AT_NH; ++h) {
          mbar_wait_g(&p_full, n_p++ & 1);                          // P of head h is in tensor memory, S is free
          if (leader) {
            tc_fence_after();
            if (h + 1 < AT_NH) issue_s(h + 1);
            const uint32_t idesc = make_idesc_f16(128, 16);
            for (int ks = 0; ks < Np / 16; ++ks) {                  // O_h = P V_h and the row sums, K = keys
              const uint32_t vb = vt_a + ks * 2 * AT_LBO_V;
              const uint64_t b_h = ((uint64_t)hi << 32) | lo(vb + 16 * h * 16, AT_LBO_V);
              const uint64_t b_1 = ((uint64_t)hi << 32) | lo(vb + 64 * 16, AT_LBO_V);
              mma_f16_ts(tmem + TC_O, tmem + TC_P + 8 * ks, b_h, idesc, ks ? 1u : 0u);
              mma_f16_ts(tmem + TC_O + 16, tmem + TC_P + 8 * ks, b_1, idesc, ks ? 1u : 0u);
            }
            mma_commit(&pv_done);
          }
          __syncwarp();
        }
        // ---- out-projection: A = normalised attention output (bf16, in the Q buffer), B = Wo
        mbar_wait_g(&bar_os, n_os++ & 1);
        if (leader) {
          tc_fence_after();
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            mma_bf16_ss2(tmem + TC_P, lo(qb_a + ks * 2 * AT_LBO_Q, AT_LBO_Q), hi, lo(wo_a + ks * 2 * AT_LBO_W64, AT_LBO_W64), hi,
                         make_idesc_bf16(128, 64), ks ? 1u : 0u);
          mma_commit(&bar_y);
        }
        __syncwarp();
      }
    }
  } else {
    // ================= compute warps =================
    const int q = warp & 3, grp = warp >> 2;        // TMEM lane quarter, key-column half
    const int row = q * 32 + lane;                  // row of the tile = TMEM lane
    const uint32_t t_lane = (uint32_t)(q * 32) << 16;
    const int col0 = grp == 0 ? 0 : NA;             // this warp's key columns [col0, col0 + ncol)
    const int ncol = grp == 0 ? NA : NB;
    uint32_t n_a = 0, n_q = 0, n_s = 0, n_pv = 0, n_y = 0;
    const float qscale = 0.25f * 1.4426950408889634f;               // 1/sqrt(16) * log2(e)
    for (long long set = blockIdx.x; set < p.B; set += gridDim.x) {
      const float* xs = p.x + (size_t)set * N * AT_DM;
      float* outs = p.out + (size_t)set * N * AT_DM;
      // ---- x -> bf16 K-major operand.  A warp iteration = 8 rows x 4 chunks: lane = (chunk, row) so that the eight
      // lanes of a store phase write one chunk of 8 consecutive rows (128 contiguous bytes, no bank conflict)
      {
        const int r_in = lane & 7, c_in = lane >> 3;
        const int items = ((N + 7) / 8) * 2;
        for (int it = warp; it < items; it += AT_COMPUTE / 32) {
          const int r = (it >> 1) * 8 + r_in, c = (it & 1) * 4 + c_in;
          if (r < N) {
            const float4 v0 = __ldg(reinterpret_cast<const float4*>(xs + (size_t)r * AT_DM + c * 8));
            const float4 v1 = __ldg(reinterpret_cast<const float4*>(xs + (size_t)r * AT_DM + c * 8 + 4));
            const uint32_t w0 = mma::pack_bf16(v0.x, v0.y), w1 = mma::pack_bf16(v0.z, v0.w), w2 = mma::pack_bf16(v1.x, v1.y),
                           w3 = mma::pack_bf16(v1.z, v1.w);
            st_shared_v4(xb_a + (uint32_t)(c * AT_LBO_X + r * 16), w0, w1, w2, w3);
            if (p.kv_shift && r == N - 1) st_shared_v4(xb_a + (uint32_t)(c * AT_LBO_X + N * 16), w0, w1, w2, w3);   // repeated last row
          }
        }
      }
      fence_async_smem();
      warp_arrive(&bar_x, lane);
      {  // next set's x -> L2 while this set is being processed
        const long long nxt = set + gridDim.x;
        if (nxt < p.B) {
          const char* nx = reinterpret_cast<const char*>(p.x + (size_t)nxt * N * AT_DM);
          for (int i = tid * 128; i < N * AT_DM * 4; i += AT_COMPUTE * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + i));
        }
      }
      // ---- phase A epilogues
      mbar_wait_g(&bar_a, n_a++ & 1);
      tc_fence_after();
      if (q < 2) {
        // V^T: lane = d (64 used rows), columns = keys: + bias -> fp16 -> [key chunk][d][16 B]
        const int d = row;
        const float bv = bias_s[128 + d];
        for (int c0 = 0; c0 < ncol; c0 += 32) {
          uint32_t r[32];
          tmem_ld32(tmem + t_lane + TC_S + col0 + c0, r);
#pragma unroll
          for (int k8 = 0; k8 < 4; ++k8) {
            const int key = col0 + c0 + 8 * k8;
            if (key < Np) {
              uint32_t w[4];
#pragma unroll
              for (int j = 0; j < 4; ++j)
                w[j] = pack_f16(__uint_as_float(r[8 * k8 + 2 * j]) + bv, __uint_as_float(r[8 * k8 + 2 * j + 1]) + bv);
              st_shared_v4(vt_a + (uint32_t)((key >> 3) * AT_LBO_V + d * 16), w[0], w[1], w[2], w[3]);
            }
          }
        }
      }
      for (int t = 0; t < tiles; ++t) {
        // K of tile t: lane = key row, this warp's 32 of the 64 columns: + bias -> bf16 -> [chunk][key][16 B]
        const int key = 128 * t + row;
        uint32_t r[32];
        tmem_ld32(tmem + t_lane + TC_P + 64 * t + 32 * grp, r);
        if (key < AT_NMAX) {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint32_t w[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int col = 32 * grp + 8 * c + 2 * j;
              w[j] = mma::pack_bf16(__uint_as_float(r[8 * c + 2 * j]) + bias_s[64 + col], __uint_as_float(r[8 * c + 2 * j + 1]) + bias_s[64 + col + 1]);
            }
            st_shared_v4(kb_a + (uint32_t)((4 * grp + c) * AT_LBO_K + key * 16), w[0], w[1], w[2], w[3]);
          }
        }
      }
      fence_async_smem();
      tc_fence_before();
      warp_arrive(&bar_free, lane);                                        // phase A accumulators drained, K / V^T stored

      for (int t = 0; t < tiles; ++t) {
        const int r0 = 128 * t;
        // ---- Q epilogue: + bias, * log2(e)/4 -> bf16 -> Q operand
        mbar_wait_g(&bar_q, n_q++ & 1);
        tc_fence_after();
        {
          uint32_t r[32];
          tmem_ld32(tmem + t_lane + TC_P + 32 * grp, r);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint32_t w[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int col = 32 * grp + 8 * c + 2 * j;
              w[j] = mma::pack_bf16((__uint_as_float(r[8 * c + 2 * j]) + bias_s[col]) * qscale,
                                    (__uint_as_float(r[8 * c + 2 * j + 1]) + bias_s[col + 1]) * qscale);
            }
            st_shared_v4(qb_a + (uint32_t)((4 * grp + c) * AT_LBO_Q + row * 16), w[0], w[1], w[2], w[3]);
          }
        }
        fence_async_smem();
        tc_fence_before();
        warp_arrive(&bar_qs, lane);

        // a warp whose 32 rows are all beyond the set (last tile) skips the softmax work: its rows of P, O and Y are
        // garbage that is never stored (an MMA output row depends on its own A row only)
        const int ncol_t = r0 + 32 * q < N ? ncol : 0;
        for (int h = 0; h < AT_NH; ++h) {
          // ---- pass 1: row maximum over this warp's key columns
          mbar_wait_g(&s_full, n_s++ & 1);
          tc_fence_after();
          float mx = -INFINITY;
          for (int c0 = 0; c0 < ncol_t; c0 += 32) {
            uint32_t r[32];
            tmem_ld32(tmem + t_lane + TC_S + col0 + c0, r);
            const int kbase = col0 + c0;
            if (kbase + 32 <= N) {
#pragma unroll
              for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(r[j]));
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (kbase + j < N) mx = fmaxf(mx, __uint_as_float(r[j]));
            }
          }
          smax[grp][row] = mx;
          named_sync(1, AT_COMPUTE);
          mx = fmaxf(mx, smax[grp ^ 1][row]);
          // ---- the previous head's O: P and O columns are free again once its MMAs have completed
          if (h > 0) {
            mbar_wait_g(&pv_done, n_pv++ & 1);
            tc_fence_after();
            if (grp == ((h - 1) & 1)) {
              uint32_t o[32];
              tmem_ld32(tmem + t_lane + TC_O, o);
              const float inv = 1.f / __uint_as_float(o[16]);
              uint32_t w[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) w[j] = mma::pack_bf16(__uint_as_float(o[2 * j]) * inv, __uint_as_float(o[2 * j + 1]) * inv);
              st_shared_v4(qb_a + (uint32_t)((2 * (h - 1)) * AT_LBO_Q + row * 16), w[0], w[1], w[2], w[3]);
              st_shared_v4(qb_a + (uint32_t)((2 * (h - 1) + 1) * AT_LBO_Q + row * 16), w[4], w[5], w[6], w[7]);
            }
          }
          // ---- pass 2: P = exp2(S - max) as packed fp16 pairs -> tensor memory (A operand of the P V product)
          for (int c0 = 0; c0 < ncol_t; c0 += 32) {
            uint32_t r[32];
            tmem_ld32(tmem + t_lane + TC_S + col0 + c0, r);
            const int kbase = col0 + c0;
            uint32_t pk[16];
            if (kbase + 32 <= N) {
#pragma unroll
              for (int j = 0; j < 16; ++j) pk[j] = mma::exp2_pair_f16(__uint_as_float(r[2 * j]) - mx, __uint_as_float(r[2 * j + 1]) - mx);
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const float a = kbase + 2 * j < N ? __uint_as_float(r[2 * j]) - mx : -INFINITY;
                const float b = kbase + 2 * j + 1 < N ? __uint_as_float(r[2 * j + 1]) - mx : -INFINITY;
                pk[j] = mma::exp2_pair_f16(a, b);
              }
            }
            tmem_st16(tmem + t_lane + TC_P + (uint32_t)(kbase >> 1), pk);
          }
          asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
          tc_fence_before();
          warp_arrive(&p_full, lane);
        }
        // ---- O of the last head, then the out-projection's operand is complete
        mbar_wait_g(&pv_done, n_pv++ & 1);
        tc_fence_after();
        if (grp == ((AT_NH - 1) & 1)) {
          uint32_t o[32];
          tmem_ld32(tmem + t_lane + TC_O, o);
          const float inv = 1.f / __uint_as_float(o[16]);
          uint32_t w[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) w[j] = mma::pack_bf16(__uint_as_float(o[2 * j]) * inv, __uint_as_float(o[2 * j + 1]) * inv);
          st_shared_v4(qb_a + (uint32_t)((2 * (AT_NH - 1)) * AT_LBO_Q + row * 16), w[0], w[1], w[2], w[3]);
          st_shared_v4(qb_a + (uint32_t)((2 * (AT_NH - 1) + 1) * AT_LBO_Q + row * 16), w[4], w[5], w[6], w[7]);
        }
        fence_async_smem();
        tc_fence_before();
        warp_arrive(&bar_os, lane);
        // ---- final epilogue: y = O Wo^T + bo + x -> LayerNorm -> out   (warp group 0: one thread per row)
        mbar_wait_g(&bar_y, n_y++ & 1);
        tc_fence_after();
        if (grp == 0) {
          float y[64];
          {
            uint32_t r[32];
            tmem_ld32(tmem + t_lane + TC_P, r);
#pragma unroll
            for (int j = 0; j < 32; ++j) y[j] = __uint_as_float(r[j]);
            tmem_ld32(tmem + t_lane + TC_P + 32, r);
#pragma unroll
            for (int j = 0; j < 32; ++j) y[32 + j] = __uint_as_float(r[j]);
          }
          const int gr = r0 + row;
          if (gr < N) {
            const float4* xr = reinterpret_cast<const float4*>(xs + (size_t)gr * AT_DM);
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float4 xv = __ldg(xr + j);
              y[4 * j] += bo_s[4 * j] + xv.x; y[4 * j + 1] += bo_s[4 * j + 1] + xv.y;
              y[4 * j + 2] += bo_s[4 * j + 2] + xv.z; y[4 * j + 3] += bo_s[4 * j + 3] + xv.w;
              s += (y[4 * j] + y[4 * j + 1]) + (y[4 * j + 2] + y[4 * j + 3]);
            }
            const float mean = s * (1.f / AT_DM);
            float v = 0.f;
#pragma unroll
            for (int j = 0; j < 64; ++j) { const float dlt = y[j] - mean; v += dlt * dlt; }
            const float rstd = rsqrtf(v * (1.f / AT_DM) + 1e-5f);
            float4* orow = reinterpret_cast<float4*>(outs + (size_t)gr * AT_DM);
#pragma unroll
            for (int j = 0; j < 16; ++j)
              orow[j] = make_float4((y[4 * j] - mean) * rstd * g_s[4 * j] + b_s[4 * j], (y[4 * j + 1] - mean) * rstd * g_s[4 * j + 1] + b_s[4 * j + 1],
                                    (y[4 * j + 2] - mean) * rstd * g_s[4 * j + 2] + b_s[4 * j + 2], (y[4 * j + 3] - mean) * rstd * g_s[4 * j + 3] + b_s[4 * j + 3]);
          }
        }
        tc_fence_before();
        warp_arrive(&bar_free, lane);                                      // Y drained: the next Q projection / set may start
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

uint16_t f2bf_a(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  return (uint16_t)((u + 0x7fffu + ((u >> 16) & 1u)) >> 16);
}

}  // namespace

bool msc_attn_tc_supported(int N, int d, int heads) { return d == AT_DM && heads == AT_NH && N >= 1 && N <= AT_NMAX; }

// W [rows][64] fp32 (rows row0 .. row0 + rows) -> bf16 K-major chunks [8][rows][8]
void msc_attn_tc_pack(const float* W, int row0, int rows, std::vector<uint16_t>& out) {
  out.assign((size_t)8 * rows * 8, 0);
  for (int n = 0; n < rows; ++n)
    for (int k = 0; k < 64; ++k) out[((size_t)(k >> 3) * rows + n) * 8 + (k & 7)] = f2bf_a(W[(size_t)(row0 + n) * 64 + k]);
}

void op_msc_attn_tc(Ctx& c, const float* x, const void* Wqk, const void* Wv, const void* Wo, const float* bqkv, const float* bo,
                    const float* ln_g, const float* ln_b, float* out, int64_t B, int N, bool kv_shift) {
  if (c.dry) return;
  PAUT_CHECK(msc_attn_tc_supported(N, AT_DM, AT_NH), PAUT_ERR_UNSUPPORTED, "attn block (tcgen05): set too long");
  AttnTcArgs p;
  p.x = x; p.out = out; p.Wqk = static_cast<const __nv_bfloat16*>(Wqk); p.Wv = static_cast<const __nv_bfloat16*>(Wv);
  p.Wo = static_cast<const __nv_bfloat16*>(Wo); p.bqkv = bqkv; p.bo = bo; p.ln_g = ln_g; p.ln_b = ln_b;
  p.B = B; p.N = N; p.kv_shift = kv_shift ? 1 : 0;
  smem_optin(c, k_msc_attn_tc);
  const long long grid = B < c.num_sms ? B : c.num_sms;
  k_msc_attn_tc<<<(unsigned)grid, AT_THREADS, AT_SMEM, c.stream>>>(p);
  c.launched("msc_attn_tc");
}

}  // namespace paut
