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
//     per head h, per block of <= 160 keys (flash-style, at most two blocks):
//                            S = Q_h K_h^T  (128 x 160 x 16, ONE K step; accumulator = 128 lanes x 160 columns)
//                            softmax: thread = row: block maximum, running maximum m (the accumulated O and row sum of
//                            the first block are rescaled by 2^(m_old - m) in tensor memory), P = exp2(S - m) as packed
//                            fp16 pairs written IN PLACE over the S columns already consumed (tcgen05.st), in the
//                            A-operand layout of the next product
//                            O_h (+)= P V_h  (TS form: A = P in TMEM, B = V^T rows 16h..16h+15; a second N = 16 block
//                            against a row of ones yields the row sums, so the denominator costs no thread work)
//                            O_h / sum -> bf16 -> the SMEM chunks Q_h occupied (dead after S)
//                            out = LayerNorm(x + O Wo^T + bo)
// TWO independent chains run per CTA: warp group g (4 warps = the four TMEM lane quarters) owns heads g and g + 2 with
// its own S and O columns; while one group waits for its MMAs the other one computes, and the MMA warp serves
// whichever group has handed over its P.  (The first version ran one chain with two warps per row: every S -> P -> O
// hand-over was an exposed round trip, 96 k cycles per set against 64 k of the mma.sync block.)
#include <cuda_fp16.h>

#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "mma_common.cuh"
#include "tc_common.cuh"

namespace paut {

using namespace tc;

namespace {

constexpr int AT_COMPUTE = 256;                   // 8 compute warps: group g = warp / 4, TMEM lane quarter = warp % 4
constexpr int AT_THREADS = AT_COMPUTE + 32;       // + MMA issuer warp
constexpr int AT_MMA_WARP = 8;
constexpr int AT_DM = 64, AT_NH = 4;
constexpr int AT_NMAX = 320;                      // keys per set (padded to 16)
constexpr int AT_KBLK = 160;                      // keys per softmax block
constexpr int AT_XROWS = 392;                     // 3 tiles of 128 rows + the shifted view's extra row, multiple of 8
constexpr int AT_LBO_X = AT_XROWS * 16;           // chunk strides (bytes) of the K-major operand tiles
constexpr int AT_LBO_WQK = 128 * 16, AT_LBO_W64 = 64 * 16, AT_LBO_Q = 128 * 16, AT_LBO_K = AT_NMAX * 16;
constexpr int AT_VROWS = 80, AT_LBO_V = AT_VROWS * 16;   // V^T rows: 64 d + a row of ones + 15 zero rows
constexpr int AT_XB = 8 * AT_LBO_X, AT_WQK = 8 * AT_LBO_WQK, AT_W64 = 8 * AT_LBO_W64, AT_QB = 8 * AT_LBO_Q,
              AT_KB = 8 * AT_LBO_K, AT_VT = (AT_NMAX / 8) * AT_LBO_V;
constexpr int AT_SMEM = AT_XB + AT_WQK + 2 * AT_W64 + AT_QB + AT_KB + AT_VT;   // 191,488 B
// tensor memory columns.  phase B: group g: S (P in place) at 192 g, O_h (16) + ones block (16) at 192 g + 160;
// Q / Y accumulator at 384.  phase A: V^T accumulator at 0 (<= 320), K accumulators of the three tiles at 320.
constexpr int TC_G = 192, TC_O = 160, TC_QY = 384, TC_VT = 0, TC_K = 320;

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

// Parity wait (a failed probe puts the warp to sleep before the next one), with a dead-lock guard: a protocol error
// traps instead of hanging the GPU.
template <int SLEEP_NS>
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
    if (SLEEP_NS > 0) asm volatile("nanosleep.u32 %0;" ::"r"(SLEEP_NS));
    if (spin > (1u << 22)) __trap();
  }
}
// non-blocking probe
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
      "selp.u32 %0, 1, 0, P1;\n"
      "}\n"
      : "=r"(done)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return done != 0;
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
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
      "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
      "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]),
      "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]),
      "r"(v[31])
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
  // MMA -> threads (count 1, tcgen05.commit):  bar_a, bar_q, s_full[g], bar_y
  // threads -> MMA (one arrival per warp):     bar_x, bar_free, bar_qs, bar_os (8 warps), p_full[g] (4 warps)
  __shared__ __align__(8) uint64_t bar_x, bar_a, bar_free, bar_q, bar_qs, s_full[2], p_full[2], bar_os, bar_y;
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float bias_s[192], bo_s[64], g_s[64], b_s[64];

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
  const int KB0 = Np < AT_KBLK ? Np : AT_KBLK, KB1 = Np - KB0;    // key blocks (KB1 = 0: one block)
  const int nblk = KB1 > 0 ? 2 : 1;
  const int tiles = (N + 127) / 128;

  // ---- one-time setup
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  if (tid == 0) {
    mbar_init(&bar_a, 1); mbar_init(&bar_q, 1); mbar_init(&bar_y, 1);
    mbar_init(&bar_x, AT_COMPUTE / 32); mbar_init(&bar_free, AT_COMPUTE / 32); mbar_init(&bar_qs, AT_COMPUTE / 32);
    mbar_init(&bar_os, AT_COMPUTE / 32);
    for (int g = 0; g < 2; ++g) { mbar_init(&s_full[g], 1); mbar_init(&p_full[g], 4); }
    fence_mbar_init();
  }
  for (int i = tid; i < AT_WQK / 16; i += AT_THREADS) reinterpret_cast<uint4*>(WQK)[i] = __ldg(reinterpret_cast<const uint4*>(p.Wqk) + i);
  for (int i = tid; i < AT_W64 / 16; i += AT_THREADS) {
    reinterpret_cast<uint4*>(WV)[i] = __ldg(reinterpret_cast<const uint4*>(p.Wv) + i);
    reinterpret_cast<uint4*>(WO)[i] = __ldg(reinterpret_cast<const uint4*>(p.Wo) + i);
  }
  for (int i = tid; i < AT_XB / 16; i += AT_THREADS) reinterpret_cast<uint4*>(XB)[i] = make_uint4(0u, 0u, 0u, 0u);
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
    uint32_t n_x = 0, n_free = 0, n_qs = 0, n_os = 0, n_p[2] = {0, 0};   // uses of the barriers this warp waits on
    // S of head h, key block kb, for group g: one K step of 16
    auto issue_s = [&](int g, int h, int kb) {
      const int n = kb ? KB1 : KB0;
      mma_bf16_ss2(tmem + (uint32_t)(TC_G * g), lo(qb_a + 2 * h * AT_LBO_Q, AT_LBO_Q), hi,
                   lo(kb_a + 2 * h * AT_LBO_K + kb * AT_KBLK * 16, AT_LBO_K), hi, make_idesc_bf16(128, n), 0u);
    };
    // O_h (+)= P V_h and the row sums over the keys of block kb (P sits in place of S: chunk c of 32 S columns holds
    // the fp16 pairs of its 32 keys in its first 16 columns)
    auto issue_pv = [&](int g, int h, int kb) {
      const int n = kb ? KB1 : KB0;
      const uint32_t idesc = make_idesc_f16(128, 16);
      for (int ks = 0; ks < n / 16; ++ks) {
        const uint32_t vb = vt_a + (uint32_t)((kb * (AT_KBLK / 8) + 2 * ks) * AT_LBO_V);
        const uint64_t b_h = ((uint64_t)hi << 32) | lo(vb + 16 * h * 16, AT_LBO_V);
        const uint64_t b_1 = ((uint64_t)hi << 32) | lo(vb + 64 * 16, AT_LBO_V);
        const uint32_t a = tmem + (uint32_t)(TC_G * g + 32 * (ks >> 1) + 8 * (ks & 1));
        const uint32_t acc = (kb | ks) ? 1u : 0u;
        mma_f16_ts(tmem + (uint32_t)(TC_G * g + TC_O), a, b_h, idesc, acc);
        mma_f16_ts(tmem + (uint32_t)(TC_G * g + TC_O + 16), a, b_1, idesc, acc);
      }
    };
    for (long long set = blockIdx.x; set < p.B; set += gridDim.x) {
      // ---- phase A: V^T and K projections of the whole set
      if (set != (long long)blockIdx.x) mbar_wait_g<0>(&bar_free, n_free++ & 1);   // the previous set's last Y is drained
      mbar_wait_g<0>(&bar_x, n_x++ & 1);                            // x of this set is in shared memory
      if (leader) {
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint32_t a = lo(wv_a + ks * 2 * AT_LBO_W64, AT_LBO_W64);
          mma_bf16_ss2(tmem + TC_VT, a, hi, lo(xb_a + sh * 16 + ks * 2 * AT_LBO_X, AT_LBO_X), hi, make_idesc_bf16(128, KB0), ks ? 1u : 0u);
          if (KB1 > 0)
            mma_bf16_ss2(tmem + TC_VT + KB0, a, hi, lo(xb_a + (sh + KB0) * 16 + ks * 2 * AT_LBO_X, AT_LBO_X), hi,
                         make_idesc_bf16(128, KB1), ks ? 1u : 0u);
        }
        for (int t = 0; t < tiles; ++t)
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            mma_bf16_ss2(tmem + TC_K + 64 * t, lo(xb_a + (sh + 128 * t) * 16 + ks * 2 * AT_LBO_X, AT_LBO_X), hi,
                         lo(wqk_a + 64 * 16 + ks * 2 * AT_LBO_WQK, AT_LBO_WQK), hi, make_idesc_bf16(128, 64), ks ? 1u : 0u);
        mma_commit(&bar_a);
      }
      __syncwarp();
      for (int t = 0; t < tiles; ++t) {
        // ---- Q projection of the tile (its accumulator is also the previous tile's Y)
        mbar_wait_g<0>(&bar_free, n_free++ & 1);
        if (leader) {
          tc_fence_after();
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            mma_bf16_ss2(tmem + TC_QY, lo(xb_a + 128 * t * 16 + ks * 2 * AT_LBO_X, AT_LBO_X), hi,
                         lo(wqk_a + ks * 2 * AT_LBO_WQK, AT_LBO_WQK), hi, make_idesc_bf16(128, 64), ks ? 1u : 0u);
          mma_commit(&bar_q);
        }
        __syncwarp();
        mbar_wait_g<0>(&bar_qs, n_qs++ & 1);                        // Q (and, for tile 0, K and V^T) are in shared memory
        if (leader) {
          tc_fence_after();
          issue_s(0, 0, 0); mma_commit(&s_full[0]);
          issue_s(1, 1, 0); mma_commit(&s_full[1]);
        }
        __syncwarp();
        // ---- serve the two groups: step k of group g = (head g + 2 (k / nblk), key block k % nblk)
        int step[2] = {0, 0};
        const int nsteps = 2 * nblk;
        uint32_t idle = 0;
        while (step[0] < nsteps || step[1] < nsteps) {
          bool served = false;
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            if (step[g] < nsteps && mbar_test(&p_full[g], n_p[g] & 1)) {
              ++n_p[g];
              const int k = step[g]++;
              const int h = g + 2 * (k / nblk), kb = k % nblk;
              if (leader) {
                tc_fence_after();
                issue_pv(g, h, kb);                                  // reads P, then S may be overwritten:
                if (k + 1 < nsteps) issue_s(g, g + 2 * ((k + 1) / nblk), (k + 1) % nblk);
                mma_commit(&s_full[g]);                              // next S ready AND this block's O accumulated
              }
              __syncwarp();
              served = true;
            }
          }
          if (!served && ++idle > (1u << 26)) __trap();
        }
        // ---- out-projection: A = normalised attention output (bf16, in the Q buffer), B = Wo
        mbar_wait_g<0>(&bar_os, n_os++ & 1);
        if (leader) {
          tc_fence_after();
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            mma_bf16_ss2(tmem + TC_QY, lo(qb_a + ks * 2 * AT_LBO_Q, AT_LBO_Q), hi, lo(wo_a + ks * 2 * AT_LBO_W64, AT_LBO_W64), hi,
                         make_idesc_bf16(128, 64), ks ? 1u : 0u);
          mma_commit(&bar_y);
        }
        __syncwarp();
      }
    }
  } else {
    // ================= compute warps =================
    const int q = warp & 3, grp = warp >> 2;        // TMEM lane quarter, group (chain)
    const int row = q * 32 + lane;                  // row of the tile = TMEM lane
    const uint32_t t_lane = (uint32_t)(q * 32) << 16;
    const uint32_t s_col = tmem + t_lane + (uint32_t)(TC_G * grp);     // this group's S / P columns, this thread's lane
    uint32_t n_a = 0, n_q = 0, n_s = 0, n_y = 0;
    const float qscale = 0.25f * 1.4426950408889634f;               // 1/sqrt(16) * log2(e)
    // phase A column split (keys of the V^T accumulator): group 0 the first block, group 1 the second
    const int col0 = grp == 0 ? 0 : KB0, ncol = grp == 0 ? KB0 : KB1;
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
      mbar_wait_g<40>(&bar_a, n_a++ & 1);
      tc_fence_after();
      if (q < 2) {
        // V^T: lane = d (64 used rows), columns = keys: + bias -> fp16 -> [key chunk][d][16 B]
        const int d = row;
        const float bv = bias_s[128 + d];
        for (int c0 = 0; c0 < ncol; c0 += 32) {
          uint32_t r[32];
          tmem_ld32(tmem + t_lane + TC_VT + col0 + c0, r);
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
        tmem_ld32(tmem + t_lane + TC_K + 64 * t + 32 * grp, r);
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
      warp_arrive(&bar_free, lane);                                  // phase A accumulators drained, K / V^T stored

      for (int t = 0; t < tiles; ++t) {
        const int r0 = 128 * t;
        // ---- Q epilogue: + bias, * log2(e)/4 -> bf16 -> Q operand
        mbar_wait_g<40>(&bar_q, n_q++ & 1);
        tc_fence_after();
        {
          uint32_t r[32];
          tmem_ld32(tmem + t_lane + TC_QY + 32 * grp, r);
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
        const bool active = r0 + 32 * q < N;
        // O_h / row sum -> bf16 -> the Q buffer's chunks of head h (the out-projection's A operand)
        auto o_epilogue = [&](int h) {
          uint32_t o[32];
          tmem_ld32(s_col + TC_O, o);
          const float inv = 1.f / __uint_as_float(o[16]);
          uint32_t w[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) w[j] = mma::pack_bf16(__uint_as_float(o[2 * j]) * inv, __uint_as_float(o[2 * j + 1]) * inv);
          st_shared_v4(qb_a + (uint32_t)((2 * h) * AT_LBO_Q + row * 16), w[0], w[1], w[2], w[3]);
          st_shared_v4(qb_a + (uint32_t)((2 * h + 1) * AT_LBO_Q + row * 16), w[4], w[5], w[6], w[7]);
        };
        for (int hi2 = 0; hi2 < 2; ++hi2) {
          const int h = grp + 2 * hi2;
          float m = -INFINITY;
          for (int kb = 0; kb < nblk; ++kb) {
            const int kn = kb ? KB1 : KB0, key0 = kb * AT_KBLK;
            mbar_wait_g<40>(&s_full[grp], n_s++ & 1);               // S of this block; the previous block's O is accumulated
            tc_fence_after();
            if (kb == 0 && hi2 > 0) o_epilogue(h - 2);               // before this head's first P V overwrites O
            if (active) {
              // ---- pass 1: block maximum
              float bm = -INFINITY;
              for (int c0 = 0; c0 < kn; c0 += 32) {
                uint32_t r[32];
                tmem_ld32(s_col + c0, r);
                if (key0 + c0 + 32 <= N) {
#pragma unroll
                  for (int j = 0; j < 32; ++j) bm = fmaxf(bm, __uint_as_float(r[j]));
                } else {
#pragma unroll
                  for (int j = 0; j < 32; ++j)
                    if (key0 + c0 + j < N) bm = fmaxf(bm, __uint_as_float(r[j]));
                }
              }
              if (kb > 0) {
                // running maximum: rescale what the first block accumulated (O_h and the row sum) in tensor memory
                const float m1 = fmaxf(m, bm);
                const float c = mma::fast_exp2(m - m1);
                m = m1;
                uint32_t o[32];
                tmem_ld32(s_col + TC_O, o);
#pragma unroll
                for (int j = 0; j < 17; ++j) o[j] = __float_as_uint(__uint_as_float(o[j]) * c);
                tmem_st32(s_col + TC_O, o);
              } else {
                m = bm;
              }
              // ---- pass 2: P = exp2(S - m) as packed fp16 pairs, in place (first 16 columns of every 32-column chunk)
              for (int c0 = 0; c0 < kn; c0 += 32) {
                uint32_t r[32];
                tmem_ld32(s_col + c0, r);
                uint32_t pk[16];
                if (key0 + c0 + 32 <= N) {
#pragma unroll
                  for (int j = 0; j < 16; ++j)
                    pk[j] = pack_f16(mma::fast_exp2(__uint_as_float(r[2 * j]) - m), mma::fast_exp2(__uint_as_float(r[2 * j + 1]) - m));
                } else {
#pragma unroll
                  for (int j = 0; j < 16; ++j) {
                    const float a = key0 + c0 + 2 * j < N ? mma::fast_exp2(__uint_as_float(r[2 * j]) - m) : 0.f;
                    const float b = key0 + c0 + 2 * j + 1 < N ? mma::fast_exp2(__uint_as_float(r[2 * j + 1]) - m) : 0.f;
                    pk[j] = pack_f16(a, b);
                  }
                }
                tmem_st16(s_col + c0, pk);
              }
              asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            }
            tc_fence_before();
            warp_arrive(&p_full[grp], lane);
          }
        }
        // ---- O of the group's last head, then the out-projection's operand is complete
        mbar_wait_g<40>(&s_full[grp], n_s++ & 1);
        tc_fence_after();
        o_epilogue(grp + 2);
        fence_async_smem();
        tc_fence_before();
        warp_arrive(&bar_os, lane);
        // ---- final epilogue: y = O Wo^T + bo + x -> LayerNorm -> out   (warp group 0: one thread per row)
        mbar_wait_g<40>(&bar_y, n_y++ & 1);
        tc_fence_after();
        if (grp == 0) {
          float y[64];
          {
            uint32_t r[32];
            tmem_ld32(tmem + t_lane + TC_QY, r);
#pragma unroll
            for (int j = 0; j < 32; ++j) y[j] = __uint_as_float(r[j]);
            tmem_ld32(tmem + t_lane + TC_QY + 32, r);
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
        warp_arrive(&bar_free, lane);                                // Y drained: the next Q projection / set may start
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
