// Fused per-A-scan encoder of MultiSignalClassifier on tcgen05 (bf16 mode):
//   x[S] -> Conv1d 1->8 k3 + ReLU -> Conv1d 8->16 k3 + ReLU -> mean over channels -> Linear S->128 + ReLU
//        -> Linear 128->64 + ReLU -> + position table            (NN_models.py:111-121, :11-14)
// One CTA = 128 A-scans; nothing but x and the 64-wide result touches HBM.
//
//  conv1 (C_in = 1, 24 MACs/position) runs on the CUDA cores in packed fp16 (HFMA2), one thread per position;
//  a thread gets the 8-channel vectors of its two neighbours by warp shuffle and writes its im2col row
//  [tap0 | tap1 | tap2 | 1,1,0..] (32 fp16 = 16 words; the constant chunk carries the bias as an fp16 hi+lo
//  pair) straight into TENSOR MEMORY with tcgen05.st.  conv2 is then two TS-form tcgen05.mma per 128
//  positions (A operand in TMEM, B = packed weights in shared memory): 16 cycles each instead of the ~60 an
//  SS-form MMA spends fetching a 4 KB A tile from shared memory, and no im2col bytes ever touch shared memory.
//  Its 32 output columns are the 16 channels (bias included) plus hi/lo halves of their sum, which turns
//  ReLU + channel-mean into   sum_c relu(y_c) = (sum_c y_c + sum_c |y_c|) / 2   -- 18 FADDs per position in
//  the epilogue instead of bias + max + add per channel.  The factor 1/32 is folded into the S->128 weights.
//  The epilogue writes f (bf16) directly in the canonical K-major operand layout of the next GEMM, so the
//  two Linear layers are tcgen05.mma on operands that never left shared memory; their accumulators live
//  in TMEM, reusing the columns of the double-buffered conv stage once it is over.
//
//  Pipeline per group of 2 A-scans (= S/64 M tiles), per TEAM of four compute warps (= one M tile): conv1(g) ->
//  the issuer warp issues the two MMAs of that tile and commits them to the tile's mbarrier -> the team runs the
//  epilogue of its tile of group g-1 while those MMAs execute (TMEM operand rows and accumulators are
//  double-buffered).  Teams only meet at the block-level Linear stages, so their latency chains interleave.  The CTA is persistent over blocks of 128
//  A-scans; the weights of both Linear layers stay resident in shared memory.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include <cuda_fp16.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace paut {

using namespace tc;

namespace {

constexpr int H0 = 128, H1 = 64;
constexpr int ENC_COMPUTE = 640;        // 20 compute warps: 2 conv1 items per thread, 1 epilogue unit (tile, quarter) per warp
constexpr int ENC_THREADS = ENC_COMPUTE + 128;  // + MMA issuer warp (20) + x loader warp (21) + idle warp (22) + edge warp (23):
                                                // warp w issues on scheduler w % 4, and the edge warp's ~60 instructions per
                                                // group are better placed on the scheduler that has no other helper
constexpr int EDGE_WARP = ENC_COMPUTE / 32 + 3;
constexpr int A2_LBO = 2048 + 16;       // chunk stride of the f operand: +16 B skews the chunks across banks
constexpr int D1_COL = 0, D2_COL = 128;     // alias the conv accumulators (dead by then)
constexpr int XS_PAD = 16;
constexpr int MAX_TILES = 5;             // M tiles (teams of 4 compute warps) per group: S / 64 <= 5
constexpr int BAR_BLOCK = 6;             // named barrier of all compute warps (ids 1..5 belong to the teams)
constexpr int XS_SLOTS = 4;              // x staging ring: cp.async prefetch runs ~3 groups ahead of conv1

struct MscEncArgs {
  const void* x;
  int x_dtype;
  int64_t A;
  int S, Nset;
  uint32_t cw[12];                 // conv1d.0.weight as fp16 pairs: [j][t] = channels (2j, 2j+1), tap t -- passed BY VALUE so
  uint32_t cb[4];                  // that the HFMA2s read them from the constant bank instead of 16 registers
  const __nv_bfloat16* Bc;         // conv2 operand [4 chunks][32 rows][8], fp16 bit patterns
  const __nv_bfloat16* W1p;        // shared_layer.0 / 32, packed [S/8][128][8]
  const float* bl1;
  const __nv_bfloat16* W2p;        // shared_layer.2 packed [16][64][8]
  const float* bl2;
  const float* pos;                // [300][64]
  float* h;                        // [A][64]
  unsigned long long* dbg;         // optional cycle probe (PAUT_ENC_DEBUG=1): [warps][8] sums for CTA 0
};

// 18 accumulator columns of one row (16 channels + hi/lo of their sum): the two loads are issued first and
// waited for later (tmem_wait18 names the destination registers so that no use can be scheduled ahead of it),
// which lets the TMEM round trip overlap other work
__device__ __forceinline__ void tmem_issue18(uint32_t taddr, uint32_t (&r)[18]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];\n" : "=r"(r[16]), "=r"(r[17]) : "r"(taddr + 16) : "memory");
}
__device__ __forceinline__ void tmem_wait18(uint32_t (&r)[18]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                 "+r"(r[16]), "+r"(r[17])
               :
               : "memory");
}

__device__ __forceinline__ void named_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// mbarrier wait / arrive on a precomputed shared-memory address (keeps the hot loop free of address arithmetic)
// (A nanosleep between failed probes was tried per role -- issuer / loader / edge / compute warps, 0-100 ns -- and
// changed nothing but the cost of the extra instruction: 1.80 -> 2.00 ms per 1 M A-scans, profiles/r02/h_enc_sleep_sweep.log.
// The waits of this kernel are short and the spinning warps are few: spinning is not what limits it.)
__device__ __forceinline__ void mbar_wait_a(uint32_t addr, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(addr),
      "r"(parity)
      : "memory");
}
// (Probing a barrier early with mbarrier.test_wait -- the x / edge barrier of the next group at the end of a pass, the MMA
// barrier of the previous group at its top -- so that the blocking wait is skipped when the answer is already "complete"
// cost more than it hid: 1.715 -> 1.769 ms per 1 M A-scans, profiles/r02/z_bench_msc_early_probes.log.)
__device__ __forceinline__ void mbar_arrive_a(uint32_t addr) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(addr) : "memory");
}

template <bool PROBE>
__global__ void __launch_bounds__(ENC_THREADS, 1) k_msc_encoder_tc(MscEncArgs p) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar_conv[2][MAX_TILES], bar_full[2][MAX_TILES], bar_x[XS_SLOTS], bar_xe[XS_SLOTS], bar_e[XS_SLOTS], bar_l1, bar_l2;
  __shared__ __align__(16) uint4 EB[XS_SLOTS][ENC_COMPUTE / 32 * 2];   // conv1 vectors just outside every warp's 32 positions
  __shared__ uint32_t tmem_slot;

  const int S = p.S;
  const int tiles = S / 64;                        // M tiles per group of 2 A-scans (S % 64 == 0)
  const int tbuf = tiles * 48;                     // TMEM columns of one group: 16 (A operand) + 32 (accumulator) per tile
  unsigned char* W1S = smem;                       // shared_layer.0 / 32, resident: [S/8 chunks][128 rows][16 B]
  unsigned char* W2S = W1S + (size_t)S * 256;      // shared_layer.2, resident: [16 chunks][64 rows][16 B]
  unsigned char* A2 = W2S + 16384;                 // [S/8 chunks][128 rows][16 B] (+16 B skew per chunk);
                                                   // its head is reused as A3 [16 chunks][128 rows][16 B]
  unsigned char* BC = A2 + (size_t)(S / 8) * A2_LBO;   // [4 chunks][32 rows][16 B]
  __nv_bfloat16* XS = reinterpret_cast<__nv_bfloat16*>(BC + 2048);   // [XS_SLOTS][2][S + 16]
  const int xs_stride = S + XS_PAD;

  const int tid = threadIdx.x, lane = tid & 31, warp = uniform_warp_id();
  const int64_t nblocks = (p.A + 127) / 128;       // blocks of 128 A-scans, walked with stride gridDim.x
  constexpr int ngroups = 64;

  // ---- one-time setup
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  if (tid == 0) {
    for (int b = 0; b < 2; ++b)
      for (int t = 0; t < MAX_TILES; ++t) { mbar_init(&bar_conv[b][t], 1); mbar_init(&bar_full[b][t], 4); }   // 4 warps hand a tile over
    // an x slot is released by every compute warp (one arrival each) once its conv1 of the group is done
    for (int i = 0; i < XS_SLOTS; ++i) { mbar_init(&bar_x[i], 32); mbar_init(&bar_xe[i], tiles * 4); mbar_init(&bar_e[i], 1); }
    mbar_init(&bar_l1, 1);
    mbar_init(&bar_l2, 1);
    fence_mbar_init();
  }
  for (int i = tid; i < 128; i += ENC_THREADS) reinterpret_cast<uint4*>(BC)[i] = reinterpret_cast<const uint4*>(p.Bc)[i];
  for (int i = tid; i < S * 16; i += ENC_THREADS) reinterpret_cast<uint4*>(W1S)[i] = __ldg(reinterpret_cast<const uint4*>(p.W1p) + i);
  for (int i = tid; i < 1024; i += ENC_THREADS) reinterpret_cast<uint4*>(W2S)[i] = __ldg(reinterpret_cast<const uint4*>(p.W2p) + i);
  for (int i = tid; i < XS_SLOTS * 2 * xs_stride; i += ENC_THREADS) XS[i] = __float2bfloat16_rn(0.f);
  // conv1 weights as fp16 pairs, channels (2j, 2j+1), taps 0..2: kernel parameters (constant bank)
  auto cwh = [&](int j, int t) { return *reinterpret_cast<const __half2*>(&p.cw[j * 3 + t]); };
  auto cbh = [&](int j) { return *reinterpret_cast<const __half2*>(&p.cb[j]); };
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t idesc_conv = make_idesc_f16(128, 32);
  const int xparts = S / 8;
  const uint32_t a2_base = smem_u32(A2);

  // conv1 + ReLU at the sample x1 with neighbours x0, x2: 8 channels as 4 fp16 pairs
  auto conv1 = [&](float x0, float x1, float x2, uint32_t (&o)[4]) {
    const __half2 h0 = __float2half2_rn(x0), h1 = __float2half2_rn(x1), h2 = __float2half2_rn(x2);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const __half2 v = __hfma2_relu(cwh(j, 2), h2, __hfma2(cwh(j, 1), h1, __hfma2(cwh(j, 0), h0, cbh(j))));   // fma.rn.relu
      o[j] = *reinterpret_cast<const uint32_t*>(&v);
    }
  };
  auto ldx = [&](uint32_t addr) {
    uint16_t h;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(h) : "r"(addr));
    return __uint_as_float((uint32_t)h << 16);
  };
  const uint32_t xs_base = smem_u32(XS), eb_base = smem_u32(&EB[0][0]);
  const uint32_t slot_bytes = (uint32_t)(2 * xs_stride * 2);
  const uint32_t a_bar_x = smem_u32(&bar_x[0]), a_bar_xe = smem_u32(&bar_xe[0]), a_bar_e = smem_u32(&bar_e[0]),
                 a_bar_conv = smem_u32(&bar_conv[0][0]), a_bar_full = smem_u32(&bar_full[0][0]);

  // Groups are numbered globally over the blocks this CTA walks: G = it * 64 + g.  Every ring / double buffer
  // (conv1 buffers, TMEM accumulators, x slots) and every barrier parity is a function of G only, so the three
  // roles run through block boundaries without any extra synchronisation.
  unsigned long long tsum[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const bool probe = PROBE && p.dbg != nullptr && blockIdx.x == 0 && lane == 0;
  if (warp == ENC_COMPUTE / 32) {
    // ================= MMA issuer warp: conv2 MMAs of every group, decoupled from the compute warps ====
    // two TS-form MMAs per tile: K chunks (tap 0, tap 1) = TMEM columns 0..7 of the tile's operand rows, then
    // (tap 2, bias chunk) = columns 8..15
    const uint64_t bd0 = make_desc(smem_u32(BC), 512, 128);
    const uint64_t bd1 = bd0 + (uint64_t)((2 * 512) >> 4);
    const bool leader = elect_one();
    const long long i0 = probe ? clock64() : 0;
    // Every M tile is handed over by its own four compute warps and committed to its own barrier; the tiles are served in
    // index order, blocking on each tile's barrier.  (Serving them in whatever order they become ready -- a round robin of
    // non-blocking mbarrier.test_wait probes -- was tried: the probing warp's shared-memory traffic slowed the compute warps
    // down, 1.72 -> 2.16 ms per 1 M A-scans, profiles/r02/z_bench_msc_issuer_ab.log.)
    {
      uint32_t G = 0;
      // (Rotating the first tile served with the group, so that no team is always last, was tried too: 1.72 -> 1.81 ms.)
      for (int64_t blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
        for (int g = 0; g < ngroups; ++g, ++G) {
          const uint32_t buf = G & 1;
          const uint32_t a0t = tmem + buf * tbuf, d0 = a0t + tiles * 16;
#pragma unroll
          for (int T = 0; T < MAX_TILES; ++T) {
            if (T < tiles) {
              mbar_wait_a(a_bar_full + (buf * MAX_TILES + T) * 8, (G >> 1) & 1);
              if (leader) {
                tc_fence_after();
                mma_f16_ts(d0 + T * 32, a0t + T * 16, bd0, idesc_conv, 0u);
                mma_f16_ts(d0 + T * 32, a0t + T * 16 + 8, bd1, idesc_conv, 1u);
                mma_commit(&bar_conv[buf][T]);
              }
              __syncwarp();
            }
          }
        }
      }
    }
    if (probe) {
      tsum[0] = clock64() - i0;    // whole issue loop
#pragma unroll
      for (int i = 0; i < 5; ++i) p.dbg[warp * 8 + i] = tsum[i];
    }
  } else if (warp == ENC_COMPUTE / 32 + 1) {
    // ================= x loader warp =================
    // lane l owns the 16-byte parts l, l+32, l+64 of the 2*S/8 parts of a group.  cp.async writes them straight
    // into the ring slot; the slot's mbarrier (32 arrivals) completes when the copies of every lane have landed.
    int al_[3], part_[3];
    uint32_t dst[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int i = lane + 32 * j;
      al_[j] = i < 2 * xparts ? i / xparts : -1;
      part_[j] = i < 2 * xparts ? i - al_[j] * xparts : 0;
      dst[j] = smem_u32(XS) + (uint32_t)((al_[j] > 0 ? al_[j] : 0) * xs_stride + 8 + part_[j] * 8) * 2;
    }
    const __nv_bfloat16* xg = static_cast<const __nv_bfloat16*>(p.x);
    auto issue_x = [&](int64_t blk, int g, uint32_t G) {
      const uint32_t so = (uint32_t)(G & (XS_SLOTS - 1)) * slot_bytes;
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        if (al_[j] < 0) continue;
        const int64_t a = blk * 128 + 2 * g + al_[j];
        if (a < p.A) {
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst[j] + so), "l"(xg + a * S + part_[j] * 8) : "memory");
        } else {
          asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(dst[j] + so), "r"(0) : "memory");
        }
      }
      asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&bar_x[G & (XS_SLOTS - 1)])) : "memory");
    };
    // classic full / empty ring over the XS_SLOTS x slots: slot s is refilled for group G (G % XS_SLOTS == s) once
    // the compute warps have released it after conv1 of group G - XS_SLOTS (bar_xe[s], one completion per use).
    // Each barrier is waited on by exactly one side that can lag by at most one phase, so a parity wait can never
    // alias -- waiting on the consumers' hand-off barrier instead could, when this warp fell two phases behind.
    uint32_t G = 0;
    for (int64_t blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
      for (int g = 0; g < ngroups; ++g, ++G) {
        if (G >= XS_SLOTS) mbar_wait_a(a_bar_xe + (G & (XS_SLOTS - 1)) * 8, ((G / XS_SLOTS) - 1) & 1);
        issue_x(blk, g, G);
      }
    }
  } else if (warp == ENC_COMPUTE / 32 + 2) {
    // idle: only there to put the edge warp on the fourth scheduler
  } else if (warp == EDGE_WARP) {
    // ================= edge warp =================
    // A compute warp owns 32 consecutive positions and needs the conv1 vectors of the two positions just outside
    // (conv2 taps -1 / +1 of its first / last row).  Computing them inside the compute warps costs a full
    // instruction stream for two lanes; this otherwise idle warp produces all 2 x 20 of them per group as soon
    // as the group's x has landed, and publishes them through bar_e (which therefore also implies "x landed").
    int epos[2], exoff[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int e = lane + 32 * j;                          // entry = 2 * warp + side
      const int w = e >> 1, side = e & 1;
      const int P0 = 32 * w, al = P0 >= S ? 1 : 0, first = P0 - al * S;
      epos[j] = e < 2 * (ENC_COMPUTE / 32) && P0 < 2 * S ? (side ? first + 32 : first - 1) : -2;   // -2: no entry
      exoff[j] = (al * xs_stride + 8 + (epos[j] > -2 ? epos[j] : 0)) * 2;
    }
    for (int64_t blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
#pragma unroll 4
      for (int g = 0; g < ngroups; ++g) {                             // slot and parity depend on g alone (64 % 4 == 0)
        const uint32_t slot = g & (XS_SLOTS - 1);
        mbar_wait_a(a_bar_x + slot * 8, (g / XS_SLOTS) & 1);         // x of the group has landed
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          if (epos[j] > -2) {
            uint32_t o[4] = {0u, 0u, 0u, 0u};
            if (epos[j] >= 0 && epos[j] < S) {                       // beyond the A-scan conv2 pads with zeros
              const uint32_t xa = xs_base + slot * slot_bytes + (uint32_t)exoff[j];
              conv1(ldx(xa - 2), ldx(xa), ldx(xa + 2), o);
            }
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(eb_base + (slot * (ENC_COMPUTE / 32 * 2) + lane + 32 * j) * 16),
                         "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3])
                         : "memory");
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive_a(a_bar_e + slot * 8);
      }
    }
  } else {
    // ================= compute warps =================
    // thread = one position of the group: P = tid -> A-scan al = P / S, position pos = P % S; warp w holds the
    // rows of M tile w / 4, lane quarter w % 4 -- exactly the TMEM lanes a warp may access
    const bool c_active = tid < 2 * S;                            // (2*S <= ENC_COMPUTE for every supported S)
    const int c_al = tid >= S ? 1 : 0, c_pos = tid - c_al * S;
    const uint32_t c_xoff = (uint32_t)(c_al * xs_stride + 8 + c_pos) * 2;
    const uint32_t c_eoff = (uint32_t)(2 * warp + (lane == 31 ? 1 : 0)) * 16;
    const int q = warp & 3, T = warp >> 2;
    const uint32_t t_lane = (uint32_t)(q * 32) << 16;
    // loop-invariant addresses, pinned in registers (the empty asm makes them opaque: otherwise every use
    // re-derives them from the warp index and the kernel arguments, ~25 of the loop's issue slots)
    uint32_t tm_st[2] = {tmem + t_lane + (uint32_t)(T * 16), tmem + t_lane + (uint32_t)(tbuf + T * 16)};
    uint32_t tm_ld[2] = {tmem + t_lane + (uint32_t)(tiles * 16 + T * 32), tmem + t_lane + (uint32_t)(tbuf + tiles * 16 + T * 32)};
    uint32_t bar_conv_t = a_bar_conv + (uint32_t)T * 8, bar_full_t = a_bar_full + (uint32_t)T * 8;
    uint32_t xs_addr = xs_base + c_xoff, eb_addr = eb_base + c_eoff;
    asm volatile("" : "+r"(tm_st[0]), "+r"(tm_st[1]), "+r"(tm_ld[0]), "+r"(tm_ld[1]), "+r"(bar_conv_t), "+r"(bar_full_t),
                      "+r"(xs_addr), "+r"(eb_addr));
    const bool is_first = lane == 0, is_last = lane == 31;
    // this lane's output element of the conv epilogue (row q*32 + lane of tile T)
    const uint32_t e_off = (uint32_t)(c_pos >> 3) * A2_LBO + (uint32_t)c_al * 16 + (uint32_t)(c_pos & 7) * 2;

    // epilogue of one conv group: f[pos] = sum_c relu(y_c) (the 1/32 lives in W1) -> bf16 K-major operand of L1
    auto epi_issue = [&](uint32_t G, uint32_t (&r)[18]) {   // wait for the MMAs of group G, request this lane's row
      const uint32_t buf = G & 1;
      mbar_wait_a(bar_conv_t + buf * (MAX_TILES * 8), (G >> 1) & 1);
      tc_fence_after();
      if (c_active) tmem_issue18(tm_ld[buf], r);
    };
    auto epi_finish = [&](int g, uint32_t (&r)[18]) {
      if (c_active) {
        tmem_wait18(r);
        // four independent accumulation chains (|.| is a free source modifier): depth 6 instead of 10 dependent FADDs
        float f0 = __uint_as_float(r[16]) + fabsf(__uint_as_float(r[0])), f1 = __uint_as_float(r[17]) + fabsf(__uint_as_float(r[1]));
        float f2 = fabsf(__uint_as_float(r[2])) + fabsf(__uint_as_float(r[3])), f3 = fabsf(__uint_as_float(r[4])) + fabsf(__uint_as_float(r[5]));
#pragma unroll
        for (int c = 6; c < 16; c += 4) {
          f0 += fabsf(__uint_as_float(r[c]));
          f1 += fabsf(__uint_as_float(r[c + 1]));
          if (c + 2 < 16) { f2 += fabsf(__uint_as_float(r[c + 2])); f3 += fabsf(__uint_as_float(r[c + 3])); }
        }
        const __nv_bfloat16 fb = __float2bfloat16_rn((f0 + f1) + (f2 + f3));
        asm volatile("st.shared.u16 [%0], %1;" ::"r"(a2_base + e_off + (uint32_t)g * 32), "h"(*reinterpret_cast<const uint16_t*>(&fb)) : "memory");
      }
    };

    const bool team_active = T < tiles;                 // whole teams are active or idle (2 * S is a multiple of 128)
    uint32_t G = 0;
    int it = 0;
    for (int64_t blk = blockIdx.x; blk < nblocks; blk += gridDim.x, ++it) {
      const int64_t a0 = blk * 128;
      G += ngroups;
      // G = it * 64 + g and 64 is a multiple of every ring length, so buffer / slot indices and barrier parities
      // depend on g alone: unrolled by four they are compile-time constants of each copy of the body
      if (team_active) {
#pragma unroll 4
      for (int g = 0; g < ngroups; ++g) {
        const uint32_t buf = g & 1, slot = g & (XS_SLOTS - 1);
        const long long c0 = probe ? clock64() : 0;
        mbar_wait_a(a_bar_e + slot * 8, (g / XS_SLOTS) & 1);          // x has landed and the edge vectors are published
        // ---- conv1 + ReLU -> im2col row in tensor memory.  The operand columns are free: the MMAs of group G-2
        // completed before the epilogue of group G-2 ran.
        if (c_active) {
          const uint32_t xa = xs_addr + slot * slot_bytes;
          const float xm = ldx(xa - 2), x0 = ldx(xa), xp = ldx(xa + 2);
          uint32_t mid[4], lft[4], rgt[4];
          conv1(xm, x0, xp, mid);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            lft[j] = __shfl_up_sync(0xffffffffu, mid[j], 1);
            rgt[j] = __shfl_down_sync(0xffffffffu, mid[j], 1);
          }
          // warp edges: the left neighbour of lane 0 and the right neighbour of lane 31 live in other warps; the
          // edge warp has published their conv1 vectors (zeros beyond the A-scan)
          // branch-free: every lane loads one edge vector (its own warp's left one, lane 31 the right one) and
          // only lanes 0 / 31 select it -- a divergent branch for two lanes cost the whole warp ~20 issue slots
          {
            uint32_t e0, e1, e2, e3;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(e0), "=r"(e1), "=r"(e2), "=r"(e3)
                         : "r"(eb_addr + slot * (uint32_t)(ENC_COMPUTE / 32 * 2 * 16)));
            lft[0] = is_first ? e0 : lft[0]; lft[1] = is_first ? e1 : lft[1];
            lft[2] = is_first ? e2 : lft[2]; lft[3] = is_first ? e3 : lft[3];
            rgt[0] = is_last ? e0 : rgt[0]; rgt[1] = is_last ? e1 : rgt[1];
            rgt[2] = is_last ? e2 : rgt[2]; rgt[3] = is_last ? e3 : rgt[3];
          }
          asm volatile(
              "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
              ::"r"(tm_st[buf]), "r"(lft[0]), "r"(lft[1]), "r"(lft[2]), "r"(lft[3]), "r"(mid[0]),
              "r"(mid[1]), "r"(mid[2]), "r"(mid[3]), "r"(rgt[0]), "r"(rgt[1]), "r"(rgt[2]), "r"(rgt[3]), "r"(0x3C003C00u), "r"(0u),
              "r"(0u), "r"(0u)
              : "memory");
        }
        // the accumulator rows of the previous group are requested while the operand store drains
        uint32_t er[18];
        if (g > 0) epi_issue(g - 1, er);
        if (c_active) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        const long long c1 = probe ? clock64() : 0;
        tc_fence_before();
        const long long c1b = probe ? clock64() : 0;
        // every warp hands its 32 operand rows over by itself (the tile's barrier counts four arrivals): no warp waits
        // for its team mates (ncu: the 128-thread named barrier that used to sit here was 9 % of the compute warps' time)
        __syncwarp();
        if (lane == 0) {
          mbar_arrive_a(bar_full_t + buf * (MAX_TILES * 8));   // operand rows of this quarter of tile T are in tensor memory
          mbar_arrive_a(a_bar_xe + slot * 8);               // and this warp's share of the x slot goes back to the loader warp
        }
        const long long c2 = probe ? clock64() : 0;
        if (g > 0) epi_finish(g - 1, er);
        if (probe) {
          const long long c4 = clock64();
          tsum[0] += c1 - c0;    // x staging + conv1
          tsum[1] += c1b - c1;   // proxy fence
          tsum[2] += c2 - c1b;   // compute-warp barrier
          tsum[4] += c4 - c2;    // barrier release + epilogue of the previous group (incl. mbarrier wait)
        }
      }
      {
        uint32_t er[18];
        epi_issue(ngroups - 1, er);
        epi_finish(ngroups - 1, er);
      }
      }
      const long long l0 = probe ? clock64() : 0;
      // ---- Linear S -> 128 (+ReLU): A = A2, B = W1S, both resident; accumulator D1 reuses the conv columns
      fence_async_smem();
      tc_fence_before();
      named_sync(BAR_BLOCK, ENC_COMPUTE);
      if (warp == 0) {
        if (elect_one()) {
          tc_fence_after();
          const uint32_t a_addr = smem_u32(A2), b_addr = smem_u32(W1S);
          for (int ks = 0; ks < S / 16; ++ks)
            mma_bf16_ss(tmem + D1_COL, make_desc(a_addr + ks * 2 * A2_LBO, A2_LBO, 128),
                        make_desc(b_addr + ks * 2 * 2048, 2048, 128), make_idesc_bf16(128, H0), ks ? 1u : 0u);
          mma_commit(&bar_l1);
        }
        __syncwarp();
      }
      mbar_wait(&bar_l1, it & 1);
      tc_fence_after();
      // ---- epilogue 1: relu(D1 + b) -> bf16 -> A3 (head of A2: its MMAs are complete), K-major for the next GEMM
      {
        unsigned char* A3 = A2;
        const int r = q * 32 + lane;
        for (int n = (warp >> 2) * 16; n < H0; n += (ENC_COMPUTE / 128) * 16) {
          float v[16];
          tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + D1_COL + n, v);
          uint32_t pk[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            __nv_bfloat162 h2 = __floats2bfloat162_rn(fmaxf(v[2 * j] + __ldg(p.bl1 + n + 2 * j), 0.f),
                                                      fmaxf(v[2 * j + 1] + __ldg(p.bl1 + n + 2 * j + 1), 0.f));
            pk[j] = *reinterpret_cast<uint32_t*>(&h2);
          }
          *reinterpret_cast<uint4*>(A3 + (size_t)(n >> 3) * 2048 + r * 16) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          *reinterpret_cast<uint4*>(A3 + (size_t)((n >> 3) + 1) * 2048 + r * 16) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        }
      }
      fence_async_smem();
      tc_fence_before();
      named_sync(BAR_BLOCK, ENC_COMPUTE);
      // ---- Linear 128 -> 64
      if (warp == 0) {
        if (elect_one()) {
          tc_fence_after();
          const uint32_t a_addr = smem_u32(A2), b_addr = smem_u32(W2S);
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)
            mma_bf16_ss(tmem + D2_COL, make_desc(a_addr + ks * 2 * 2048, 2048, 128),
                        make_desc(b_addr + ks * 2 * 1024, 1024, 128), make_idesc_bf16(128, H1), ks ? 1u : 0u);
          mma_commit(&bar_l2);
        }
        __syncwarp();
      }
      mbar_wait(&bar_l2, it & 1);
      tc_fence_after();
      // ---- epilogue 2: relu(D2 + b) + position table -> h (fp32)
      {
        const int64_t a = a0 + q * 32 + lane;
        for (int n = (warp >> 2) * 16; n < H1; n += (ENC_COMPUTE / 128) * 16) {
          float v[16];
          tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + D2_COL + n, v);
          if (a < p.A) {
            const float* pr = p.pos + (a % p.Nset) * H1 + n;
            float4* dst = reinterpret_cast<float4*>(p.h + a * H1 + n);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bl2 + n) + j);
              const float4 t4 = __ldg(reinterpret_cast<const float4*>(pr) + j);
              dst[j] = make_float4(fmaxf(v[4 * j] + b4.x, 0.f) + t4.x, fmaxf(v[4 * j + 1] + b4.y, 0.f) + t4.y,
                                   fmaxf(v[4 * j + 2] + b4.z, 0.f) + t4.z, fmaxf(v[4 * j + 3] + b4.w, 0.f) + t4.w);
            }
          }
        }
      }
      // the next block's conv epilogues rewrite A2 and its MMAs rewrite the TMEM columns of D1 / D2.  A team hands
      // its first tile over without waiting for the other teams, so every warp must have finished reading D2
      // before any team moves on: one block-wide barrier per 128 A-scans (the hand-off itself then orders the
      // MMAs behind it: tcgen05 fence + team barrier + arrive)
      tc_fence_before();
      named_sync(BAR_BLOCK, ENC_COMPUTE);
      if (probe) tsum[5] += clock64() - l0;
    }
    if (probe) {
#pragma unroll
      for (int i = 0; i < 6; ++i) p.dbg[warp * 8 + i] = tsum[i];
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

uint16_t f2bf(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  const uint32_t r = 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)((u + r) >> 16);
}
float bf2f(uint16_t h) {
  uint32_t u = (uint32_t)h << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

}  // namespace

bool msc_encoder_tc_supported(int S, int h0, int h1) { return h0 == H0 && h1 == H1 && S % 64 == 0 && S >= 64 && S <= 320; }

// conv2 operand [4 chunks][32 rows][8] as fp16: rows 0..15 = channels (taps 0..2 in chunks 0..2, bias hi/lo in
// chunk 3), row 16 / 17 = hi / lo halves of the column sums (so that y16 + y17 = sum_c y_c), rows 18..31 = 0.
// The weights are first rounded to bf16 (the precision contract of the bf16 mode); every such value with an
// exponent >= -14 is exactly representable in fp16, whose 11-bit significand also makes the hi/lo pairs tighter.
void msc_pack_conv2(const float* w2 /*[16][8][3]*/, const float* b2 /*[16]*/, std::vector<uint16_t>& out) {
  out.assign(4 * 32 * 8, 0);
  auto f2h = [](float f) -> uint16_t { const __half h = __float2half_rn(f); uint16_t u; memcpy(&u, &h, 2); return u; };
  auto h2f = [](uint16_t u) -> float { __half_raw r; r.x = u; return __half2float(__half(r)); };
  auto at = [&](int chunk, int row, int e) -> uint16_t& { return out[((size_t)chunk * 32 + row) * 8 + e]; };
  float wsum[3][8] = {}, bsum = 0.f;
  for (int n = 0; n < 16; ++n) {
    for (int t = 0; t < 3; ++t)
      for (int ci = 0; ci < 8; ++ci) {
        const uint16_t h = f2h(bf2f(f2bf(w2[(n * 8 + ci) * 3 + t])));
        at(t, n, ci) = h;
        wsum[t][ci] += h2f(h);
      }
    const uint16_t hi = f2h(b2[n]);
    const uint16_t lo = f2h(b2[n] - h2f(hi));
    at(3, n, 0) = hi;
    at(3, n, 1) = lo;
    bsum += h2f(hi) + h2f(lo);
  }
  for (int t = 0; t < 3; ++t)
    for (int ci = 0; ci < 8; ++ci) {
      const uint16_t hi = f2h(wsum[t][ci]);
      at(t, 16, ci) = hi;
      at(t, 17, ci) = f2h(wsum[t][ci] - h2f(hi));
    }
  const uint16_t bhi = f2h(bsum);
  at(3, 16, 0) = bhi;
  at(3, 16, 1) = f2h(bsum - h2f(bhi));
}

void op_msc_encoder_tc(Ctx& c, const void* x, int x_dtype, int64_t A, int S, int Nset, const float* w1_host, const float* b1_host,
                       const void* Bc, const void* W1p, const float* bl1, const void* W2p, const float* bl2,
                       const float* pos, float* h) {
  if (c.dry) return;
  PAUT_CHECK(msc_encoder_tc_supported(S, H0, H1), PAUT_ERR_UNSUPPORTED, "msc encoder: unsupported signal length");
  PAUT_CHECK(x_dtype == PAUT_BF16, PAUT_ERR_INVALID, "msc encoder: input must be bf16 (cast fp32 first)");
  MscEncArgs p;
  p.x = x; p.x_dtype = x_dtype; p.A = A; p.S = S; p.Nset = Nset;
  for (int j = 0; j < 4; ++j) {
    auto pack = [](float lo, float hi) {
      const __half2 h = __floats2half2_rn(lo, hi);
      uint32_t u;
      memcpy(&u, &h, 4);
      return u;
    };
    p.cb[j] = pack(b1_host[2 * j], b1_host[2 * j + 1]);
    for (int t = 0; t < 3; ++t) p.cw[j * 3 + t] = pack(w1_host[(2 * j) * 3 + t], w1_host[(2 * j + 1) * 3 + t]);
  }
  p.Bc = static_cast<const __nv_bfloat16*>(Bc); p.W1p = static_cast<const __nv_bfloat16*>(W1p); p.bl1 = bl1;
  p.W2p = static_cast<const __nv_bfloat16*>(W2p); p.bl2 = bl2; p.pos = pos; p.h = h;
  const size_t smem = (size_t)S * 256 + 16384 + (size_t)(S / 8) * A2_LBO + 2048 + (size_t)XS_SLOTS * 2 * (S + XS_PAD) * 2;
  PAUT_CHECK((int)smem <= c.smem_optin, PAUT_ERR_UNSUPPORTED, "msc encoder: shared memory budget exceeded");
  static const bool debug = std::getenv("PAUT_ENC_DEBUG") != nullptr;
  void (*kern)(MscEncArgs) = debug ? k_msc_encoder_tc<true> : k_msc_encoder_tc<false>;
  smem_optin(c, kern);
  const int64_t nblocks = (A + 127) / 128;
  PAUT_CHECK(nblocks < (int64_t(1) << 24), PAUT_ERR_INVALID, "msc encoder: too many A-scans in one launch");
  const int64_t grid = nblocks < c.num_sms ? nblocks : c.num_sms;      // persistent: one CTA per SM
  p.dbg = nullptr;
  if (debug) PAUT_CUDA(cudaMalloc(&p.dbg, sizeof(unsigned long long) * 8 * (ENC_THREADS / 32)));
  kern<<<(unsigned)grid, ENC_THREADS, smem, c.stream>>>(p);
  c.launched("msc_encoder_tc");
  if (debug) {
    std::vector<unsigned long long> hbuf(8 * (ENC_THREADS / 32));
    PAUT_CUDA(cudaMemcpy(hbuf.data(), p.dbg, hbuf.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    cudaFree(p.dbg);
    fprintf(stderr, "[enc probe] per-group cycles (CTA 0): warp  conv1  fence  barrier  mma_issue  epilogue | linear stage per block\n");
    const unsigned long long nb0 = (unsigned long long)((nblocks + grid - 1) / grid), ng = 64 * nb0;   // CTA 0's share
    for (int w = 0; w <= ENC_COMPUTE / 32; w += (w < 18 ? 6 : 1))
      fprintf(stderr, "[enc probe] %4d %6llu %6llu %8llu %9llu %9llu | %llu\n", w, hbuf[w * 8] / ng, hbuf[w * 8 + 1] / ng,
              hbuf[w * 8 + 2] / ng, hbuf[w * 8 + 3] / ng, hbuf[w * 8 + 4] / ng, hbuf[w * 8 + 5] / nb0);
  }
}

}  // namespace paut
