// Fused encoder of TwoStageDefectDetector for the bf16 mode (MultiScaleSignalEncoder.forward,
// two_stage_model.py:92-118): per A-scan, four branches
//     Conv1d 1->32 (k = 3/5/7/11) + BN + ReLU -> Conv1d 32->32 (same k) + BN + ReLU
// concatenated to 128 channels and averaged over the signal length.  ONE persistent kernel; nothing but the
// input samples (640 B per A-scan, through TMA) and the pooled 128-vector (512 B) touches HBM.
//
// Work units.  A CTA walks blocks of 16 A-scans.  Inside a block the A-scans are laid on a flat row axis with period
// Lp = S + 8 (S signal rows followed by 8 zero rows, the padding of every convolution); 16 * Lp is a multiple of 128,
// so a block is a whole number of 128-row M tiles (41 for S = 320) and a tile touches at most two A-scans.
//
// Warp roles (18 warps):
//   TMA warp    : x of the NEXT block -> shared memory (cp.async.bulk.tensor.2d, one box per A-scan, mbarrier
//                 complete_tx), double buffered.
//   stem warps  : two groups of 5 warps, group g takes every other tile.  A thread owns one of the 138 input rows
//                 of the tile's window (128 rows + 5 on each side), runs the four C_in = 1 stem convolutions in packed
//                 fp16 (HFMA2, weights and folded BN shift in the constant bank = kernel parameters) and stores the
//                 fp16 row [128 channels] into the stage buffer in the canonical K-major operand layout
//                 [branch][4 chunks][144 rows][16 B].  Rows outside an A-scan are stored as zeros.
//   MMA warp    : per tile 52 tcgen05.mma 128x32x16 (SS form, fp16 x fp16 -> fp32): branch b, tap t, K step ks reads
//                 the SAME stage buffer through a descriptor advanced by (5 - k/2 + t) rows -- no im2col copy --
//                 against the resident folded weights; 4 x 32 accumulator columns, two accumulator buffers in TMEM.
//   epilogue    : 4 warps: TMEM -> registers, + BN shift, ReLU, and the pooled sum over the tile's rows with a
//                 transposing butterfly (31 shuffles per 32 columns); per-tile partial sums of the four lane quarters
//                 are combined in a fixed order and accumulated per A-scan in shared memory; the finished mean is
//                 written once per A-scan.  The summation order depends only on the A-scan's index modulo 16.
//
// Roofline: 17.0 MFLOP per A-scan on the tensor pipe; an SS-form 128x32x16 MMA is paced by the 4 KB A-tile fetch
// from shared memory (measured 40 cycles against the 16 of the pipe), which bounds this kernel at ~0.4 of the MMA
// rate at the clock it runs at.
#include <cuda.h>
#include <cuda_fp16.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "tc_common.cuh"

namespace paut {

using namespace tc;

namespace {

constexpr int TE_EPI_WARPS = 8;                   // warps 0-7: warp w = TMEM lane quarter w % 4, column half w / 4
constexpr int TE_EPI = TE_EPI_WARPS * 32;
constexpr int TE_MMA_WARP = 8, TE_TMA_WARP = 9;
constexpr int TE_STEM_WARP0 = 10;
constexpr int TE_STEM_WARPS = 8;                  // stem warp sw: chunk sw % 4 (8 channels) of branches {3, 0} (sw < 4) or {2, 1}
constexpr int TE_STEM = TE_STEM_WARPS * 32;
constexpr int TE_THREADS = (TE_STEM_WARP0 + TE_STEM_WARPS) * 32;   // 576
constexpr int TE_PAD = 5;                         // k = 11
constexpr int TE_ROWS = 128 + 2 * TE_PAD;         // 138 input rows per tile
constexpr int TE_BROWS = 160;                     // rows per chunk of a stage buffer (5 row blocks of 32)
constexpr int TE_STAGES = 3;
constexpr int TE_BLOCK_A = 16;
constexpr int TE_HALO = 8;
constexpr int TE_LBO_A = TE_BROWS * 16;           // chunk stride of the A operand (bytes)
constexpr int TE_STAGE_BYTES = 16 * TE_LBO_A;     // 4 branches x 4 chunks
constexpr int TE_NTAPS = 3 + 5 + 7 + 11;          // 26 (branch, tap) pairs
constexpr int TE_W2_BYTES = TE_NTAPS * 2048;      // [26][4 chunks][32 rows][16 B]
constexpr int TE_XPAD = 64;                       // zero elements in front of / behind the x staging area

struct TsEncArgs {
  const uint32_t* sw;             // stem weights (BN scale folded) as fp16 pairs: [(branch, tap)][j] = channels (2j, 2j+1)
  const uint32_t* sb;             // stem shift (folded bias) as fp16 pairs [4][16]
  const __half* W2;               // second convolutions, BN scale folded, fp16 [26][4][32][8]
  const float* shift2;            // [128]
  float* feat;                    // [A][128] mean over the signal length
  long long A;
  long long nblk;                 // blocks of 16 A-scans
  int S, Lp;                      // signal length, row period S + 8
  int tpb;                        // tiles per block = 16 * Lp / 128
  int xs_stride;                  // elements between A-scans in the x staging buffer (multiple of 64)
  int xs_buf_bytes;               // one staging buffer
  int rows_per_ascan;             // rows of the [rows, W] tensor-map view one A-scan occupies (S / W)
  float invS;
  unsigned long long* dbg;        // optional cycle probe (PAUT_TS_DEBUG=1): role timings of CTA 0
};

__device__ __forceinline__ void named_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Parity wait that does not burn issue slots: a failed probe puts the warp to sleep (nanosleep) before the next one
// -- a spinning warp issues ~0.6 instructions per cycle and takes them from the working warps of its scheduler (ncu
// of the first version: 37 % of all issued instructions were wait loops).  With a dead-lock guard: a protocol error
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
    asm volatile("nanosleep.u32 %0;" ::"r"(SLEEP_NS));
    if (spin > (1u << 22)) __trap();
  }
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

// one box of the 2-D tensor map -> shared memory, completion counted in bytes on the mbarrier
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}

// TMEM -> registers, 16 lanes x 32 columns: thread t holds rows t/4 and t/4 + 8 of the 16-lane window, columns
// 8j + 2(t%4) + {0, 1}, j = 0..3:  r[4j + 0, 1] = row t/4, r[4j + 2, 3] = row t/4 + 8   (the mma C-fragment layout)
__device__ __forceinline__ void tmem_ld_16x256b_x4(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

// Stem convolution of one branch, one chunk of 8 channels (4 fp16 pairs), one row: K taps with the weights in
// registers (w[t][j]); xh[i] = sample l - 5 + i of the A-scan in both halves.  The last tap carries the ReLU
// (fma.rn.relu).  vmask zeroes rows that are not signal rows (bitwise: garbage inputs cannot leak).
template <int K, bool MASK>
__device__ __forceinline__ void stem_chunk(const __half2 (&w)[K][4], const __half2 (&sh)[4], const __half2 (&xh)[11],
                                           uint32_t dst, uint32_t vmask) {
  uint32_t o[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    __half2 acc = sh[j];
#pragma unroll
    for (int t = 0; t < K - 1; ++t) acc = __hfma2(w[t][j], xh[TE_PAD - K / 2 + t], acc);
    acc = __hfma2_relu(w[K - 1][j], xh[TE_PAD + K / 2], acc);
    o[j] = *reinterpret_cast<const uint32_t*>(&acc);
    if (MASK) o[j] &= vmask;
  }
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3]) : "memory");
}

template <int K>
__device__ __forceinline__ void load_stem_weights(const TsEncArgs& p, int woff, int b, int c, __half2 (&w)[K][4], __half2 (&sh)[4]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t s = __ldg(p.sb + b * 16 + c * 4 + j);
    sh[j] = *reinterpret_cast<const __half2*>(&s);
#pragma unroll
    for (int t = 0; t < K; ++t) {
      const uint32_t v = __ldg(p.sw + (woff + t) * 16 + c * 4 + j);
      w[t][j] = *reinterpret_cast<const __half2*>(&v);
    }
  }
}

// One stem warp: chunk c of branches (KA taps at weight offset WA, branch BA) and (KB, WB, BB), weights in registers
// for the whole kernel.  Lane = row of a 32-row block of the tile's 138-row window.
template <int KA, int WA, int BA, int KB, int WB, int BB>
__device__ __forceinline__ void stem_loop(const TsEncArgs& p, int c, int lane, int nt_local, uint32_t xs_base, uint32_t abuf_base,
                                          uint64_t* full, uint64_t* empty, uint64_t* x_full, uint64_t* x_empty, unsigned long long* dbg) {
  __half2 wa[KA][4], sa[4], wb[KB][4], sbb[4];
  load_stem_weights<KA>(p, WA, BA, c, wa, sa);
  load_stem_weights<KB>(p, WB, BB, c, wb, sbb);
  const int S = p.S, Lp = p.Lp, tpb = p.tpb;
  int cur_blk = -1;
  unsigned long long pt0 = 0, pt1 = 0, pt2 = 0;
  for (int g = 0; g < nt_local; ++g) {
    const int i_blk = g / tpb, T = g - i_blk * tpb;
    const int stage = g % TE_STAGES;
    const uint32_t use = (uint32_t)(g / TE_STAGES);
    const long long s0 = dbg ? clock64() : 0;
    if (i_blk != cur_blk) {
      if (cur_blk >= 0) {                                              // this warp has read the last x of the block
        __syncwarp();
        if (lane == 0) mbar_arrive(&x_empty[cur_blk & 1]);
      }
      cur_blk = i_blk;
      mbar_wait_g<200>(&x_full[i_blk & 1], (i_blk >> 1) & 1);
      // bf16 -> fp16 in place, once per block (16 x S samples over the 256 stem threads), so that the per-row loads
      // below need no conversion; the zero padding between the A-scans is the same bit pattern in both formats
      {
        const uint32_t xb0 = xs_base + (uint32_t)(i_blk & 1) * p.xs_buf_bytes + TE_XPAD * 2;
        const int st = (int)(threadIdx.x) - TE_STEM_WARP0 * 32;
        const int wpa = S / 2;                                           // 32-bit words per A-scan
        for (int i = st; i < TE_BLOCK_A * wpa; i += TE_STEM) {
          const int a = i / wpa, w = i - a * wpa;
          const uint32_t addr = xb0 + (uint32_t)(a * p.xs_stride * 2 + w * 4);
          uint32_t v;
          asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
          const __half2 h = __floats2half2_rn(__uint_as_float(v << 16), __uint_as_float(v & 0xffff0000u));
          asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(*reinterpret_cast<const uint32_t*>(&h)) : "memory");
        }
        named_sync(2, TE_STEM);
      }
    }
    if (use > 0) mbar_wait_g<100>(&empty[stage], (use - 1) & 1);            // the MMAs that read this buffer are done
    const long long s1 = dbg ? clock64() : 0;
    // first row of the window: flat row 128 T - 5 of the block -> A-scan a0 (may be -1), position l0
    const int fa = 128 * T - TE_PAD + Lp;                              // >= 0
    const int a1 = (int)((unsigned)fa / (unsigned)Lp);
    const int a0 = a1 - 1, l0 = fa - a1 * Lp;
    const long long a_blk = ((long long)blockIdx.x + (long long)i_blk * gridDim.x) * TE_BLOCK_A;
    const uint32_t xb = xs_base + (uint32_t)(i_blk & 1) * p.xs_buf_bytes;
    const uint32_t sbuf = abuf_base + (uint32_t)stage * TE_STAGE_BYTES;
#pragma unroll 1
    for (int rb = 0; rb < (TE_ROWS + 31) / 32; ++rb) {
      const int i = rb * 32 + lane;                                    // row of the window
      int a_loc = a0, l = l0 + i;
      while (l >= Lp) { l -= Lp; ++a_loc; }
      const bool valid = i < TE_ROWS && a_loc >= 0 && a_loc < TE_BLOCK_A && l < S && a_blk + a_loc < p.A;
      const int aa = valid ? a_loc : 0, ll = valid ? l : 0;
      const uint32_t xa = xb + (uint32_t)(TE_XPAD + aa * p.xs_stride + ll - TE_PAD) * 2;
      __half2 xh[11];
#pragma unroll
      for (int k = 0; k < 11; ++k) {
        uint16_t h;
        asm volatile("ld.shared.u16 %0, [%1];" : "=h"(h) : "r"(xa + 2 * k) : "memory");
        xh[k] = __half2half2(__ushort_as_half(h));
      }
      const uint32_t dst = sbuf + (uint32_t)i * 16;
      if (__all_sync(0xffffffffu, valid)) {                            // the common case: no masking instructions
        stem_chunk<KA, false>(wa, sa, xh, dst + (uint32_t)((BA * 4 + c) * TE_LBO_A), 0xffffffffu);
        stem_chunk<KB, false>(wb, sbb, xh, dst + (uint32_t)((BB * 4 + c) * TE_LBO_A), 0xffffffffu);
      } else {
        const uint32_t vmask = valid ? 0xffffffffu : 0u;
        stem_chunk<KA, true>(wa, sa, xh, dst + (uint32_t)((BA * 4 + c) * TE_LBO_A), vmask);
        stem_chunk<KB, true>(wb, sbb, xh, dst + (uint32_t)((BB * 4 + c) * TE_LBO_A), vmask);
      }
    }
    fence_async_smem();                                                // generic-proxy stores -> tensor-core operand reads
    __syncwarp();
    if (lane == 0) mbar_arrive(&full[stage]);
    if (dbg) { pt0 += s1 - s0; pt1 += clock64() - s1; pt2 += 1; }
  }
  if (dbg) { dbg[0] = pt0; dbg[1] = pt1; dbg[2] = pt2; }
  if (cur_blk >= 0) {
    __syncwarp();
    if (lane == 0) mbar_arrive(&x_empty[cur_blk & 1]);
  }
}

__global__ void __launch_bounds__(TE_THREADS, 1)
    k_ts_encoder(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ TsEncArgs p) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t full[TE_STAGES], empty[TE_STAGES], acc_full[2], acc_empty[2], x_full[2], x_empty[2];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float pool_s[2][128];          // running sums of the (at most two) open A-scans
  __shared__ __align__(16) float part_s[2][4][2][128];    // [tile parity][lane quarter][segment][column]

  unsigned char* W2S = smem;                              // resident weights of the second convolutions
  unsigned char* ABUF = W2S + TE_W2_BYTES;                // [TE_STAGES] stage buffers
  unsigned char* XS = ABUF + TE_STAGES * TE_STAGE_BYTES;  // [2] x staging buffers (bf16)

  const int tid = threadIdx.x, lane = tid & 31, warp = uniform_warp_id();
  const int S = p.S, Lp = p.Lp, tpb = p.tpb;
  const int nb_local = (int)((p.nblk - blockIdx.x + gridDim.x - 1) / gridDim.x);
  const int nt_local = nb_local * tpb;

  // ---- one-time setup
  if (warp == 0) tmem_alloc(&tmem_slot, 256);
  if (tid == 0) {
    for (int s = 0; s < TE_STAGES; ++s) { mbar_init(&full[s], TE_STEM_WARPS); mbar_init(&empty[s], 1); }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], TE_EPI_WARPS);
      mbar_init(&x_full[a], 1); mbar_init(&x_empty[a], TE_STEM_WARPS);
    }
    fence_mbar_init();
  }
  for (int i = tid; i < TE_W2_BYTES / 16; i += TE_THREADS)
    reinterpret_cast<uint4*>(W2S)[i] = __ldg(reinterpret_cast<const uint4*>(p.W2) + i);
  for (int i = tid; i < 2 * p.xs_buf_bytes / 16; i += TE_THREADS) reinterpret_cast<uint4*>(XS)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (tid < 128) { pool_s[0][tid] = 0.f; pool_s[1][tid] = 0.f; }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t xs_base = smem_u32(XS), abuf_base = smem_u32(ABUF);
  const bool probe = p.dbg != nullptr && blockIdx.x == 0 && lane == 0;
  unsigned long long pt[4] = {0, 0, 0, 0};

  if (warp == TE_TMA_WARP) {
    // ================= TMA producer: x of block i -> staging buffer i & 1 =================
    if (elect_one()) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap)) : "memory");
      for (int i = 0; i < nb_local; ++i) {
        const int xb = i & 1;
        if (i >= 2) mbar_wait_g<1000>(&x_empty[xb], ((i >> 1) - 1) & 1);       // every stem warp is done with block i - 2
        mbar_expect_tx(&x_full[xb], (uint32_t)(TE_BLOCK_A * S * 2));
        const long long a0 = ((long long)blockIdx.x + (long long)i * gridDim.x) * TE_BLOCK_A;
        const uint32_t dst0 = xs_base + (uint32_t)xb * p.xs_buf_bytes + TE_XPAD * 2;
#pragma unroll 1
        for (int j = 0; j < TE_BLOCK_A; ++j)     // rows beyond the volume are zero-filled by the TMA unit
          tma_load_2d(dst0 + (uint32_t)(j * p.xs_stride * 2), &tmap, 0, (int)((a0 + j) * p.rows_per_ascan), &x_full[xb]);
      }
    }
    __syncwarp();
  } else if (warp == TE_MMA_WARP) {
    // ================= MMA issuer =================
    const bool leader = elect_one();
    const uint32_t idesc = make_idesc_f16(128, 32);
    const uint32_t desc_hi = (uint32_t)(128 >> 4) | (1u << 14);          // SBO = 128 B, descriptor version 1
    const uint32_t w_u = smem_u32(W2S) >> 4;                             // 16-byte units
    for (int g = 0; g < nt_local; ++g) {
      const int stage = g % TE_STAGES, acc = g & 1;
      const uint32_t use = (uint32_t)(g / TE_STAGES);
      const long long q0 = probe ? clock64() : 0;
      if (g >= 2) mbar_wait_g<40>(&acc_empty[acc], ((g >> 1) - 1) & 1);      // the epilogue drained this accumulator
      const long long q1 = probe ? clock64() : 0;
      mbar_wait_g<40>(&full[stage], use & 1);                                // the stem rows of the tile are stored
      const long long q2 = probe ? clock64() : 0;
      if (leader) {
        tc_fence_after();
        const uint32_t a_u = (abuf_base + (uint32_t)stage * TE_STAGE_BYTES) >> 4;
        auto branch = [&](int b, int K, int woff) {
          const uint32_t d = tmem + (uint32_t)(acc * 128 + b * 32);
          uint32_t accum = 0u;
#pragma unroll
          for (int t = 0; t < K; ++t) {
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
              // A: rows (5 - K/2 + t) .. +127 of chunks (2ks, 2ks+1) of branch b;  B: [32 x 16] block of tap t
              const uint32_t ad = ((a_u + (uint32_t)((b * 4 + ks * 2) * TE_BROWS + TE_PAD - K / 2 + t)) & 0x3FFFu) |
                                  ((uint32_t)TE_BROWS << 16);
              const uint32_t bd = ((w_u + (uint32_t)((woff + t) * 128 + ks * 64)) & 0x3FFFu) | (32u << 16);
              mma_bf16_ss2(d, ad, desc_hi, bd, desc_hi, idesc, accum);
              accum = 1u;
            }
          }
        };
        branch(0, 3, 0);
        branch(1, 5, 3);
        branch(2, 7, 8);
        branch(3, 11, 15);
        mma_commit(&empty[stage]);
        mma_commit(&acc_full[acc]);
      }
      __syncwarp();
      if (probe) { const long long q3 = clock64(); pt[0] += q1 - q0; pt[1] += q2 - q1; pt[2] += q3 - q2; pt[3] += 1; }
    }
    if (probe) { p.dbg[0] = pt[0]; p.dbg[1] = pt[1]; p.dbg[2] = pt[2]; p.dbg[3] = pt[3]; }
  } else if (warp >= TE_STEM_WARP0) {
    // ================= stem warps =================
    const int sw = warp - TE_STEM_WARP0, c = sw & 3;
    unsigned long long* sd = probe && (sw == 0 || sw == 4) ? p.dbg + 16 + 2 * sw : nullptr;
    if (sw < 4) stem_loop<11, 15, 3, 3, 0, 0>(p, c, lane, nt_local, xs_base, abuf_base, full, empty, x_full, x_empty, sd);
    else stem_loop<7, 8, 2, 5, 3, 1>(p, c, lane, nt_local, xs_base, abuf_base, full, empty, x_full, x_empty, sd);
  } else {
    // ================= epilogue: warp = (TMEM lane quarter q, column half hf) =================
    // A thread holds 4 rows x 8 columns of every 32-column chunk (16x256b loads): + BN shift, ReLU, the 4 rows are
    // added locally, then a 3-level transposing butterfly over the 8 lanes that hold the same columns leaves ONE
    // column sum of the warp's 32 rows in every lane (column 8 (i / 2) + 2 (lane % 4) + i % 2 with i = lane / 4).
    const int q = warp & 3, hf = warp >> 2;
    const int t4 = lane & 3, i8 = lane >> 2;
    float shv[2][8];                                   // BN shift of this thread's columns (fixed for the kernel)
#pragma unroll
    for (int cc = 0; cc < 2; ++cc)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        shv[cc][2 * j] = __ldg(p.shift2 + 64 * hf + 32 * cc + 8 * j + 2 * t4);
        shv[cc][2 * j + 1] = __ldg(p.shift2 + 64 * hf + 32 * cc + 8 * j + 2 * t4 + 1);
      }
    const int my_col = 8 * (i8 >> 1) + 2 * t4 + (i8 & 1);
    for (int g = 0; g < nt_local; ++g) {
      const int i_blk = g / tpb, T = g - i_blk * tpb;
      const int acc = g & 1;
      const long long a_blk = ((long long)blockIdx.x + (long long)i_blk * gridDim.x) * TE_BLOCK_A;
      // geometry of the tile: first A-scan, its position at the tile's first row; a second A-scan starts at row Lp - l0
      const int t0 = 128 * T;
      const int a_first = (int)((unsigned)t0 / (unsigned)Lp), l0 = t0 - a_first * Lp;
      // the warp's 32 rows are all signal rows of the first A-scan (warp-uniform, the common case)
      const bool uni0 = l0 + 32 * q + 31 < S && a_blk + a_first < p.A;
      bool m0[4], m1[4];                               // rows i8 + 8k of the quarter: valid row of segment 0 / 1
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int lr = l0 + 32 * q + i8 + 8 * k;
        const int seg = lr >= Lp ? 1 : 0, l = lr - seg * Lp;
        const bool valid = l < S && a_blk + a_first + seg < p.A;
        m0[k] = valid && seg == 0;
        m1[k] = valid && seg == 1;
      }
      const long long e0 = probe ? clock64() : 0;
      mbar_wait_g<60>(&acc_full[acc], (g >> 1) & 1);
      const long long e1 = probe ? clock64() : 0;
      tc_fence_after();
      float* pq = &part_s[acc][q][0][64 * hf];
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        uint32_t ra[16], rb[16];
        const uint32_t col = (uint32_t)(acc * 128 + 64 * hf + 32 * cc);
        tmem_ld_16x256b_x4(tmem + ((uint32_t)(q * 32) << 16) + col, ra);          // rows i8, i8 + 8
        tmem_ld_16x256b_x4(tmem + ((uint32_t)(q * 32 + 16) << 16) + col, rb);     // rows i8 + 16, i8 + 24
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (cc == 1) {                                                   // accumulator drained: tile g + 2 may start
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[acc]);
        }
        float s0[8], s1[8];
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const float sh = shv[cc][2 * j + e];
            const float v0 = fmaxf(__uint_as_float(ra[4 * j + e]) + sh, 0.f), v1 = fmaxf(__uint_as_float(ra[4 * j + 2 + e]) + sh, 0.f);
            const float v2 = fmaxf(__uint_as_float(rb[4 * j + e]) + sh, 0.f), v3 = fmaxf(__uint_as_float(rb[4 * j + 2 + e]) + sh, 0.f);
            if (uni0) {
              s0[2 * j + e] = (v0 + v1) + (v2 + v3);
              s1[2 * j + e] = 0.f;
            } else {
              s0[2 * j + e] = ((m0[0] ? v0 : 0.f) + (m0[1] ? v1 : 0.f)) + ((m0[2] ? v2 : 0.f) + (m0[3] ? v3 : 0.f));
              s1[2 * j + e] = ((m1[0] ? v0 : 0.f) + (m1[1] ? v1 : 0.f)) + ((m1[2] ? v2 : 0.f) + (m1[3] ? v3 : 0.f));
            }
          }
        // transposing butterfly over the lanes with equal lane % 4 (xor 16, 8, 4): 8 values -> 1
        auto bfly = [&](float (&v)[8]) {
#pragma unroll
          for (int m = 4; m >= 1; m >>= 1) {
            const bool hi = (lane & (4 * m)) != 0;
#pragma unroll
            for (int i = 0; i < m; ++i) {
              const float keep = hi ? v[m + i] : v[i];
              const float send = hi ? v[i] : v[m + i];
              v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4 * m);
            }
          }
        };
        bfly(s0);
        pq[32 * cc + my_col] = s0[0];
        if (uni0) {
          pq[128 + 32 * cc + my_col] = 0.f;
        } else {
          bfly(s1);
          pq[128 + 32 * cc + my_col] = s1[0];
        }
      }
      named_sync(1, TE_EPI);
      // ---- combine the four lane quarters (fixed order) and accumulate per A-scan: threads 0..127 = the columns of the
      // tile's first A-scan, threads 128..255 = the columns of a second A-scan that starts in this tile
      {
        const int c = tid & 127, sg = tid >> 7;
        const float s = ((part_s[acc][0][sg][c] + part_s[acc][1][sg][c]) + part_s[acc][2][sg][c]) + part_s[acc][3][sg][c];
        const int par0 = a_first & 1;
        if (sg == 0) {
          const float sum0 = pool_s[par0][c] + s;
          const bool flush0 = l0 < S && S - 1 - l0 <= 127;               // the first A-scan's last row is in this tile
          if (flush0) {
            if (a_blk + a_first < p.A) p.feat[(a_blk + a_first) * 128 + c] = sum0 * p.invS;
            pool_s[par0][c] = 0.f;
          } else {
            pool_s[par0][c] = sum0;
          }
        } else if (l0 + 127 >= Lp) {
          pool_s[par0 ^ 1][c] += s;                                      // (S >= 128: a second A-scan cannot end here)
        }
      }
      if (probe) { pt[0] += e1 - e0; pt[1] += clock64() - e1; pt[2] += 1; }
    }
    if (probe && warp == 0) { p.dbg[8] = pt[0]; p.dbg[9] = pt[1]; p.dbg[10] = pt[2]; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

uint16_t f2h_bits(float f) {
  const __half h = __float2half_rn(f);
  uint16_t u;
  memcpy(&u, &h, 2);
  return u;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// the driver API is resolved at run time (libpaut.so does not link libcuda: it must load on a machine without a GPU)
EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      f = nullptr;
    return reinterpret_cast<EncodeTiledFn>(f);
  }();
  return fn;
}

}  // namespace

bool ts_encoder_supported(int S, int d_model) { return d_model == 128 && S % 16 == 0 && S >= 128 && S <= 512; }

// Host-side packing (BatchNorm already folded by the caller):
//   w1[b]: [k_b][32] stem weights, sh1[b]: [32];  w2[b]: [k_b][32 in][32 out], sh2: [128]
void ts_encoder_pack(const float* const* w1, const float* const* sh1, const float* const* w2, std::vector<uint32_t>& sw,
                     std::vector<uint32_t>& sb, std::vector<uint16_t>& W2) {
  const int taps[4] = {3, 5, 7, 11};
  sw.assign(TE_NTAPS * 16, 0);
  sb.assign(4 * 16, 0);
  W2.assign((size_t)TE_NTAPS * 4 * 32 * 8, 0);
  int woff = 0;
  for (int b = 0; b < 4; ++b) {
    for (int j = 0; j < 16; ++j) {
      sb[b * 16 + j] = (uint32_t)f2h_bits(sh1[b][2 * j]) | ((uint32_t)f2h_bits(sh1[b][2 * j + 1]) << 16);
      for (int t = 0; t < taps[b]; ++t)
        sw[(woff + t) * 16 + j] =
            (uint32_t)f2h_bits(w1[b][t * 32 + 2 * j]) | ((uint32_t)f2h_bits(w1[b][t * 32 + 2 * j + 1]) << 16);
    }
    for (int t = 0; t < taps[b]; ++t)
      for (int ci = 0; ci < 32; ++ci)
        for (int co = 0; co < 32; ++co)
          W2[(((size_t)(woff + t) * 4 + ci / 8) * 32 + co) * 8 + ci % 8] = f2h_bits(w2[b][((size_t)t * 32 + ci) * 32 + co]);
    woff += taps[b];
  }
}

void op_ts_encoder(Ctx& c, const void* x_bf16, int64_t A, int S, const uint32_t* sw_dev, const uint32_t* sb_dev,
                   const void* W2, const float* shift2, float* feat) {
  if (c.dry) return;
  PAUT_CHECK(ts_encoder_supported(S, 128), PAUT_ERR_UNSUPPORTED, "two-stage encoder: unsupported signal length");
  PAUT_CHECK((reinterpret_cast<uintptr_t>(x_bf16) & 15) == 0, PAUT_ERR_INVALID, "two-stage encoder: x must be 16-byte aligned");
  TsEncArgs p;
  p.sw = sw_dev; p.sb = sb_dev;
  p.W2 = static_cast<const __half*>(W2); p.shift2 = shift2; p.feat = feat;
  p.A = A; p.nblk = (A + TE_BLOCK_A - 1) / TE_BLOCK_A;
  p.S = S; p.Lp = S + TE_HALO; p.tpb = TE_BLOCK_A * p.Lp / 128;
  p.xs_stride = (S + 16 + 63) / 64 * 64;
  p.xs_buf_bytes = (2 * TE_XPAD + TE_BLOCK_A * p.xs_stride) * 2;
  p.invS = 1.0f / (float)S;
  // x [A, S] bf16 viewed as [A * S/W rows, W] with W <= 256 (the box limit): one box = one A-scan
  const int W = S <= 256 ? S : S / 2;
  p.rows_per_ascan = S / W;
  EncodeTiledFn enc = encode_tiled();
  PAUT_CHECK(enc != nullptr, PAUT_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  CUtensorMap tmap;
  const cuuint64_t gdim[2] = {(cuuint64_t)W, (cuuint64_t)(A * p.rows_per_ascan)};
  const cuuint64_t gstride[1] = {(cuuint64_t)W * 2};
  const cuuint32_t box[2] = {(cuuint32_t)W, (cuuint32_t)p.rows_per_ascan};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(x_bf16), gdim, gstride, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  PAUT_CHECK(r == CUDA_SUCCESS, PAUT_ERR_CUDA, "cuTensorMapEncodeTiled failed (two-stage encoder input)");
  const size_t smem = (size_t)TE_W2_BYTES + (size_t)TE_STAGES * TE_STAGE_BYTES + 2 * (size_t)p.xs_buf_bytes;
  smem_optin(c, k_ts_encoder);
  cudaFuncAttributes fa;
  PAUT_CUDA(cudaFuncGetAttributes(&fa, k_ts_encoder));
  PAUT_CHECK(smem + fa.sharedSizeBytes <= (size_t)c.smem_optin, PAUT_ERR_UNSUPPORTED, "two-stage encoder: shared memory budget exceeded");
  const long long grid = p.nblk < c.num_sms ? p.nblk : c.num_sms;     // persistent: one CTA per SM
  static const bool debug = std::getenv("PAUT_TS_DEBUG") != nullptr;
  p.dbg = nullptr;
  if (debug) { PAUT_CUDA(cudaMalloc(&p.dbg, 32 * sizeof(unsigned long long))); PAUT_CUDA(cudaMemset(p.dbg, 0, 32 * 8)); }
  k_ts_encoder<<<(unsigned)grid, TE_THREADS, smem, c.stream>>>(tmap, p);
  if (debug) {
    unsigned long long h[32];
    PAUT_CUDA(cudaMemcpy(h, p.dbg, sizeof(h), cudaMemcpyDeviceToHost));
    cudaFree(p.dbg);
    auto per = [](unsigned long long v, unsigned long long n) { return n ? v / n : 0ull; };
    fprintf(stderr, "[ts probe] per tile (CTA 0, %llu tiles) | mma: wait_acc_empty %llu wait_full %llu issue %llu | "
                    "epilogue w0: wait_acc_full %llu work %llu | stem w0 (k11+k3): wait %llu work %llu | stem w4 (k7+k5): wait %llu work %llu\n",
            h[3], per(h[0], h[3]), per(h[1], h[3]), per(h[2], h[3]), per(h[8], h[10]), per(h[9], h[10]), per(h[16], h[18]),
            per(h[17], h[18]), per(h[24], h[26]), per(h[25], h[26]));
  }
  c.launched("ts_encoder");
}

}  // namespace paut
