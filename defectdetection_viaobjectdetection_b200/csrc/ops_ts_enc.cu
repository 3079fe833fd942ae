// Fused encoder of TwoStageDefectDetector for the bf16 mode (MultiScaleSignalEncoder.forward,
// two_stage_model.py:92-118): per A-scan, four branches
//     Conv1d 1->32 (k = 3/5/7/11) + BN + ReLU -> Conv1d 32->32 (same k) + BN + ReLU
// concatenated to 128 channels and averaged over the signal length.  ONE persistent kernel; nothing but the
// input samples (640 B per A-scan, through TMA) and the pooled 128-vector (512 B) touches HBM.
//
// Work units.  A CTA walks blocks of 16 A-scans.  Inside a block the A-scans are laid on a flat row axis with period
// Lp = S + 8 (S signal rows followed by 8 zero rows, the padding of every convolution); 16 * Lp is a multiple of 128,
// so a block is a whole number of 128-row M tiles (41 for S = 320), a tile touches at most two A-scans, and the
// tiles of consecutive blocks form one continuous row stream.
//
// BOTH convolutions run on the tensor cores:
//   stem (C_in = 1): the 11-sample window of a row is an im2col row of K = 16 (taps 0..10 = x[l-5 .. l+5], taps 11, 12
//          = 1.0 carrying the folded BN shift as a bf16 hi + lo pair, rest 0).  ONE tcgen05.mma 128 x 128 x 16 per
//          tile computes all four branches (smaller kernels are zero-padded, centred); rows outside an A-scan are
//          all-zero im2col rows, so their outputs are exactly zero -- the zero padding of the second convolution.
//          Its accumulator is read back once: cvt.rn.relu.f16x2 (ReLU + fp16 in one instruction) -> the fp16
//          activation rows [128 channels] in a shared-memory RING of 4 tiles (+ 8 mirrored rows at each end), in
//          the canonical K-major operand layout [branch][4 chunks][rows][16 B].
//   second convolutions: per tile 52 tcgen05.mma 128x32x16 (SS form, fp16 x fp16 -> fp32): branch b, tap t, K step
//          ks reads the ring through a descriptor advanced by (t - k/2) rows -- no im2col copy, no recomputed halo
//          -- against the resident folded weights; 4 x 32 accumulator columns, two accumulator buffers.
// Measured on the way here (profiles/r02): a version with the stems on HFMA2 (register-resident weights) issued 14.7 k
// warp instructions per tile and was bound by instruction issue at 3.3 IPC (4360 cycles per tile against 2224 for the
// MMAs); as an MMA the stem costs 64 tensor-pipe cycles and ~600 instructions per tile.
//
// Warp roles (16 warps):
//   TMA warp      : x of the next block -> shared memory (cp.async.bulk.tensor.2d, one box per A-scan, mbarrier
//                   complete_tx).
//   stem warps    : 4 warps, thread = row of the tile: build the im2col row (11 ld.shared.u16, 2 st.shared.v4), and
//                   one tile later turn the stem accumulator into ring rows.
//   MMA warp      : stem MMA of tile g, second-convolution MMAs of tile g - 2 (it needs the first rows of tile g - 1).
//   epilogue      : 8 warps (lane quarter x column half): TMEM -> registers (16x256b loads: a thread holds 4 rows x 8
//                   columns), + BN shift, ReLU, the 4 rows are added locally, a 3-level transposing butterfly leaves
//                   one column sum per lane; per-tile partial sums of the four lane quarters are combined in a fixed
//                   order and accumulated per A-scan in shared memory; the finished mean is written once per
//                   A-scan.  The summation order depends only on the A-scan's index modulo 16.
//
// Roofline: 17.0 MFLOP per A-scan on the tensor pipe; an SS-form 128x32x16 MMA is paced by the 4 KB A-tile fetch
// from shared memory (measured 40 cycles against the 16 of the pipe, tools/mma_probe2.py), which bounds this kernel
// at ~0.4 of the MMA rate at the clock it runs at.
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
constexpr int TE_MMA_WARP = 8, TE_TMA_WARP = 9;   // warps 10, 11 idle
constexpr int TE_STEM_WARP0 = 12;                 // warps 12-15: stem group, warp = TMEM lane quarter
constexpr int TE_STEM_WARPS = 4;
constexpr int TE_THREADS = (TE_STEM_WARP0 + TE_STEM_WARPS) * 32;   // 512
constexpr int TE_PAD = 5;                         // k = 11
constexpr int TE_SLOTS = 4;                       // ring of 4 tiles
constexpr int X_FULL = 0, X_EMPTY = 1, XA_FULL = 2, XA_EMPTY = 4, SACC_FULL = 6, SACC_EMPTY = 8, ACC_FULL = 10, ACC_EMPTY = 12,
              RING_FULL = 14, RING_EMPTY = 14 + TE_SLOTS, TE_NBAR = 14 + 2 * TE_SLOTS;     // barrier indices
constexpr int TE_MIRROR = 8;                      // mirrored rows at each end of the ring
constexpr int TE_RROWS = TE_MIRROR + TE_SLOTS * 128 + TE_MIRROR;   // 528
constexpr int TE_LBO_R = TE_RROWS * 16;           // chunk stride of the ring (bytes)
constexpr int TE_RING_BYTES = 16 * TE_LBO_R;      // 4 branches x 4 chunks
constexpr int TE_BLOCK_A = 16;
constexpr int TE_HALO = 8;
constexpr int TE_NTAPS = 3 + 5 + 7 + 11;          // 26 (branch, tap) pairs
constexpr int TE_W2_BYTES = TE_NTAPS * 2048;      // [26][4 chunks][32 rows][16 B]
constexpr int TE_WST_BYTES = 2 * 128 * 16;        // stem weights [2 chunks][128 rows][16 B], bf16
constexpr int TE_XA_BYTES = 2 * 128 * 16;         // one im2col tile [2 chunks][128 rows][16 B], bf16
constexpr int TE_XPAD = 64;                       // zero elements in front of / behind the x staging area
constexpr int TC_ACC = 0, TC_STEM = 256;          // TMEM columns: 2 x 128 second-convolution, 2 x 128 stem accumulators

struct TsEncArgs {
  const __nv_bfloat16* Wst;       // stem weights + BN shift as a [128 x 16] bf16 B operand, K-major chunks
  const __half* W2;               // second convolutions, BN scale folded, fp16 [26][4][32][8]
  const float* shift2;            // [128]
  float* feat;                    // [A][128] mean over the signal length
  long long A;
  long long nblk;                 // blocks of 16 A-scans
  int S, Lp;                      // signal length, row period S + 8
  int tpb;                        // tiles per block = 16 * Lp / 128
  int xs_stride;                  // elements between A-scans in the x staging buffer (multiple of 64)
  int xs_buf_bytes;               // the staging buffer
  int rows_per_ascan;             // rows of the [rows, W] tensor-map view one A-scan occupies (S / W)
  float invS;
  unsigned long long* dbg;        // optional cycle probe (PAUT_TS_DEBUG=1): role timings of CTA 0
};

__device__ __forceinline__ void named_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Parity wait; a failed probe puts the warp to sleep (nanosleep) before the next one, with a dead-lock guard: a
// protocol error traps instead of hanging the GPU.
template <int SLEEP_NS>
__device__ __forceinline__ void mbar_wait_g(uint32_t addr, uint32_t parity) {
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
// one arrival per warp: every lane has executed its fences, lane 0 arrives for the warp
__device__ __forceinline__ void warp_arrive(uint32_t bar, int lane) {
  __syncwarp();
  if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

// one box of the 2-D tensor map -> shared memory, completion counted in bytes on the mbarrier
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar)
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

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// relu + fp32 -> fp16 of two values in one instruction: lo in the low half
__device__ __forceinline__ uint32_t relu_pack_f16(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}

__global__ void __launch_bounds__(TE_THREADS, 1)
    k_ts_encoder(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ TsEncArgs p) {
  extern __shared__ __align__(128) unsigned char smem[];
  // all mbarriers in ONE array: their shared-window addresses are one pinned register + compile-time offsets (as separate
  // variables every use re-derived its address with S2R + LEA)
  __shared__ __align__(8) uint64_t bars[TE_NBAR];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float pool_s[2][128];          // running sums of the (at most two) open A-scans
  __shared__ __align__(16) float part_s[2][4][2][128];    // [tile parity][lane quarter][segment][column]

  unsigned char* W2S = smem;                              // resident weights of the second convolutions
  unsigned char* WST = W2S + TE_W2_BYTES;                 // stem weights (B operand)
  unsigned char* RING = WST + TE_WST_BYTES;               // fp16 activation rows of 4 tiles (+ mirrors)
  unsigned char* XA = RING + TE_RING_BYTES;               // [2] im2col tiles of the stem
  unsigned char* XS = XA + 2 * TE_XA_BYTES;               // x staging buffer (bf16)

  const int tid = threadIdx.x, lane = tid & 31, warp = uniform_warp_id();
  const int S = p.S, Lp = p.Lp, tpb = p.tpb;
  const int nb_local = (int)((p.nblk - blockIdx.x + gridDim.x - 1) / gridDim.x);
  const int nt = nb_local * tpb;                          // tiles of this CTA; the stem side runs one more (all zero)

  // ---- one-time setup
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  if (tid == 0) {
    mbar_init(&bars[X_FULL], 1); mbar_init(&bars[X_EMPTY], TE_STEM_WARPS);
    for (int a = 0; a < 2; ++a) {
      mbar_init(&bars[XA_FULL + a], TE_STEM_WARPS); mbar_init(&bars[XA_EMPTY + a], 1);
      mbar_init(&bars[SACC_FULL + a], 1); mbar_init(&bars[SACC_EMPTY + a], TE_STEM_WARPS);
      mbar_init(&bars[ACC_FULL + a], 1); mbar_init(&bars[ACC_EMPTY + a], TE_EPI_WARPS);
    }
    for (int s = 0; s < TE_SLOTS; ++s) { mbar_init(&bars[RING_FULL + s], TE_STEM_WARPS); mbar_init(&bars[RING_EMPTY + s], 1); }
    fence_mbar_init();
  }
  for (int i = tid; i < TE_W2_BYTES / 16; i += TE_THREADS)
    reinterpret_cast<uint4*>(W2S)[i] = __ldg(reinterpret_cast<const uint4*>(p.W2) + i);
  for (int i = tid; i < TE_WST_BYTES / 16; i += TE_THREADS)
    reinterpret_cast<uint4*>(WST)[i] = __ldg(reinterpret_cast<const uint4*>(p.Wst) + i);
  // ring (the rows in front of the first tile are the zero rows before the first A-scan), im2col tiles, x staging pads
  for (int i = tid; i < (TE_RING_BYTES + 2 * TE_XA_BYTES + p.xs_buf_bytes) / 16; i += TE_THREADS)
    reinterpret_cast<uint4*>(RING)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (tid < 128) { pool_s[0][tid] = 0.f; pool_s[1][tid] = 0.f; }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  uint32_t bar_u = smem_u32(bars);
  asm volatile("" : "+r"(bar_u));                     // opaque: keeps the address in a register instead of re-deriving it
  auto BA = [&](int idx) { return bar_u + (uint32_t)idx * 8u; };
  const uint32_t xs_base = smem_u32(XS), ring_base = smem_u32(RING), xa_base = smem_u32(XA);
  const bool probe = p.dbg != nullptr && blockIdx.x == 0 && lane == 0;
  unsigned long long pt[4] = {0, 0, 0, 0};

  if (warp == TE_TMA_WARP) {
    // ================= TMA producer: x of block i -> the staging buffer =================
    if (elect_one()) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap)) : "memory");
      for (int i = 0; i < nb_local; ++i) {
        if (i >= 1) mbar_wait_g<500>(BA(X_EMPTY), (i - 1) & 1);            // the im2col rows of block i - 1 are built
        mbar_expect_tx(BA(X_FULL), (uint32_t)(TE_BLOCK_A * S * 2));
        const long long a0 = ((long long)blockIdx.x + (long long)i * gridDim.x) * TE_BLOCK_A;
        const uint32_t dst0 = xs_base + TE_XPAD * 2;
#pragma unroll 1
        for (int j = 0; j < TE_BLOCK_A; ++j)     // rows beyond the volume are zero-filled by the TMA unit
          tma_load_2d(dst0 + (uint32_t)(j * p.xs_stride * 2), &tmap, 0, (int)((a0 + j) * p.rows_per_ascan), BA(X_FULL));
      }
    }
    __syncwarp();
  } else if (warp == TE_MMA_WARP) {
    // ================= MMA issuer =================
    const bool leader = elect_one();
    const uint32_t idesc2 = make_idesc_f16(128, 32), idesc_st = make_idesc_bf16(128, 128);
    const uint32_t desc_hi = (uint32_t)(128 >> 4) | (1u << 14);          // SBO = 128 B, descriptor version 1
    const uint32_t w_u = smem_u32(W2S) >> 4, wst_u = smem_u32(WST) >> 4, r_u = ring_base >> 4, xa_u = xa_base >> 4;   // 16-byte units
    for (int g = 0; g <= nt + 1; ++g) {
      if (g <= nt) {
        // ---- stem MMA of tile g: [128 rows x 16 taps] x [16 taps x 128 channels]
        const int b = g & 1;
        mbar_wait_g<0>(BA(XA_FULL + (b)), (g >> 1) & 1);
        if (g >= 2) mbar_wait_g<0>(BA(SACC_EMPTY + (b)), ((g >> 1) - 1) & 1);
        if (leader) {
          tc_fence_after();
          mma_bf16_ss2(tmem + (uint32_t)(TC_STEM + 128 * b), ((xa_u + (uint32_t)(b * (TE_XA_BYTES / 16))) & 0x3FFFu) | (128u << 16), desc_hi,
                       (wst_u & 0x3FFFu) | (128u << 16), desc_hi, idesc_st, 0u);
          commit_a(BA(XA_EMPTY + (b)));
          commit_a(BA(SACC_FULL + (b)));
        }
        __syncwarp();
      }
      if (g >= 2) {
        // ---- second convolutions of tile u: its window reaches 5 rows into tiles u - 1 and u + 1
        const int u = g - 2, slot = u & (TE_SLOTS - 1), acc = u & 1;
        const long long q0 = probe ? clock64() : 0;
        mbar_wait_g<0>(BA(RING_FULL + ((u + 1) & (TE_SLOTS - 1))), ((u + 1) >> 2) & 1);
        const long long q1 = probe ? clock64() : 0;
        if (u >= 2) mbar_wait_g<0>(BA(ACC_EMPTY + (acc)), ((u >> 1) - 1) & 1);    // the epilogue drained this accumulator
        const long long q2 = probe ? clock64() : 0;
        if (leader) {
          tc_fence_after();
          const uint32_t a_u = r_u + (uint32_t)(TE_MIRROR + 128 * slot);       // ring row of the tile's first output row
          auto branch = [&](int b, int K, int woff) {
            const uint32_t d = tmem + (uint32_t)(TC_ACC + acc * 128 + b * 32);
            uint32_t accum = 0u;
#pragma unroll
            for (int t = 0; t < K; ++t) {
#pragma unroll
              for (int ks = 0; ks < 2; ++ks) {
                // A: rows (t - K/2) .. +127 of chunks (2ks, 2ks+1) of branch b;  B: [32 x 16] block of tap t
                const uint32_t ad = ((a_u + (uint32_t)((b * 4 + ks * 2) * TE_RROWS + t - K / 2)) & 0x3FFFu) | ((uint32_t)TE_RROWS << 16);
                const uint32_t bd = ((w_u + (uint32_t)((woff + t) * 128 + ks * 64)) & 0x3FFFu) | (32u << 16);
                mma_bf16_ss2(d, ad, desc_hi, bd, desc_hi, idesc2, accum);
                accum = 1u;
              }
            }
          };
          branch(0, 3, 0);
          branch(1, 5, 3);
          branch(2, 7, 8);
          branch(3, 11, 15);
          commit_a(BA(ACC_FULL + (acc)));
          commit_a(BA(RING_EMPTY + ((u + TE_SLOTS - 1) & (TE_SLOTS - 1))));        // tile u - 1 has no reader left
        }
        __syncwarp();
        if (probe) { const long long q3 = clock64(); pt[0] += q1 - q0; pt[1] += q2 - q1; pt[2] += q3 - q2; pt[3] += 1; }
      }
    }
    if (probe) { p.dbg[0] = pt[0]; p.dbg[1] = pt[1]; p.dbg[2] = pt[2]; p.dbg[3] = pt[3]; }
  } else if (warp >= TE_STEM_WARP0) {
    // ================= stem group: thread = row i of the tile =================
    const int qw = warp - TE_STEM_WARP0;
    const int i = qw * 32 + lane;
    const uint32_t t_lane = (uint32_t)(qw * 32) << 16;
    // stem accumulator of tile u -> relu -> fp16 rows of ring slot u % 4 (+ the mirrored rows at the ring's ends)
    auto stem_epilogue = [&](int u) {
      const int b = u & 1, slot = u & (TE_SLOTS - 1), lap = u >> 2;
      const long long s0 = probe ? clock64() : 0;
      mbar_wait_g<40>(BA(SACC_FULL + (b)), (u >> 1) & 1);
      // the slot's previous tile (u - 4) was last read by the MMAs of tile u - 3; slot 3 also owns the mirrored rows
      // in front of the ring, which tile 0 reads as its zero rows: ring_empty[3] carries an extra first phase
      if (slot == TE_SLOTS - 1) mbar_wait_g<100>(BA(RING_EMPTY + (slot)), lap & 1);
      else if (lap >= 1) mbar_wait_g<100>(BA(RING_EMPTY + (slot)), (lap - 1) & 1);
      const long long s1 = probe ? clock64() : 0;
      tc_fence_after();
      const uint32_t row = ring_base + (uint32_t)(TE_MIRROR + 128 * slot + i) * 16;
      const bool mir_hi = slot == 0 && i < TE_MIRROR;                      // first rows of the ring -> behind its end
      const bool mir_lo = slot == TE_SLOTS - 1 && i >= 128 - TE_MIRROR;    // last rows of the ring -> in front of it
#pragma unroll 1
      for (int cc = 0; cc < 4; ++cc) {
        uint32_t r[32];
        tmem_ld32(tmem + t_lane + (uint32_t)(TC_STEM + 128 * b + 32 * cc), r);
        if (cc == 3) {                                                     // accumulator drained
          tc_fence_before();
          warp_arrive(BA(SACC_EMPTY + (b)), lane);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint32_t w[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) w[e] = relu_pack_f16(__uint_as_float(r[8 * j + 2 * e]), __uint_as_float(r[8 * j + 2 * e + 1]));
          const uint32_t a = row + (uint32_t)((4 * cc + j) * TE_LBO_R);
          st_shared_v4(a, w[0], w[1], w[2], w[3]);
          if (mir_hi) st_shared_v4(a + TE_SLOTS * 128 * 16, w[0], w[1], w[2], w[3]);
          if (mir_lo) st_shared_v4(a - TE_SLOTS * 128 * 16, w[0], w[1], w[2], w[3]);
        }
      }
      fence_async_smem();                                                  // generic-proxy stores -> tensor-core operand reads
      warp_arrive(BA(RING_FULL + (slot)), lane);
      if (probe && qw == 0) { pt[0] += s1 - s0; pt[1] += clock64() - s1; pt[2] += 1; }
    };
    int a_loc = 0, l = i;                                                  // row i of tile T of the block: A-scan, position
    int cur_blk = -1;
    for (int t = 0; t <= nt; ++t) {
      const int i_blk = t / tpb, T = t - i_blk * tpb;
      const int b = t & 1;
      bool valid = false;
      uint32_t xa = 0;
      if (t < nt) {
        if (i_blk != cur_blk) {
          cur_blk = i_blk;
          a_loc = 0; l = i;                                                // (Lp >= 136: the first 128 rows are A-scan 0)
          mbar_wait_g<200>(BA(X_FULL), i_blk & 1);
        }
        const long long a_blk = ((long long)blockIdx.x + (long long)i_blk * gridDim.x) * TE_BLOCK_A;
        valid = l < S && a_blk + a_loc < p.A;
        xa = xs_base + (uint32_t)(TE_XPAD + a_loc * p.xs_stride + l - TE_PAD) * 2;
      }
      // ---- im2col row: taps 0..10 = x[l-5 .. l+5] (bf16 as stored), taps 11, 12 = 1.0 (BN shift hi + lo), rest 0
      uint32_t w[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
      if (valid) {
        uint16_t h[11];
#pragma unroll
        for (int k = 0; k < 11; ++k) asm volatile("ld.shared.u16 %0, [%1];" : "=h"(h[k]) : "r"(xa + 2 * k) : "memory");
#pragma unroll
        for (int k = 0; k < 5; ++k) w[k] = (uint32_t)h[2 * k] | ((uint32_t)h[2 * k + 1] << 16);
        w[5] = (uint32_t)h[10] | 0x3F800000u;
        w[6] = 0x00003F80u;
      }
      if (t >= 2) mbar_wait_g<40>(BA(XA_EMPTY + (b)), ((t >> 1) - 1) & 1);        // the stem MMA of tile t - 2 has read the buffer
      const uint32_t dst = xa_base + (uint32_t)(b * TE_XA_BYTES + i * 16);
      st_shared_v4(dst, w[0], w[1], w[2], w[3]);
      st_shared_v4(dst + 128 * 16, w[4], w[5], w[6], w[7]);
      fence_async_smem();
      warp_arrive(BA(XA_FULL + (b)), lane);
      if (t < nt) {
        if (T == tpb - 1) warp_arrive(BA(X_EMPTY), lane);                     // the block's x has been read
        l += 128;                                                          // the same row of the next tile
        if (l >= Lp) { l -= Lp; ++a_loc; }
      }
      if (t >= 1) stem_epilogue(t - 1);
    }
    stem_epilogue(nt);
    if (probe && qw == 0) { p.dbg[16] = pt[0]; p.dbg[17] = pt[1]; p.dbg[18] = pt[2]; }
  } else if (warp < TE_EPI_WARPS) {
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
    // transposing butterfly over the lanes with equal lane % 4 (xor 16, 8, 4): 8 values -> 1
    auto bfly = [&](float (&v)[8]) {
#pragma unroll
      for (int m = 4; m >= 1; m >>= 1) {
        const bool hi = (lane & (4 * m)) != 0;
#pragma unroll
        for (int k = 0; k < m; ++k) {
          const float keep = hi ? v[m + k] : v[k];
          const float send = hi ? v[k] : v[m + k];
          v[k] = keep + __shfl_xor_sync(0xffffffffu, send, 4 * m);
        }
      }
    };
    int a_first = 0, l0 = 0;                           // geometry of the tile: first A-scan, its position at the first row
    for (int g = 0; g < nt; ++g) {
      const int i_blk = g / tpb, T = g - i_blk * tpb;
      const int acc = g & 1;
      if (T == 0) { a_first = 0; l0 = 0; }
      const long long a_blk = ((long long)blockIdx.x + (long long)i_blk * gridDim.x) * TE_BLOCK_A;
      // the warp's 32 rows are all signal rows of the first A-scan (warp-uniform, the common case)
      const bool uni0 = l0 + 32 * q + 31 < S && a_blk + a_first < p.A;
      const bool uni1 = l0 + 32 * q >= Lp && l0 + 32 * q + 31 - Lp < S && a_blk + a_first + 1 < p.A;   // ... of the second one
      const long long e0 = probe ? clock64() : 0;
      mbar_wait_g<60>(BA(ACC_FULL + (acc)), (g >> 1) & 1);
      const long long e1 = probe ? clock64() : 0;
      tc_fence_after();
      float* pq = &part_s[acc][q][0][64 * hf];
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        uint32_t ra[16], rb[16];
        const uint32_t col = (uint32_t)(TC_ACC + acc * 128 + 64 * hf + 32 * cc);
        tmem_ld_16x256b_x4(tmem + ((uint32_t)(q * 32) << 16) + col, ra);          // rows i8, i8 + 8
        tmem_ld_16x256b_x4(tmem + ((uint32_t)(q * 32 + 16) << 16) + col, rb);     // rows i8 + 16, i8 + 24
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (cc == 1) {                                                   // accumulator drained: tile g + 2 may start
          tc_fence_before();
          warp_arrive(BA(ACC_EMPTY + (acc)), lane);
        }
        float s0[8];
        if (uni0 || uni1) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const float sh = shv[cc][2 * j + e];
              s0[2 * j + e] = (fmaxf(__uint_as_float(ra[4 * j + e]) + sh, 0.f) + fmaxf(__uint_as_float(ra[4 * j + 2 + e]) + sh, 0.f)) +
                              (fmaxf(__uint_as_float(rb[4 * j + e]) + sh, 0.f) + fmaxf(__uint_as_float(rb[4 * j + 2 + e]) + sh, 0.f));
            }
          bfly(s0);
          pq[32 * cc + my_col] = uni0 ? s0[0] : 0.f;
          pq[128 + 32 * cc + my_col] = uni0 ? 0.f : s0[0];
        } else {
          // rows i8 + 8k of the quarter: valid row of the first / of a second A-scan (or a zero row between them)
          float f0[4], f1[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int lr = l0 + 32 * q + i8 + 8 * k;
            const int seg = lr >= Lp ? 1 : 0, l = lr - seg * Lp;
            const bool valid = l < S && a_blk + a_first + seg < p.A;
            f0[k] = valid && seg == 0 ? 1.f : 0.f;
            f1[k] = valid && seg == 1 ? 1.f : 0.f;
          }
          float s1[8];
#pragma unroll
          for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const float sh = shv[cc][2 * j + e];
              const float v0 = fmaxf(__uint_as_float(ra[4 * j + e]) + sh, 0.f), v1 = fmaxf(__uint_as_float(ra[4 * j + 2 + e]) + sh, 0.f);
              const float v2 = fmaxf(__uint_as_float(rb[4 * j + e]) + sh, 0.f), v3 = fmaxf(__uint_as_float(rb[4 * j + 2 + e]) + sh, 0.f);
              s0[2 * j + e] = (f0[0] * v0 + f0[1] * v1) + (f0[2] * v2 + f0[3] * v3);
              s1[2 * j + e] = (f1[0] * v0 + f1[1] * v1) + (f1[2] * v2 + f1[3] * v3);
            }
          bfly(s0);
          bfly(s1);
          pq[32 * cc + my_col] = s0[0];
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
      l0 += 128;                                                         // next tile (Lp >= 136: at most one A-scan further)
      if (l0 >= Lp) { l0 -= Lp; ++a_first; }
      if (probe) { pt[0] += e1 - e0; pt[1] += clock64() - e1; pt[2] += 1; }
    }
    if (probe && warp == 0) { p.dbg[8] = pt[0]; p.dbg[9] = pt[1]; p.dbg[10] = pt[2]; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
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

bool ts_encoder_supported(int S, int d_model) { return d_model == 128 && S % 16 == 0 && S >= 128 && S <= 448; }

uint16_t f2bf_bits(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  return (uint16_t)((u + 0x7fffu + ((u >> 16) & 1u)) >> 16);
}
float bf2f_bits(uint16_t h) {
  uint32_t u = (uint32_t)h << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

// Host-side packing (BatchNorm already folded by the caller):
//   w1[b]: [k_b][32] stem weights, sh1[b]: [32];  w2[b]: [k_b][32 in][32 out]
//   Wst: the stem as a [128 x 16] bf16 B operand ([2 chunks][128 rows][8]): row 32 b + co, K index = window tap
//        (branch b's k_b taps centred in the 11-tap window), taps 11 / 12 = hi / lo halves of the folded shift
//   W2 : fp16 [26 (branch, tap)][4 chunks][32 rows (co)][8 (ci)]
void ts_encoder_pack(const float* const* w1, const float* const* sh1, const float* const* w2, std::vector<uint16_t>& Wst,
                     std::vector<uint16_t>& W2) {
  const int taps[4] = {3, 5, 7, 11};
  Wst.assign((size_t)2 * 128 * 8, 0);
  W2.assign((size_t)TE_NTAPS * 4 * 32 * 8, 0);
  auto st = [&](int n, int k) -> uint16_t& { return Wst[((size_t)(k >> 3) * 128 + n) * 8 + (k & 7)]; };
  int woff = 0;
  for (int b = 0; b < 4; ++b) {
    for (int co = 0; co < 32; ++co) {
      for (int t = 0; t < taps[b]; ++t) st(32 * b + co, TE_PAD - taps[b] / 2 + t) = f2bf_bits(w1[b][t * 32 + co]);
      const uint16_t hi = f2bf_bits(sh1[b][co]);
      st(32 * b + co, 11) = hi;
      st(32 * b + co, 12) = f2bf_bits(sh1[b][co] - bf2f_bits(hi));
    }
    for (int t = 0; t < taps[b]; ++t)
      for (int ci = 0; ci < 32; ++ci)
        for (int co = 0; co < 32; ++co)
          W2[(((size_t)(woff + t) * 4 + ci / 8) * 32 + co) * 8 + ci % 8] = f2h_bits(w2[b][((size_t)t * 32 + ci) * 32 + co]);
    woff += taps[b];
  }
}

void op_ts_encoder(Ctx& c, const void* x_bf16, int64_t A, int S, const void* Wst, const void* W2, const float* shift2, float* feat) {
  if (c.dry) return;
  PAUT_CHECK(ts_encoder_supported(S, 128), PAUT_ERR_UNSUPPORTED, "two-stage encoder: unsupported signal length");
  PAUT_CHECK((reinterpret_cast<uintptr_t>(x_bf16) & 15) == 0, PAUT_ERR_INVALID, "two-stage encoder: x must be 16-byte aligned");
  TsEncArgs p;
  p.Wst = static_cast<const __nv_bfloat16*>(Wst);
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
  const size_t smem = (size_t)TE_W2_BYTES + TE_WST_BYTES + TE_RING_BYTES + 2 * TE_XA_BYTES + (size_t)p.xs_buf_bytes;
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
    fprintf(stderr, "[ts probe] per tile (CTA 0, %llu tiles) | mma: wait_ring_full %llu wait_acc_empty %llu issue %llu | "
                    "epilogue w0: wait_acc_full %llu work %llu | stem group w0, stem epilogue: wait %llu work %llu\n",
            h[3], per(h[0], h[3]), per(h[1], h[3]), per(h[2], h[3]), per(h[8], h[10]), per(h[9], h[10]), per(h[16], h[18]),
            per(h[17], h[18]));
  }
  c.launched("ts_encoder");
}

}  // namespace paut
