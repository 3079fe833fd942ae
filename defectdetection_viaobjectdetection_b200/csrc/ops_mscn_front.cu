// Fused front end of MultiSignalClassifier_N for the bf16 mode (NN_models.py:227-234): per A-scan
//     Conv1d 1->8 k3 + ReLU -> Conv1d 8->16 k3 + ReLU -> y - depthwise Conv1d(16ch, k11)(y) -> mean over the channels
// = f [S], the input of shared_layer.  ONE persistent kernel, every contraction on tcgen05; nothing but x (TMA) and f
// touches HBM.  Same row geometry as the two-stage encoder (ops_ts_enc.cu): blocks of 16 A-scans on a flat row axis
// with period Lp = S + 8 (the 8 zero rows are the padding of every convolution), 128-row M tiles, and the tiles of
// consecutive blocks form one continuous row stream.
//
//   stage 1  conv1 as ONE MMA 128 x 16 x 16 per tile: the im2col row of a position is [x[l-1], x[l], x[l+1], 1, 1, 0..]
//            (bf16; the two ones carry the bias as a hi + lo pair), rows outside an A-scan are all zero, so their
//            outputs are exactly zero.  Epilogue: cvt.rn.relu.f16x2 -> the 8 channels of a row are ONE 16-byte chunk
//            of ring 1.
//   stage 2  conv2 as TWO MMAs 128 x 16 x 16: with a chunk stride (LBO) of 16 bytes the second K chunk of a row is
//            the NEXT row of ring 1, so (tap 0 | tap 1) and (tap 2 | zero weights) are descriptors that start one row
//            before / after the tile -- no im2col copy.  Epilogue: + bias, ReLU -> fp16, rows outside an A-scan
//            forced to zero (the zero padding of the background extractor) -> 2 chunks per row of ring 2.
//   stage 3  "x - depthwise(x), mean over channels" is one 16 -> 1 convolution with 11 taps,
//            f[l] = sum_t sum_c M[c][t] y_c[l + t - 5] - mean(b_bg),  M[c][t] = ([t == 5] - w_bg[c][t]) / 16:
//            11 MMAs 128 x 16 x 16 over ring 2 through descriptors advanced by (t - 5) rows (only output column 0 is
//            used; an SS-form MMA costs the same 39 cycles for N = 16 as for N = 32: it is paced by the A-tile fetch).
//            Epilogue: column 0 -> f (fp32) -> HBM.
//
// (Round 2 also tried the contraction the other way round -- channel mix G[l][t] = sum_c y_c[l] M[c][t] as ONE MMA, then the
// diagonal gather f[l] = sum_t G[l + t - 5][t] by the worker threads from a transposed fp16 ring: 4 MMAs per tile instead of
// 14, parity-green, and SLOWER, 9.4 against 7.5 ms per 1 M A-scans.  The in-kernel probe explains both versions
// (profiles/r02/3_mscn_front_v2_experiment.log): a worker warp runs ~700 instructions per pass at ~9 cycles each -- one
// thread per row, one warp per scheduler and CTA -- so the kernel is bound by the worker warps' instruction latency, not by
// the tensor pipe (546 cycles per tile here, 156 there), not by sleeping waits (spinning changes nothing) and not by the
// number of hand-overs per pass (merging the five stages of a pass into one TMEM round trip + one fence changed nothing).
// What it needs is the instruction diet the MSC encoder got in round 1 -- constant ring indices through unrolling, no
// 64-bit index arithmetic or wait-loop bookkeeping in the pass -- before the 4-MMA formulation can pay off.)
//
// One CTA = 4 worker warps (thread = row of the tile: im2col rows and the three epilogues), one MMA warp, one TMA
// warp; ~50 KB of shared memory and 128 TMEM columns, so FOUR CTAs share an SM and hide each other's hand-over
// latencies (the chain of one tile is seven steps long).
#include <cuda.h>
#include <cuda_fp16.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <type_traits>

#include "common.cuh"
#include "tc_common.cuh"

namespace paut {

using namespace tc;

namespace {

constexpr int MF_WORKERS = 4;                      // warps 0-3: worker warp = TMEM lane quarter
constexpr int MF_MMA_WARP = 4, MF_TMA_WARP = 5;
constexpr int MF_THREADS = 6 * 32;
constexpr int MF_SLOTS = 4, MF_MIRROR = 8;
constexpr int MF_RROWS = MF_MIRROR + MF_SLOTS * 128 + MF_MIRROR;   // 528 rows per ring chunk
constexpr int MF_R1_BYTES = MF_RROWS * 16;         // ring 1: one chunk (8 channels) per row
constexpr int MF_R2_BYTES = 2 * MF_RROWS * 16;     // ring 2: two chunks (16 channels) per row
constexpr int MF_XA_BYTES = 2 * 128 * 16;          // im2col tile [2 chunks][128 rows][16 B]
constexpr int MF_W_BYTES = (1 + 2 + 11) * 512;     // B operands: conv1, conv2 (2), stencil taps (11): [2 chunks][16 rows][16 B]
constexpr int MF_BLOCK_A = 16, MF_HALO = 8, MF_XPAD = 64;
// barrier indices
constexpr int X_FULL = 0, X_EMPTY = 1, XA_FULL = 2, XA_EMPTY = 4, D_FULL = 6, D_EMPTY = 12, R_FULL = 18, R_EMPTY = 26, MF_NBAR = 34;
constexpr int TC_D1 = 0, TC_D2 = 32, TC_D3 = 64;   // TMEM columns: 2 x 16 per stage

struct MscnArgs {
  const uint16_t* W;              // packed B operands (MF_W_BYTES): conv1 (bf16), conv2 a / b (fp16), 11 stencil taps (fp16)
  float b2[16];                   // conv1d.2 bias
  float f_const;                  // mean over the channels of the background extractor's bias
  float* f;                       // [A][S]
  long long A, nblk;
  int S, Lp, tpb, xs_stride, xs_buf_bytes, rows_per_ascan;
};

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
    if (spin > (1u << 22)) __trap();               // dead-lock guard: a protocol error traps instead of hanging
  }
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint32_t relu_pack_f16(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t tmem_ld1(uint32_t taddr) {
  uint32_t r;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];\n" : "=r"(r) : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  return r;
}

__global__ void __launch_bounds__(MF_THREADS, 4)
    k_mscn_front(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ MscnArgs p) {
  extern __shared__ __align__(128) unsigned char smem[];
  // all mbarriers in ONE array: their shared-window addresses are then one pinned register + compile-time offsets (as
  // separate variables every use re-derived its address with S2R + LEA, ~5 instructions and an S2R latency each)
  __shared__ __align__(8) uint64_t bars[MF_NBAR];
  __shared__ uint32_t tmem_slot;

  unsigned char* WS = smem;                               // B operands
  unsigned char* R1 = WS + MF_W_BYTES;                    // ring 1 (conv1 output, fp16, 8 channels per row)
  unsigned char* R2 = R1 + MF_R1_BYTES;                   // ring 2 (conv2 output, fp16, 16 channels per row)
  unsigned char* XA = R2 + MF_R2_BYTES;                   // [2] im2col tiles
  unsigned char* XS = XA + 2 * MF_XA_BYTES;               // x staging buffer (bf16)

  const int tid = threadIdx.x, lane = tid & 31, warp = uniform_warp_id();
  const int S = p.S, Lp = p.Lp, tpb = p.tpb;
  const int nb_local = (int)((p.nblk - blockIdx.x + gridDim.x - 1) / gridDim.x);
  const int nt = nb_local * tpb;

  if (warp == 0) tmem_alloc(&tmem_slot, 128);
  if (tid == 0) {
    mbar_init(&bars[X_FULL], 1); mbar_init(&bars[X_EMPTY], MF_WORKERS);
    for (int a = 0; a < 2; ++a) {
      mbar_init(&bars[XA_FULL + a], MF_WORKERS); mbar_init(&bars[XA_EMPTY + a], 1);
      for (int s = 0; s < 3; ++s) { mbar_init(&bars[D_FULL + 2 * s + a], 1); mbar_init(&bars[D_EMPTY + 2 * s + a], MF_WORKERS); }
    }
    for (int r = 0; r < 2; ++r)
      for (int s = 0; s < MF_SLOTS; ++s) { mbar_init(&bars[R_FULL + 4 * r + s], MF_WORKERS); mbar_init(&bars[R_EMPTY + 4 * r + s], 1); }
    fence_mbar_init();
  }
  for (int i = tid; i < MF_W_BYTES / 16; i += MF_THREADS) reinterpret_cast<uint4*>(WS)[i] = __ldg(reinterpret_cast<const uint4*>(p.W) + i);
  for (int i = tid; i < (MF_R1_BYTES + MF_R2_BYTES + 2 * MF_XA_BYTES + p.xs_buf_bytes) / 16; i += MF_THREADS)
    reinterpret_cast<uint4*>(R1)[i] = make_uint4(0u, 0u, 0u, 0u);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  uint32_t bar_u = smem_u32(bars);
  asm volatile("" : "+r"(bar_u));                     // opaque: keeps the address in a register instead of re-deriving it
  auto BA = [&](int idx) { return bar_u + (uint32_t)idx * 8u; };
  const uint32_t xs_base = smem_u32(XS), r1_base = smem_u32(R1), r2_base = smem_u32(R2), xa_base = smem_u32(XA);

  if (warp == MF_TMA_WARP) {
    // ================= TMA producer: x of block i -> the staging buffer =================
    if (elect_one()) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap)) : "memory");
      for (int i = 0; i < nb_local; ++i) {
        if (i >= 1) mbar_wait_g<500>(BA(X_EMPTY), (i - 1) & 1);
        mbar_expect_tx(BA(X_FULL), (uint32_t)(MF_BLOCK_A * S * 2));
        const long long a0 = ((long long)blockIdx.x + (long long)i * gridDim.x) * MF_BLOCK_A;
#pragma unroll 1
        for (int j = 0; j < MF_BLOCK_A; ++j)
          tma_load_2d(xs_base + (uint32_t)((MF_XPAD + j * p.xs_stride) * 2), &tmap, 0, (int)((a0 + j) * p.rows_per_ascan), BA(X_FULL));
      }
    }
    __syncwarp();
  } else if (warp == MF_MMA_WARP) {
    // ================= MMA issuer: conv1 of tile g, conv2 of tile g - 2, stencil of tile g - 5 =================
    const bool leader = elect_one();
    const uint32_t id_bf16 = make_idesc_bf16(128, 16), id_f16 = make_idesc_f16(128, 16);
    const uint32_t hi = (uint32_t)(128 >> 4) | (1u << 14);               // SBO = 128 B, descriptor version 1
    const uint32_t w_u = smem_u32(WS) >> 4, r1_u = r1_base >> 4, r2_u = r2_base >> 4, xa_u = xa_base >> 4;   // 16-byte units
    auto lo = [](uint32_t start_u, uint32_t lbo_u) { return (start_u & 0x3FFFu) | (lbo_u << 16); };
    for (int g = 0; g < nt + 5; ++g) {
      if (g < nt) {
        const int b = g & 1;
        mbar_wait_g<0>(BA(XA_FULL + (b)), (g >> 1) & 1);
        if (g >= 2) mbar_wait_g<0>(BA(D_EMPTY + 2 * (0) + (b)), ((g >> 1) - 1) & 1);
        if (leader) {
          tc_fence_after();
          mma_bf16_ss2(tmem + (uint32_t)(TC_D1 + 16 * b), lo(xa_u + (uint32_t)(b * (MF_XA_BYTES / 16)), 128), hi, lo(w_u, 16), hi, id_bf16, 0u);
          commit_a(BA(XA_EMPTY + (b)));
          commit_a(BA(D_FULL + 2 * (0) + (b)));
        }
        __syncwarp();
      }
      const int u = g - 2;
      if (u >= 0 && u < nt) {
        const int b = u & 1, slot = u & (MF_SLOTS - 1);
        mbar_wait_g<0>(BA(R_FULL + 4 * (0) + ((u + 1) & (MF_SLOTS - 1))), ((u + 1) >> 2) & 1);
        if (u >= 2) mbar_wait_g<0>(BA(D_EMPTY + 2 * (1) + (b)), ((u >> 1) - 1) & 1);
        if (leader) {
          tc_fence_after();
          const uint32_t row0 = r1_u + (uint32_t)(MF_MIRROR + 128 * slot);
          // K chunk 1 of a row = the next row (LBO = 16 B): (tap 0 | tap 1) starts one row early, (tap 2 | 0) one row late
          mma_bf16_ss2(tmem + (uint32_t)(TC_D2 + 16 * b), lo(row0 - 1, 1), hi, lo(w_u + 32, 16), hi, id_f16, 0u);
          mma_bf16_ss2(tmem + (uint32_t)(TC_D2 + 16 * b), lo(row0 + 1, 1), hi, lo(w_u + 64, 16), hi, id_f16, 1u);
          commit_a(BA(D_FULL + 2 * (1) + (b)));
          commit_a(BA(R_EMPTY + 4 * (0) + ((u + MF_SLOTS - 1) & (MF_SLOTS - 1))));   // ring 1 tile u - 1 has no reader left
        }
        __syncwarp();
      }
      const int v = g - 5;
      if (v >= 0 && v < nt) {
        const int b = v & 1, slot = v & (MF_SLOTS - 1);
        mbar_wait_g<0>(BA(R_FULL + 4 * (1) + ((v + 1) & (MF_SLOTS - 1))), ((v + 1) >> 2) & 1);
        if (v >= 2) mbar_wait_g<0>(BA(D_EMPTY + 2 * (2) + (b)), ((v >> 1) - 1) & 1);
        if (leader) {
          tc_fence_after();
          const uint32_t row0 = r2_u + (uint32_t)(MF_MIRROR + 128 * slot);
#pragma unroll
          for (int t = 0; t < 11; ++t)
            mma_bf16_ss2(tmem + (uint32_t)(TC_D3 + 16 * b), lo(row0 + (uint32_t)(t - 5), MF_RROWS), hi, lo(w_u + 96 + 32 * t, 16), hi, id_f16,
                         t ? 1u : 0u);
          commit_a(BA(D_FULL + 2 * (2) + (b)));
          commit_a(BA(R_EMPTY + 4 * (1) + ((v + MF_SLOTS - 1) & (MF_SLOTS - 1))));
        }
        __syncwarp();
      }
    }
  } else if (warp < MF_WORKERS) {
    // ================= worker warps: thread = row i of the tile =================
    const int i = warp * 32 + lane;
    const uint32_t t_lane = (uint32_t)(warp * 32) << 16;
    // The pass loop is unrolled by four: with g = 4k + J every ring slot, double-buffer index and most barrier parities of
    // the four stages of a pass (im2col of tile g, epilogue 1 of g - 1, epilogue 2 of g - 3, epilogue 3 of g - 6) are
    // compile-time constants of J; only the lap parity (k & 1) stays in a register.  The probe of round 2 showed the
    // kernel bound by the worker warps' instruction stream (~9 cycles per instruction at one worker warp per scheduler and
    // CTA), so what is removed here -- index arithmetic on the uniform datapath, 64-bit A-scan indices, guarded wait loops
    // with a sleep and a spin counter, per-lane tests for the mirrored ring rows -- is time.
    // ring rule: the previous tile of a slot (tile - 4) was last read by the MMAs of tile - 3; slot 3 also owns the
    // mirrored rows in front of the ring (tile 0 reads them as zero rows): r_empty[.][3] carries an extra first phase
    auto spin = [](uint32_t bar, uint32_t parity) {                     // lean parity wait: probe + branch
      asm volatile(
          "{\n"
          ".reg .pred P1;\n"
          "W_%=:\n"
          "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
          "@!P1 bra W_%=;\n"
          "}\n" ::"r"(bar),
          "r"(parity)
          : "memory");
    };
    // row cursors: (A-scan, sample) of this thread's row in the tiles handled by im2col / epilogue 2 / epilogue 3
    struct Cur { int a, l, T, a_blk; };
    const int a_first = (int)blockIdx.x * MF_BLOCK_A, a_step = (int)gridDim.x * MF_BLOCK_A, nA = (int)p.A;
    auto cur_advance = [&](Cur& q) {
      if (++q.T == tpb) { q.T = 0; q.a_blk += a_step; q.a = q.a_blk; q.l = i; return; }
      q.l += 128;
      if (q.l >= Lp) { q.l -= Lp; ++q.a; }                            // Lp >= 136: at most one A-scan further
    };
    Cur c0{a_first, i, 0, a_first}, c2 = c0, c3 = c0;
    int x_blk = -1, blk0 = 0;                                            // block whose x is staged / block of the im2col cursor
    const bool first_rows = warp == 0, last_rows = warp == MF_WORKERS - 1;   // warps that own mirrored ring rows
    auto pass = [&](auto Jc, int k) {
      constexpr int J = decltype(Jc)::value;
      const int g = 4 * k + J;
      const uint32_t kp = (uint32_t)k & 1u;
      // ---- im2col row of tile g
      if (g < nt) {
        constexpr int b = J & 1;
        if (blk0 != x_blk) {
          x_blk = blk0;
          spin(BA(X_FULL), (uint32_t)x_blk & 1u);
        }
        uint32_t w0 = 0u, w1 = 0u, w2 = 0u;
        if (c0.l < S && c0.a < nA) {
          const uint32_t xa = xs_base + (uint32_t)(MF_XPAD + (c0.a - c0.a_blk) * p.xs_stride + c0.l - 1) * 2;
          uint16_t h0, h1, h2;
          asm volatile("ld.shared.u16 %0, [%1];" : "=h"(h0) : "r"(xa) : "memory");
          asm volatile("ld.shared.u16 %0, [%1];" : "=h"(h1) : "r"(xa + 2) : "memory");
          asm volatile("ld.shared.u16 %0, [%1];" : "=h"(h2) : "r"(xa + 4) : "memory");
          w0 = (uint32_t)h0 | ((uint32_t)h1 << 16);
          w1 = (uint32_t)h2 | 0x3F800000u;           // tap 3 = 1.0 (bias hi)
          w2 = 0x00003F80u;                          // tap 4 = 1.0 (bias lo)
        }
        if (g >= 2) spin(BA(XA_EMPTY + (b)), (uint32_t)(((J >> 1) + 1) & 1));   // ((g >> 1) - 1) & 1
        const uint32_t dst = xa_base + (uint32_t)(b * MF_XA_BYTES) + (uint32_t)i * 16;
        st_shared_v4(dst, w0, w1, w2, 0u);
        st_shared_v4(dst + 128 * 16, 0u, 0u, 0u, 0u);
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
          arrive_a(BA(XA_FULL + (b)));
          if (c0.T == tpb - 1) arrive_a(BA(X_EMPTY));                   // the block's x has been read
        }
        if (c0.T == tpb - 1) ++blk0;
        cur_advance(c0);
      }
      // ---- epilogue 1 of tile u = g - 1: relu(conv1) -> ring 1 (tile nt: zero rows, the stream ends)
      {
        constexpr int uJ = (J + 3) & 3, b = uJ & 1;
        const int u = g - 1;
        const uint32_t lap_p = J >= 1 ? kp : kp ^ 1u;                    // parity of the tile's lap (u >> 2)
        if (u >= 0 && u <= nt) {
          uint32_t w[4] = {0u, 0u, 0u, 0u};
          if (u < nt) {
            spin(BA(D_FULL + 2 * (0) + (b)), (uint32_t)((uJ >> 1) & 1));
            tc_fence_after();
            uint32_t r[8];
            tmem_ld8(tmem + t_lane + (uint32_t)(TC_D1 + 16 * b), r);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) arrive_a(BA(D_EMPTY + 2 * (0) + (b)));
#pragma unroll
            for (int e = 0; e < 4; ++e) w[e] = relu_pack_f16(__uint_as_float(r[2 * e]), __uint_as_float(r[2 * e + 1]));
          }
          if (uJ == MF_SLOTS - 1) spin(BA(R_EMPTY + 4 * (0) + (uJ)), lap_p);
          else if (u >= MF_SLOTS) spin(BA(R_EMPTY + 4 * (0) + (uJ)), lap_p ^ 1u);
          const uint32_t a = r1_base + (uint32_t)(MF_MIRROR + 128 * uJ) * 16 + (uint32_t)i * 16;
          st_shared_v4(a, w[0], w[1], w[2], w[3]);
          if (uJ == 0 && first_rows) { if (lane < MF_MIRROR) st_shared_v4(a + MF_SLOTS * 128 * 16, w[0], w[1], w[2], w[3]); }
          if (uJ == MF_SLOTS - 1 && last_rows) { if (lane >= 32 - MF_MIRROR) st_shared_v4(a - MF_SLOTS * 128 * 16, w[0], w[1], w[2], w[3]); }
          fence_async_smem();
          __syncwarp();
          if (lane == 0) arrive_a(BA(R_FULL + 4 * (0) + (uJ)));
        }
      }
      // ---- epilogue 2 of tile v = g - 3: relu(conv2 + bias), zero outside the A-scans -> ring 2
      {
        constexpr int vJ = (J + 1) & 3, b = vJ & 1;
        const int v = g - 3;
        const uint32_t lap_p = J >= 3 ? kp : kp ^ 1u;
        if (v >= 0 && v <= nt) {
          uint32_t w[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
          if (v < nt) {
            spin(BA(D_FULL + 2 * (1) + (b)), (uint32_t)((vJ >> 1) & 1));
            tc_fence_after();
            float y[16];
            tmem_ld16(tmem + t_lane + (uint32_t)(TC_D2 + 16 * b), y);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) arrive_a(BA(D_EMPTY + 2 * (1) + (b)));
            if (c2.l < S && c2.a < nA) {
#pragma unroll
              for (int e = 0; e < 8; ++e) w[e] = relu_pack_f16(y[2 * e] + p.b2[2 * e], y[2 * e + 1] + p.b2[2 * e + 1]);
            }
            cur_advance(c2);
          }
          if (vJ == MF_SLOTS - 1) spin(BA(R_EMPTY + 4 * (1) + (vJ)), lap_p);
          else if (v >= MF_SLOTS) spin(BA(R_EMPTY + 4 * (1) + (vJ)), lap_p ^ 1u);
          const uint32_t a = r2_base + (uint32_t)(MF_MIRROR + 128 * vJ) * 16 + (uint32_t)i * 16;
#pragma unroll
          for (int ch = 0; ch < 2; ++ch) {
            const uint32_t ac = a + (uint32_t)(ch * MF_RROWS * 16);
            st_shared_v4(ac, w[4 * ch], w[4 * ch + 1], w[4 * ch + 2], w[4 * ch + 3]);
            if (vJ == 0 && first_rows) {
              if (lane < MF_MIRROR) st_shared_v4(ac + MF_SLOTS * 128 * 16, w[4 * ch], w[4 * ch + 1], w[4 * ch + 2], w[4 * ch + 3]);
            }
            if (vJ == MF_SLOTS - 1 && last_rows) {
              if (lane >= 32 - MF_MIRROR) st_shared_v4(ac - MF_SLOTS * 128 * 16, w[4 * ch], w[4 * ch + 1], w[4 * ch + 2], w[4 * ch + 3]);
            }
          }
          fence_async_smem();
          __syncwarp();
          if (lane == 0) arrive_a(BA(R_FULL + 4 * (1) + (vJ)));
        }
      }
      // ---- epilogue 3 of tile w = g - 6: f = column 0 of the stencil accumulator - mean(b_bg)
      {
        constexpr int wJ = (J + 2) & 3, b = wJ & 1;
        const int wv = g - 6;
        if (wv >= 0 && wv < nt) {
          spin(BA(D_FULL + 2 * (2) + (b)), (uint32_t)((wJ >> 1) & 1));
          tc_fence_after();
          const float fv = __uint_as_float(tmem_ld1(tmem + t_lane + (uint32_t)(TC_D3 + 16 * b))) - p.f_const;
          tc_fence_before();
          __syncwarp();
          if (lane == 0) arrive_a(BA(D_EMPTY + 2 * (2) + (b)));
          if (c3.l < S && c3.a < nA) p.f[(size_t)c3.a * S + c3.l] = fv;
          cur_advance(c3);
        }
      }
    };
    for (int k = 0; 4 * k < nt + 7; ++k) {
      pass(std::integral_constant<int, 0>{}, k);
      pass(std::integral_constant<int, 1>{}, k);
      pass(std::integral_constant<int, 2>{}, k);
      pass(std::integral_constant<int, 3>{}, k);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 128);
}

uint16_t mf_f2h(float f) {
  const __half h = __float2half_rn(f);
  uint16_t u;
  memcpy(&u, &h, 2);
  return u;
}
uint16_t mf_f2bf(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  return (uint16_t)((u + 0x7fffu + ((u >> 16) & 1u)) >> 16);
}
float mf_bf2f(uint16_t h) {
  uint32_t u = (uint32_t)h << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

typedef CUresult (*EncodeTiledFn2)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn2 encode_tiled2() {
  static EncodeTiledFn2 fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess)
      f = nullptr;
    return reinterpret_cast<EncodeTiledFn2>(f);
  }();
  return fn;
}

}  // namespace

bool mscn_front_supported(int S) { return S % 16 == 0 && S >= 128 && S <= 512; }

// B operands [14][2 chunks][16 rows][8]: 0 = conv1 (bf16: taps 0..2, bias hi / lo at K = 3, 4; rows 8..15 zero),
// 1 = conv2 (tap 0 | tap 1), 2 = conv2 (tap 2 | 0), 3..13 = stencil taps (row 0 = M[c][t], rows 1..15 zero)   (fp16)
void mscn_front_pack(const float* w1 /*[8][3]*/, const float* b1 /*[8]*/, const float* w2 /*[16][8][3]*/,
                     const float* wbg /*[16][11]*/, std::vector<uint16_t>& W) {
  W.assign(MF_W_BYTES / 2, 0);
  auto at = [&](int op, int k, int row) -> uint16_t& { return W[(size_t)op * 256 + ((size_t)(k >> 3) * 16 + row) * 8 + (k & 7)]; };
  for (int co = 0; co < 8; ++co) {
    for (int t = 0; t < 3; ++t) at(0, t, co) = mf_f2bf(w1[co * 3 + t]);
    const uint16_t hi = mf_f2bf(b1[co]);
    at(0, 3, co) = hi;
    at(0, 4, co) = mf_f2bf(b1[co] - mf_bf2f(hi));
  }
  for (int co = 0; co < 16; ++co)
    for (int ci = 0; ci < 8; ++ci) {
      at(1, ci, co) = mf_f2h(w2[(co * 8 + ci) * 3 + 0]);
      at(1, 8 + ci, co) = mf_f2h(w2[(co * 8 + ci) * 3 + 1]);
      at(2, ci, co) = mf_f2h(w2[(co * 8 + ci) * 3 + 2]);
    }
  for (int t = 0; t < 11; ++t)
    for (int c = 0; c < 16; ++c) at(3 + t, c, 0) = mf_f2h(((t == 5 ? 1.f : 0.f) - wbg[c * 11 + t]) / 16.f);
}

void op_mscn_front(Ctx& c, const void* x_bf16, int64_t A, int S, const void* W, const float* b2_host, float f_const, float* f) {
  if (c.dry) return;
  PAUT_CHECK(mscn_front_supported(S), PAUT_ERR_UNSUPPORTED, "msc_n front end: unsupported signal length");
  PAUT_CHECK((reinterpret_cast<uintptr_t>(x_bf16) & 15) == 0, PAUT_ERR_INVALID, "msc_n front end: x must be 16-byte aligned");
  MscnArgs p;
  p.W = static_cast<const uint16_t*>(W);
  memcpy(p.b2, b2_host, sizeof(p.b2));
  p.f_const = f_const; p.f = f;
  p.A = A; p.nblk = (A + MF_BLOCK_A - 1) / MF_BLOCK_A;
  p.S = S; p.Lp = S + MF_HALO; p.tpb = MF_BLOCK_A * p.Lp / 128;
  p.xs_stride = (S + 16 + 63) / 64 * 64;
  p.xs_buf_bytes = (2 * MF_XPAD + MF_BLOCK_A * p.xs_stride) * 2;
  const int Wd = S <= 256 ? S : S / 2;
  p.rows_per_ascan = S / Wd;
  EncodeTiledFn2 enc = encode_tiled2();
  PAUT_CHECK(enc != nullptr, PAUT_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  CUtensorMap tmap;
  const cuuint64_t gdim[2] = {(cuuint64_t)Wd, (cuuint64_t)(A * p.rows_per_ascan)};
  const cuuint64_t gstride[1] = {(cuuint64_t)Wd * 2};
  const cuuint32_t box[2] = {(cuuint32_t)Wd, (cuuint32_t)p.rows_per_ascan};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(x_bf16), gdim, gstride, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  PAUT_CHECK(r == CUDA_SUCCESS, PAUT_ERR_CUDA, "cuTensorMapEncodeTiled failed (msc_n front end input)");
  const size_t smem = (size_t)MF_W_BYTES + MF_R1_BYTES + MF_R2_BYTES + 2 * MF_XA_BYTES + (size_t)p.xs_buf_bytes;
  smem_optin(c, k_mscn_front);
  static const int ctas_per_sm = [] { const char* e = std::getenv("PAUT_MSCN_CTAS"); const int v = e ? atoi(e) : 4; return v >= 1 && v <= 4 ? v : 4; }();
  long long grid = (long long)c.num_sms * ctas_per_sm;  // four CTAs per SM (shared memory ~50 KB, 128 TMEM columns each)
  if (grid > p.nblk) grid = p.nblk;
  k_mscn_front<<<(unsigned)grid, MF_THREADS, smem, c.stream>>>(tmap, p);
  c.launched("mscn_front");
}

}  // namespace paut
