// tcgen05 implicit-GEMM Conv1d (+ folded BatchNorm shift, residual, ReLU, pooled mean) for the bf16 mode.
//
// Activation layout ("flat rows"): every A-scan owns L rows of C channels (bf16, channels-last) followed by
// HALO zero rows; HALO more zero rows precede the first A-scan:  row(a, l) = HALO + a*(L+HALO) + l.
// Because the halo rows are zero and at least pad*dilation wide, a convolution is a plain 1-D convolution
// over the whole flat row axis: no per-A-scan edge handling, and an M tile is simply 128 consecutive rows.
//
// K loop: for each block of 64 input channels, the window rows [r0 - pad*dil, r0 + 128 + pad*dil) are
// staged once in shared memory in the canonical K-major layout ([8 chunks][window rows][16 B]); tap t of
// the kernel is the SAME window addressed through a descriptor whose start address is advanced by
// t*dil rows (rows are uniformly 16 B apart inside a chunk), so no im2col copy exists anywhere.
// The matching weight blocks ([tap][8 chunks][NT rows][16 B]) ride in the same pipeline stage.
//
// Warp roles (416 threads): warps 9-12 stream the windows with cp.async into a 5-deep ring (weights of the CTA's
// N tile are resident in shared memory); one elected thread of warp 8 issues the tcgen05.mma of each stage
// (commit -> stage-free barrier; last stage of a tile also commits -> accumulator-full barrier); warps 0-7 are
// the epilogue: TMEM -> registers -> per-warp swizzled shared-memory transposition -> shift, residual, ReLU ->
// bf16 rows in HBM (halo rows are written as zeros to keep the layout invariant) and/or per-tile column sums
// for the pooled mean, with 8 lanes per row so that loads and stores are 64 B per row.  Two TMEM accumulators
// let tile i+1's MMAs overlap tile i's epilogue.  CTAs are persistent over (row tile, N tile) pairs.
//
// Stride-2 convolutions (the enhanced encoder's pyramid) use the same kernel through a space-to-depth view:
// flat rows [R, C] with an even period are read as [R/2, 2C] (row pair (2m, 2m+1) = one row of 2C channels), and
//   y[l] = W0 x[2l-1] + W1 x[2l] + W2 x[2l+1]  =  [0 | W0] x2[l-1] + [W1 | W2] x2[l]
// is a stride-1 convolution with two taps (offsets -1, 0) over 2C channels.  The all-zero half of tap 0 is
// never multiplied (`skip_lo`), so the MMA work equals the real 3-tap convolution.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "tc_common.cuh"

namespace paut {

using namespace tc;

namespace {

constexpr int CT_STAGES = 5;
constexpr int CT_EPI_WARPS = 8;          // warps 0-7 epilogue (2 per TMEM lane quarter), 8 MMA issuer, 9-12 cp.async loaders
constexpr int CT_EPI = CT_EPI_WARPS * 32;
constexpr int CT_THREADS = CT_EPI + 32 + 128;
constexpr int CT_LOADERS = 128;

struct ConvTcArgs {
  const __nv_bfloat16* in;      // flat rows [R, Cin]
  int64_t R;                    // total flat rows
  int Cin, Cout;
  const __nv_bfloat16* Wp;      // [Cout/NT][Cin/CB blocks][taps][CB/8 chunks][NT rows][8]
  const float* shift;           // [Cout]
  int taps, dil, pad;           // stride 1
  int NT, CB;                   // N tile (<= 128), input-channel block (<= 64, multiple of 16)
  int relu;
  const __nv_bfloat16* res;     // flat rows [R, ldr] (nullable), same row geometry
  int ldr;
  __nv_bfloat16* out;           // flat rows [R, ldc] at column offset coff (nullable)
  int ldc, coff;
  float* pool;                  // per-tile column sums [tiles][nseg segments][Cout] (nullable)
  int nseg;                     // A-scans a 128-row tile can touch: 2 (period >= 127 rows) or 3
  int skip_lo;                  // space-to-depth stride-2 view: tap 0 multiplies only the upper half of Cin
  int grouped;                  // 1: channel block cb is its own conv (gtaps[cb] taps, N = NT / ncb columns);
                                // 2: ngroups convs share the ONE channel block and differ in dilation gdil[g]
  int ngroups;
  int gtaps[4], goff[4], gdil[4];   // taps, resident-weight offset (16-byte units), dilation of every group
  int L, Lp, H0;                // geometry: valid rows per A-scan, period, leading halo
  int64_t A;
  int wrows;                    // window rows = 128 + (taps-1)*dil
  int a_stage_bytes, w_bytes;   // one A stage; all weights of one N tile
  int64_t num_tiles;            // row tiles
  unsigned long long* dbg;      // optional cycle probe (PAUT_CONV_DEBUG=1): CTA 0 role timings
};

__device__ __forceinline__ void named_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Persistent CTA = one N tile (weights resident in shared memory) x a strided set of 128-row tiles.
template <bool OUT, bool POOL, bool RES>
__global__ void __launch_bounds__(CT_THREADS, 1) k_conv_tc(ConvTcArgs p) {
  extern __shared__ __align__(128) unsigned char smem[];
  // all mbarriers in ONE array behind a pinned base address (as separate variables every use re-derived its shared-window
  // address with S2R + LEA): full[CT_STAGES] | empty[CT_STAGES] | acc_full[2] | acc_empty[2]
  __shared__ __align__(8) uint64_t bars[2 * CT_STAGES + 4];
  constexpr int FULL = 0, EMPTY = CT_STAGES, ACC_FULL = 2 * CT_STAGES, ACC_EMPTY = 2 * CT_STAGES + 2;
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float pool_s[4 * 2 * 128];                    // [TMEM quarter][segment][NT]; 4*nseg*NT <= 1024
  __shared__ __align__(128) float tsm[CT_EPI_WARPS][32 * 32];   // per-warp 32x32 fp32 transposition tile (swizzled)
  __shared__ __align__(16) float shift_s[128];

  const int tid = threadIdx.x, lane = tid & 31, warp = uniform_warp_id();
  const int NT = p.NT;
  const int ntn = p.Cout / NT;
  const int ncb = p.Cin / p.CB;
  const int chunks = p.CB / 8;
  const int nt = blockIdx.x % ntn;                         // this CTA's N tile
  const int64_t tile0 = blockIdx.x / ntn, tstep = gridDim.x / ntn;
  unsigned char* Wres = smem;                              // [ncb][taps][chunks][NT][16 B]
  unsigned char* Aring = smem + p.w_bytes;                 // [CT_STAGES][chunks][wrows][16 B]

  if (warp == 0) tmem_alloc(&tmem_slot, 256);
  if (tid == 0) {
    for (int s = 0; s < CT_STAGES; ++s) { mbar_init(&bars[FULL + s], CT_LOADERS); mbar_init(&bars[EMPTY + s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&bars[ACC_FULL + a], 1); mbar_init(&bars[ACC_EMPTY + a], CT_EPI); }
    fence_mbar_init();
  }
  {  // resident weights of this N tile: one contiguous block in the packed layout
    const uint4* src = reinterpret_cast<const uint4*>(p.Wp) + (size_t)nt * (p.w_bytes / 16);
    uint4* dst = reinterpret_cast<uint4*>(Wres);
    for (int i = tid; i < p.w_bytes / 16; i += CT_THREADS) dst[i] = __ldg(src + i);
  }
  for (int i = tid; i < NT; i += CT_THREADS) shift_s[i] = __ldg(p.shift + nt * NT + i);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  uint32_t bar_u = smem_u32(bars);
  asm volatile("" : "+r"(bar_u));
  auto BA = [&](int idx) { return bar_u + (uint32_t)idx * 8u; };
  const bool probe = p.dbg != nullptr && blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == CT_EPI_WARPS || warp == CT_EPI_WARPS + 1);
  unsigned long long pt[4] = {0, 0, 0, 0};

  if (warp > CT_EPI_WARPS) {
    // ================= loaders: cp.async of the A windows, up to CT_STAGES stages of look-ahead =================
    const int l = tid - (CT_EPI_WARPS + 1) * 32;            // 0..127
    const int ch = l & 7;                                   // fixed chunk (8 lanes cover one 128-byte row)
    const int rsub = l >> 3;                                // rows rsub, rsub+16, ...
    const uint32_t a_ring = smem_u32(Aring);
    int stage = 0;
    uint32_t use = 0;
    for (int64_t tile = tile0; tile < p.num_tiles; tile += tstep) {
      const int64_t wr0 = tile * 128 - (int64_t)p.pad * p.dil;
      for (int cb = 0; cb < ncb; ++cb) {
        const long long q0 = probe ? clock64() : 0;
        if (use > 0) wait_a(BA(EMPTY + (stage)), (use - 1) & 1);        // MMAs that read this slot are done
        const long long q1 = probe ? clock64() : 0;
        if (ch < chunks) {
          uint32_t dst = a_ring + (uint32_t)stage * p.a_stage_bytes + (uint32_t)(ch * p.wrows + rsub) * 16;
          if (wr0 >= 0 && wr0 + p.wrows <= p.R) {
            // interior tile (all but the first/last of the volume): no per-row bounds checks, incremental addressing
            const __nv_bfloat16* src = p.in + (size_t)(wr0 + rsub) * p.Cin + (size_t)cb * p.CB + ch * 8;
            const size_t sstep = (size_t)16 * p.Cin;
            for (int r = rsub; r < p.wrows; r += 16) {
              asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
              dst += 256;
              src += sstep;
            }
          } else {
            const __nv_bfloat16* src0 = p.in + (size_t)cb * p.CB + ch * 8;
            for (int r = rsub; r < p.wrows; r += 16) {
              const int64_t row = wr0 + r;
              const bool ok = row >= 0 && row < p.R;
              // src-size 0 zero-fills the 16 bytes (rows outside the volume)
              asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src0 + (ok ? row : 0) * p.Cin),
                           "r"(ok ? 16 : 0)
                           : "memory");
              dst += 256;
            }
          }
        }
        // the slot's "full" barrier counts this thread in as soon as ITS copies have landed -- no wait here, and no
        // coupling of the hand-over to the request of a later stage (a wait_group-based hand-over of stage j only
        // happens once stage j + LAG can be requested, i.e. after the MMAs of stage j - 1 have completed: that
        // serialised consecutive stages and left the tensor pipe idle between them)
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(BA(FULL + (stage))) : "memory");
        if (probe) { const long long q2 = clock64(); pt[0] += q1 - q0; pt[1] += q2 - q1; pt[2] += 1; }
        if (++stage == CT_STAGES) { stage = 0; ++use; }
      }
    }
    if (probe) { p.dbg[0] = pt[0]; p.dbg[1] = pt[1]; p.dbg[2] = pt[2]; }
    asm volatile("cp.async.wait_group 0;" ::: "memory");   // nothing may be in flight when the CTA retires
  } else if (warp == CT_EPI_WARPS) {
    // ================= MMA issuer (one elected lane) =================
    const bool leader = elect_one();
    const uint32_t idesc = make_idesc_bf16(128, NT);
    const uint32_t a_ring = smem_u32(Aring), w_addr = smem_u32(Wres);
    int stage = 0, it = 0;
    uint32_t use = 0;
    for (int64_t tile = tile0; tile < p.num_tiles; tile += tstep, ++it) {
      const int acc = it & 1;
      const long long m0 = probe ? clock64() : 0;
      if (it >= 2) wait_a(BA(ACC_EMPTY + (acc)), ((it >> 1) - 1) & 1);         // epilogue drained this accumulator
      if (probe) pt[0] += clock64() - m0;
      for (int cb = 0; cb < ncb; ++cb) {
        const long long m1 = probe ? clock64() : 0;
        wait_a(BA(FULL + (stage)), use & 1);
        const long long m2 = probe ? clock64() : 0;
        if (leader) {
          tc_fence_after();
          fence_async_smem();                                // cp.async (generic proxy) writes -> tensor-core operand reads
          const uint32_t a_addr = a_ring + (uint32_t)stage * p.a_stage_bytes;
          // descriptors as (lo, hi) words: hi (SBO = 128 B, version 1) is constant; lo = start address | LBO << 16
          // advances by constant 16-byte-unit amounts with one 32-bit add per MMA
          const uint32_t desc_hi = (uint32_t)(128 >> 4) | (1u << 14);
          const uint32_t a_ks = (uint32_t)(2 * p.wrows), a_t = (uint32_t)p.dil;
          if (p.grouped == 2) {
            // multi-dilation branches over the same input block: branch g = taps at (t - pad) * gdil[g] around the
            // centre of the common window, output columns [g * GN, (g + 1) * GN)
            const int GN = NT / p.ngroups;
            const uint32_t idesc_g = make_idesc_bf16(128, GN);
            const uint32_t b_step = (uint32_t)(2 * GN);
            for (int g = 0; g < p.ngroups; ++g) {
              const uint32_t d = tmem + acc * 128 + g * GN;
              const uint32_t gd = (uint32_t)p.gdil[g];
              uint32_t ad_t = (((a_addr >> 4) + (uint32_t)(p.pad * (p.dil - p.gdil[g]))) & 0x3FFFu) | ((uint32_t)p.wrows << 16);
              uint32_t bd = (((w_addr >> 4) + (uint32_t)p.goff[g]) & 0x3FFFu) | ((uint32_t)GN << 16);
              uint32_t accum = 0u;
              for (int t = 0; t < p.gtaps[g]; ++t) {
                uint32_t ad = ad_t;
                for (int ks = 0; ks < chunks / 2; ++ks) {
                  mma_bf16_ss2(d, ad, desc_hi, bd, desc_hi, idesc_g, accum);
                  accum = 1u;
                  ad += a_ks;
                  bd += b_step;
                }
                ad_t += gd;
              }
            }
          } else if (p.grouped) {
            // branch cb: its own tap count (centred in the common window), its own 32-column block of the accumulator
            const int GN = NT / ncb, tg = p.gtaps[cb];
            const uint32_t idesc_g = make_idesc_bf16(128, GN);
            const uint32_t d = tmem + acc * 128 + cb * GN;
            uint32_t ad_t = (((a_addr >> 4) + (uint32_t)((p.pad - tg / 2) * p.dil)) & 0x3FFFu) | ((uint32_t)p.wrows << 16);
            uint32_t bd = (((w_addr >> 4) + (uint32_t)p.goff[cb]) & 0x3FFFu) | ((uint32_t)GN << 16);
            const uint32_t b_step = (uint32_t)(2 * GN);
            uint32_t accum = 0u;
            for (int t = 0; t < tg; ++t) {
              uint32_t ad = ad_t;
              for (int ks = 0; ks < chunks / 2; ++ks) {
                mma_bf16_ss2(d, ad, desc_hi, bd, desc_hi, idesc_g, accum);
                accum = 1u;
                ad += a_ks;
                bd += b_step;
              }
              ad_t += a_t;
            }
          } else {
            const uint32_t d = tmem + acc * 128;
            uint32_t ad_t = ((a_addr >> 4) & 0x3FFFu) | ((uint32_t)p.wrows << 16);
            uint32_t bd = (((w_addr >> 4) + (uint32_t)(cb * p.taps * chunks * NT)) & 0x3FFFu) | ((uint32_t)NT << 16);
            const uint32_t b_step = (uint32_t)(2 * NT);
            uint32_t accum = cb ? 1u : 0u;
            for (int t = 0; t < p.taps; ++t) {
              if (p.skip_lo && t == 0 && 2 * cb < ncb) {     // [0 | W0]: the lower channel half of tap 0 is zero
                ad_t += a_t;
                bd += b_step * (uint32_t)(chunks / 2);
                continue;
              }
              uint32_t ad = ad_t;
              for (int ks = 0; ks < chunks / 2; ++ks) {
                mma_bf16_ss2(d, ad, desc_hi, bd, desc_hi, idesc, accum);
                accum = 1u;
                ad += a_ks;
                bd += b_step;
              }
              ad_t += a_t;
            }
          }
          commit_a(BA(EMPTY + (stage)));
          if (cb == ncb - 1) commit_a(BA(ACC_FULL + (acc)));
        }
        __syncwarp();
        if (probe) { pt[1] += m2 - m1; pt[2] += clock64() - m2; pt[3] += 1; }
        if (++stage == CT_STAGES) { stage = 0; ++use; }
      }
    }
    if (probe) { p.dbg[8] = pt[0]; p.dbg[9] = pt[1]; p.dbg[10] = pt[2]; p.dbg[11] = pt[3]; }
  } else {
    // ================= epilogue: warp w owns TMEM lane quarter q = w%4 and the 32-column passes h, h+2, ... (h = w/4)
    // Row-per-lane accumulators go through a per-warp swizzled shared-memory tile once; everything else
    // (shift, residual, ReLU, halo masking, bf16 stores, pooled sums) happens in the transposed domain where
    // 8 lanes cover 32 consecutive columns of one row: residual loads and output stores touch 4 rows x 64 B
    // per instruction.  OUT / POOL / RES are compile-time, tile-uniform conditions (tile fully inside the
    // volume, warp rows free of halo rows, tile inside one A-scan) select straight-line fast paths, and all
    // addresses advance incrementally.
    const int q = warp & 3, h = warp >> 2;
    const int rr = lane >> 3, cg = lane & 7;               // transposed domain: rows rr + 4i, columns 4cg..4cg+3
    const uint32_t tw = smem_u32(&tsm[warp][0]);
    const int npass = (NT - 32 * h + 63) / 64;              // passes at c0 = 32h + 64k < NT
    const float lo = p.relu ? 0.f : -INFINITY;              // ReLU as a branch-free max
    float4 sh4[2];
    bool col_ok[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int c = 32 * h + 64 * k + 4 * cg;
      col_ok[k] = k < npass && c < NT;
      sh4[k] = col_ok[k] ? *reinterpret_cast<const float4*>(shift_s + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    uint32_t st_addr[8], ld_addr[8];                        // swizzled tile addresses (loop invariant)
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      st_addr[c] = tw + (uint32_t)lane * 128 + (uint32_t)((c ^ (lane & 7)) * 16);
      const int rl = rr + 4 * c;
      ld_addr[c] = tw + (uint32_t)rl * 128 + (uint32_t)((cg ^ (rl & 7)) * 16);
    }
    // element offsets of this lane's first row (tile 0 of this CTA) in the output / residual; advanced per tile
    const int64_t row0 = tile0 * 128 + q * 32 + rr;
    int64_t o_off = row0 * p.ldc + p.coff + nt * NT + 32 * h + 4 * cg;
    int64_t r_off = row0 * p.ldr + nt * NT + 32 * h + 4 * cg;
    const int64_t o_tile = tstep * 128 * (int64_t)p.ldc, r_tile = tstep * 128 * (int64_t)p.ldr;
    const int o_row4 = 4 * p.ldc, r_row4 = 4 * p.ldr;       // 4 rows further (elements)
    int it = 0;
    for (int64_t tile = tile0; tile < p.num_tiles; tile += tstep, ++it, o_off += o_tile, r_off += r_tile) {
      const int acc = it & 1;
      // ---- geometry of the tile (one division, tile-uniform): first A-scan, first rows of the next two
      const int t0 = (int)(tile * 128);
      const int rel0 = t0 - p.H0;
      const int a_first = rel0 >= 0 ? (int)((unsigned)rel0 / (unsigned)p.Lp) : 0;
      const int seg1_row = p.H0 + (a_first + 1) * p.Lp - t0;
      const int split1 = seg1_row < 128 ? seg1_row : 128;
      const int split2 = (p.nseg > 2 && seg1_row + p.Lp < 128) ? seg1_row + p.Lp : 128;
      const bool has_seg1 = split1 < 128, has_seg2 = split2 < 128;
      // validity of this lane's own row -> bit mask of the warp's 32 rows
      const int trow = q * 32 + lane;
      const int a_of_row = a_first + (trow >= split1 ? 1 : 0) + (trow >= split2 ? 1 : 0);
      const int l_of_row = rel0 + trow - a_of_row * p.Lp;
      const bool valid = l_of_row >= 0 && l_of_row < p.L && a_of_row < p.A && (int64_t)t0 + trow < p.R;
      const uint32_t vmask = __ballot_sync(0xffffffffu, valid);
      const bool all_valid = vmask == 0xffffffffu;           // warp-uniform
      const bool in_volume = (int64_t)t0 + 128 <= p.R;       // tile-uniform: no row bound checks needed
      // (A direct epilogue -- the lane keeps its row and stores its 64 bytes with four 16-byte accesses, no transposition
      // through shared memory -- was measured in round 2: the conv of MSC-Conv1D went from 72.6 to 79.5 ms per 1 M
      // A-scans.  A store instruction whose 32 lanes touch 32 different lines costs the LSU more than the transposition
      // costs the shared-memory port.)
      // ---- residual rows of both passes, requested before the accumulator is waited for
      uint2 rres[2][8];
      if (RES) {
#pragma unroll
        for (int k = 0; k < 2; ++k) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            rres[k][i] = make_uint2(0u, 0u);
            if (p.res != nullptr && col_ok[k] && (all_valid || ((vmask >> (rr + 4 * i)) & 1u)))
              rres[k][i] = __ldg(reinterpret_cast<const uint2*>(p.res + r_off + 64 * k + (int64_t)i * r_row4));
          }
        }
      }
      const long long e0 = probe ? clock64() : 0;
      wait_a(BA(ACC_FULL + (acc)), (it >> 1) & 1);
      const long long e1 = probe ? clock64() : 0;
      tc_fence_after();
      if (npass == 0) {                                      // narrow N tile: this warp only keeps the barrier phases in step
        tc_fence_before();
        arrive_a(BA(ACC_EMPTY + (acc)));
      }
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        if (k < npass) {
          uint32_t t32[32];
          tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + acc * 128 + 32 * h + 64 * k, t32);
          if (k == npass - 1) {                              // accumulator drained: the MMAs of tile it+2 may start
            tc_fence_before();
            arrive_a(BA(ACC_EMPTY + (acc)));
          }
#pragma unroll
          for (int c = 0; c < 8; ++c)
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(st_addr[c]), "r"(t32[4 * c]), "r"(t32[4 * c + 1]),
                         "r"(t32[4 * c + 2]), "r"(t32[4 * c + 3])
                         : "memory");
          __syncwarp();
          float4 t[8];
#pragma unroll
          for (int i = 0; i < 8; ++i)
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(t[i].x), "=f"(t[i].y), "=f"(t[i].z), "=f"(t[i].w)
                         : "r"(ld_addr[i]));
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            t[i].x += sh4[k].x; t[i].y += sh4[k].y; t[i].z += sh4[k].z; t[i].w += sh4[k].w;
            if (RES) {
              const __nv_bfloat162 r0 = *reinterpret_cast<const __nv_bfloat162*>(&rres[k][i].x);
              const __nv_bfloat162 r1 = *reinterpret_cast<const __nv_bfloat162*>(&rres[k][i].y);
              t[i].x += __low2float(r0); t[i].y += __high2float(r0); t[i].z += __low2float(r1); t[i].w += __high2float(r1);
            }
            t[i].x = fmaxf(t[i].x, lo); t[i].y = fmaxf(t[i].y, lo); t[i].z = fmaxf(t[i].z, lo); t[i].w = fmaxf(t[i].w, lo);
          }
          if (!all_valid) {                                  // halo rows stay zero
#pragma unroll
            for (int i = 0; i < 8; ++i)
              if (!((vmask >> (rr + 4 * i)) & 1u)) t[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
          if (OUT && col_ok[k]) {
            __nv_bfloat16* op = p.out + o_off + 64 * k;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              __nv_bfloat162 o0 = __floats2bfloat162_rn(t[i].x, t[i].y), o1 = __floats2bfloat162_rn(t[i].z, t[i].w);
              if (in_volume || (int64_t)t0 + q * 32 + rr + 4 * i < p.R)
                *reinterpret_cast<uint2*>(op + (int64_t)i * o_row4) =
                    make_uint2(*reinterpret_cast<uint32_t*>(&o0), *reinterpret_cast<uint32_t*>(&o1));
            }
          }
          if (POOL) {
            float4 ps0 = make_float4(0.f, 0.f, 0.f, 0.f), ps1 = ps0, ps2 = ps0;
#pragma unroll
            for (int i = 0; i < 8; ++i) { ps0.x += t[i].x; ps0.y += t[i].y; ps0.z += t[i].z; ps0.w += t[i].w; }
            if (has_seg1) {
#pragma unroll
              for (int i = 0; i < 8; ++i)
                if (q * 32 + rr + 4 * i >= split1) { ps1.x += t[i].x; ps1.y += t[i].y; ps1.z += t[i].z; ps1.w += t[i].w; }
            }
            if (has_seg2) {
#pragma unroll
              for (int i = 0; i < 8; ++i)
                if (q * 32 + rr + 4 * i >= split2) { ps2.x += t[i].x; ps2.y += t[i].y; ps2.z += t[i].z; ps2.w += t[i].w; }
            }
            // sums over the warp's 32 rows: lanes with equal cg (xor 8, 16), fixed order
            auto red = [&](float4& v) {
#pragma unroll
              for (int m = 8; m <= 16; m <<= 1) {
                v.x += __shfl_xor_sync(0xffffffffu, v.x, m); v.y += __shfl_xor_sync(0xffffffffu, v.y, m);
                v.z += __shfl_xor_sync(0xffffffffu, v.z, m); v.w += __shfl_xor_sync(0xffffffffu, v.w, m);
              }
            };
            red(ps0);
            if (has_seg1) red(ps1);
            if (has_seg2) red(ps2);
            if (lane < 8 && col_ok[k]) {
              float* ps = pool_s + (q * p.nseg) * NT + 32 * h + 64 * k + 4 * cg;
              *reinterpret_cast<float4*>(ps) = make_float4(ps0.x - ps1.x, ps0.y - ps1.y, ps0.z - ps1.z, ps0.w - ps1.w);
              *reinterpret_cast<float4*>(ps + NT) = make_float4(ps1.x - ps2.x, ps1.y - ps2.y, ps1.z - ps2.z, ps1.w - ps2.w);
              if (p.nseg > 2) *reinterpret_cast<float4*>(ps + 2 * NT) = ps2;
            }
          }
          __syncwarp();                                      // the tile is rewritten by the next pass
        }
      }
      if (POOL) {
        // combine the four lane quarters of this column half (128 threads, named barrier 2 + h), fixed order
        named_sync(2 + h, 128);
        const int tl = q * 32 + lane;                        // 0..127 within the half
        const int qs = p.nseg * NT;
        for (int i = tl; i < p.nseg * 64; i += 128) {
          const int sgm = i >> 6, cc = i & 63;               // cc: 0..31 -> pass 0, 32..63 -> pass 1
          const int col = 32 * h + 64 * (cc >> 5) + (cc & 31);
          if ((cc >> 5) < npass && col < NT) {
            const float* ps = pool_s + sgm * NT + col;
            p.pool[((size_t)tile * p.nseg + sgm) * p.Cout + nt * NT + col] = ps[0] + ps[qs] + ps[2 * qs] + ps[3 * qs];
          }
        }
        named_sync(2 + h, 128);
      }
      if (probe) { pt[0] += e1 - e0; pt[1] += clock64() - e1; pt[2] += 1; }
    }
    if (probe && warp == 0) { p.dbg[16] = pt[0]; p.dbg[17] = pt[1]; p.dbg[18] = pt[2]; p.dbg[19] = pt[3]; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

// pooled mean from the per-tile partial sums, in a fixed order: out[a, poff + c] = sum / L
__global__ void k_pool_finish(const float* __restrict__ partial, float* __restrict__ out, int ldp, int poff,
                              int64_t A, int C, int L, int Lp, int H0, int nseg) {
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= A * C) return;
  const int64_t a = idx / C;
  const int c = (int)(idx - a * C);
  const int64_t first = H0 + a * Lp, last = first + L - 1;
  float s = 0.f;
  for (int64_t t = first / 128; t <= last / 128; ++t) {
    const int64_t rel0 = t * 128 - H0;
    const int64_t a_first = rel0 >= 0 ? rel0 / Lp : 0;
    const int seg = (int)(a - a_first);
    if (seg >= 0 && seg < nseg) s += partial[((size_t)t * nseg + seg) * C + c];
  }
  out[a * ldp + poff + c] = s / (float)L;
}

// x [A, S] (fp32 or bf16) -> Conv1d 1->Cout (+folded BN, ReLU) -> flat rows bf16 (halo rows zeroed)
__global__ void k_stem_flat(const void* __restrict__ x, int x_dtype, int64_t A, int S, const float* __restrict__ w,
                            const float* __restrict__ shift, int k, int Cout, int relu, __nv_bfloat16* __restrict__ out,
                            int ldc, int coff, int Lp, int H0, int64_t R) {
  const int c8n = Cout >> 3;
  const int64_t total = R * c8n;
  const int half = k >> 1;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(idx % c8n);
    const int64_t row = idx / c8n;
    const int64_t rel = row - H0;
    const int64_t a = rel >= 0 ? rel / Lp : -1;
    const int l = rel >= 0 ? (int)(rel - a * Lp) : S;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    if (a >= 0 && a < A && l < S) {
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = shift[c8 * 8 + j];
      for (int t = 0; t < k; ++t) {
        const int li = l + t - half;
        if (li >= 0 && li < S) {
          const float xv = x_dtype == PAUT_BF16
                               ? __bfloat162float(static_cast<const __nv_bfloat16*>(x)[a * S + li])
                               : static_cast<const float*>(x)[a * S + li];
          const float4 w0 = __ldg(reinterpret_cast<const float4*>(w + t * Cout + c8 * 8));
          const float4 w1 = __ldg(reinterpret_cast<const float4*>(w + t * Cout + c8 * 8 + 4));
          acc[0] = fmaf(xv, w0.x, acc[0]); acc[1] = fmaf(xv, w0.y, acc[1]);
          acc[2] = fmaf(xv, w0.z, acc[2]); acc[3] = fmaf(xv, w0.w, acc[3]);
          acc[4] = fmaf(xv, w1.x, acc[4]); acc[5] = fmaf(xv, w1.y, acc[5]);
          acc[6] = fmaf(xv, w1.z, acc[6]); acc[7] = fmaf(xv, w1.w, acc[7]);
        }
      }
      if (relu) {
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaxf(acc[j], 0.f);
      }
    }
    uint32_t pk[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      __nv_bfloat162 h2 = __floats2bfloat162_rn(acc[2 * j], acc[2 * j + 1]);
      pk[j] = *reinterpret_cast<uint32_t*>(&h2);
    }
    *reinterpret_cast<uint4*>(out + row * ldc + coff + c8 * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  }
}

// Same stem, specialised: the tap count is a template parameter, a thread keeps the K x 8 weights of its 8
// channels in registers and produces 4 consecutive rows per step from one sliding window of K+3 input samples
// (independent FMA chains, 32-bit index arithmetic); consecutive threads write the consecutive 16-byte channel
// groups of one row (full 128-byte lines for 64 channels).
constexpr int STEM_RPT = 4;
template <int K>
__global__ void __launch_bounds__(256, 2) k_stem_flat_t(const void* __restrict__ x, int x_dtype, int A, int S,
                                                     const float* __restrict__ w, const float* __restrict__ shift,
                                                     int Cout, int relu, __nv_bfloat16* __restrict__ out, int ldc,
                                                     int coff, int Lp, int H0, int R) {
  const int c8n = Cout >> 3;
  const int c8 = threadIdx.x % c8n, rsub = threadIdx.x / c8n, rpb = (256 / c8n) * STEM_RPT;
  float wr[K][8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) sh[j] = __ldg(shift + c8 * 8 + j);
#pragma unroll
  for (int t = 0; t < K; ++t)
#pragma unroll
    for (int j = 0; j < 8; ++j) wr[t][j] = __ldg(w + t * Cout + c8 * 8 + j);
  const __nv_bfloat16* xb = static_cast<const __nv_bfloat16*>(x);
  const float* xf = static_cast<const float*>(x);
  const float lo = relu ? 0.f : -INFINITY;
  for (int row0 = blockIdx.x * rpb + rsub * STEM_RPT; row0 < R; row0 += gridDim.x * rpb) {
    const int rel = row0 - H0;
    const int a = rel >= 0 ? (int)((unsigned)rel / (unsigned)Lp) : -1;
    const int l = rel - a * Lp;                              // position of the first row inside its period
    const bool scan_ok = a >= 0 && a < A;
    // sliding window: samples l - K/2 .. l + RPT - 1 + K/2 of A-scan a (zero outside [0, S))
    float xw[K + STEM_RPT - 1];
    const size_t base = (size_t)(scan_ok ? a : 0) * S;
#pragma unroll
    for (int i = 0; i < K + STEM_RPT - 1; ++i) {
      const int li = l - K / 2 + i;
      xw[i] = 0.f;
      if (scan_ok && li >= 0 && li < S)
        xw[i] = x_dtype == PAUT_BF16 ? __uint_as_float((uint32_t)__ldg(reinterpret_cast<const unsigned short*>(xb) + base + li) << 16)
                                     : __ldg(xf + base + li);
    }
#pragma unroll
    for (int j = 0; j < STEM_RPT; ++j) {
      const int row = row0 + j;
      if (row >= R) break;
      // rows l+j >= S are halo rows (zero); a group that runs past the period (l+j >= Lp) starts the next A-scan
      // and takes the general path below
      float acc[8];
      const int lj = l + j;
      if (lj < Lp) {
        const bool ok = scan_ok && lj < S;
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[c] = ok ? sh[c] : 0.f;
        if (ok) {
#pragma unroll
          for (int t = 0; t < K; ++t)
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[c] = fmaf(xw[j + t], wr[t][c], acc[c]);
#pragma unroll
          for (int c = 0; c < 8; ++c) acc[c] = fmaxf(acc[c], lo);
        }
      } else {
        const int a2 = a + 1, l2 = lj - Lp;
        const bool ok = a2 >= 0 && a2 < A && l2 < S;
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[c] = ok ? sh[c] : 0.f;
        if (ok) {
          const size_t b2 = (size_t)a2 * S;
#pragma unroll
          for (int t = 0; t < K; ++t) {
            const int li = l2 + t - K / 2;
            float xv = 0.f;
            if (li >= 0 && li < S)
              xv = x_dtype == PAUT_BF16 ? __uint_as_float((uint32_t)__ldg(reinterpret_cast<const unsigned short*>(xb) + b2 + li) << 16)
                                        : __ldg(xf + b2 + li);
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[c] = fmaf(xv, wr[t][c], acc[c]);
          }
#pragma unroll
          for (int c = 0; c < 8; ++c) acc[c] = fmaxf(acc[c], lo);
        }
      }
      uint32_t pk[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        __nv_bfloat162 h2 = __floats2bfloat162_rn(acc[2 * c], acc[2 * c + 1]);
        pk[c] = *reinterpret_cast<uint32_t*>(&h2);
      }
      *reinterpret_cast<uint4*>(out + (size_t)row * ldc + coff + c8 * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
  }
}

__global__ void k_unflatten(const __nv_bfloat16* __restrict__ flat, int64_t A, int L, int Lp, int H0, int C,
                            float* __restrict__ out) {
  const int c8n = C >> 3;
  const int64_t total = A * L * c8n;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(idx % c8n);
    const int64_t al = idx / c8n;
    const int64_t a = al / L;
    const int l = (int)(al - a * L);
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(flat + (H0 + a * Lp + l) * C + c8 * 8));
    const uint32_t rr[4] = {v.x, v.y, v.z, v.w};
    float o[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const __nv_bfloat162 h2 = *reinterpret_cast<const __nv_bfloat162*>(&rr[j]);
      o[2 * j] = __low2float(h2);
      o[2 * j + 1] = __high2float(h2);
    }
    float4* dst = reinterpret_cast<float4*>(out + al * C + c8 * 8);
    dst[0] = make_float4(o[0], o[1], o[2], o[3]);
    dst[1] = make_float4(o[4], o[5], o[6], o[7]);
  }
}

uint16_t f2bf_(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  const uint32_t r = 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)((u + r) >> 16);
}

}  // namespace

// Tile plan of one convolution: N tile NT (<= 128, divides Cout) and input-channel block CB (64/32/16, divides
// Cin) such that the resident weights of one N tile plus the CT_STAGES-deep window ring fit the dynamic shared
// memory left beside the kernel's static arrays.  Prefers the widest N tile (the A windows are re-read once
// per N tile), then the deepest channel block.
constexpr size_t CT_DYN_SMEM = 232448 - 40 * 1024;
bool conv_tc_plan(int Cin, int taps, int Cout, int max_dil, int* NT_out, int* CB_out) {
  if (Cout % 16 != 0 || Cin % 16 != 0) return false;
  for (int nt = 128; nt >= 16; nt -= 16) {
    if (Cout % nt != 0) continue;
    const size_t w = (size_t)taps * Cin * nt * 2;
    for (int cb : {64, 32, 16}) {
      if (Cin % cb != 0) continue;
      const size_t a = (((size_t)(cb / 8) * (128 + (taps - 1) * max_dil) * 16) + 127) & ~(size_t)127;
      if (w + CT_STAGES * a <= CT_DYN_SMEM) {
        *NT_out = nt;
        *CB_out = cb;
        return true;
      }
    }
  }
  return false;
}

// weights [taps][Cin][Cout] fp32 (BN scale folded, as packed for the fp32 path) -> bf16 blocks
// [Cout/NT][Cin/CB][taps][CB/8][NT][8]
void conv_tc_pack(const float* w, int taps, int Cin, int Cout, int NT, int CB, std::vector<uint16_t>& out) {
  const int ncb = Cin / CB, chunks = CB / 8;
  out.assign((size_t)taps * Cin * Cout, 0);
  for (int co = 0; co < Cout; ++co)
    for (int ci = 0; ci < Cin; ++ci)
      for (int t = 0; t < taps; ++t) {
        const int nt = co / NT, r = co % NT, cb = ci / CB, ch = (ci % CB) / 8, e = ci % 8;
        const size_t idx = (((((size_t)nt * ncb + cb) * taps + t) * chunks + ch) * NT + r) * 8 + e;
        out[idx] = f2bf_(w[((size_t)t * Cin + ci) * Cout + co]);
      }
}

// Grouped variant (the two-stage encoder's four branches in one launch): group g is a conv of CB -> GN channels
// with taps[g] taps; weights [g][t][CB/8][GN][8] back to back.  w[g] is [taps[g]][CB][GN] fp32 (BN scale folded).
void conv_tc_pack_grouped(const float* const* w, const int* taps, int groups, int CB, int GN, std::vector<uint16_t>& out,
                          int* goff16) {
  const int chunks = CB / 8;
  size_t total = 0;
  for (int g = 0; g < groups; ++g) { goff16[g] = (int)(total / 8); total += (size_t)taps[g] * CB * GN; }
  out.assign(total, 0);
  for (int g = 0; g < groups; ++g)
    for (int t = 0; t < taps[g]; ++t)
      for (int ci = 0; ci < CB; ++ci)
        for (int co = 0; co < GN; ++co) {
          const size_t idx = (size_t)goff16[g] * 8 + ((((size_t)t * chunks + ci / 8) * GN + co) * 8 + ci % 8);
          out[idx] = f2bf_(w[g][((size_t)t * CB + ci) * GN + co]);
        }
}

size_t flat_rows(int64_t A, int L, int halo) { return (size_t)halo + (size_t)A * (L + halo); }

void op_conv_tc(Ctx& c, const ConvTcLaunch& a) {
  if (c.dry) return;
  ConvTcArgs p;
  p.in = static_cast<const __nv_bfloat16*>(a.in); p.Cin = a.Cin; p.Cout = a.Cout;
  p.Wp = static_cast<const __nv_bfloat16*>(a.Wp); p.shift = a.shift; p.taps = a.taps; p.dil = a.dil; p.pad = a.pad;
  p.NT = a.NT; p.CB = a.CB;
  // geometry: row(a, l) = H0 + a*Lp + l; defaults to the standard flat layout of (L, halo)
  p.L = a.L; p.Lp = a.Lp ? a.Lp : a.L + a.halo; p.H0 = a.Lp ? a.H0 : a.halo; p.A = a.A;
  p.R = (int64_t)p.H0 + a.A * (int64_t)p.Lp;
  PAUT_CHECK(p.R < (int64_t(1) << 31) - 256, PAUT_ERR_UNSUPPORTED, "conv_tc: too many rows in one launch");
  PAUT_CHECK(p.NT > 0 && p.CB > 0 && a.Cout % p.NT == 0 && a.Cin % p.CB == 0, PAUT_ERR_UNSUPPORTED,
             "conv_tc: channel counts must be multiples of 16 (no tile plan)");
  // zero rows on both sides of every A-scan must cover the taps' reach
  const int reach_l = a.pad * a.dil, reach_r = (a.taps - 1 - a.pad) * a.dil, gap = p.Lp - p.L;
  PAUT_CHECK(reach_l >= 0 && reach_r >= 0 && reach_l <= gap && reach_r <= gap && reach_l <= p.H0, PAUT_ERR_UNSUPPORTED,
             "conv_tc: padding wider than the zero rows between A-scans");
  p.nseg = p.Lp >= 127 ? 2 : 3;
  PAUT_CHECK(!a.pool_partial || (p.Lp >= 64 && 4 * p.nseg * p.NT <= 1024), PAUT_ERR_UNSUPPORTED,
             "conv_tc: pooled mean needs a row period >= 64 (and N tile <= 64 below 127)");
  p.skip_lo = a.skip_lo ? 1 : 0;
  p.grouped = a.groups > 0 ? (a.shared_input ? 2 : 1) : 0;
  p.ngroups = a.groups;
  for (int g = 0; g < 4; ++g) { p.gtaps[g] = a.gtaps[g]; p.goff[g] = a.goff[g]; p.gdil[g] = a.gdil[g]; }
  PAUT_CHECK(p.grouped != 1 || (a.groups == a.Cin / p.CB && a.groups <= 4 && p.NT == a.Cout && (p.NT / a.groups) % 16 == 0),
             PAUT_ERR_UNSUPPORTED, "conv_tc: grouped mode needs one channel block and >= 16 output columns per group");
  PAUT_CHECK(p.grouped != 2 || (a.Cin == p.CB && a.groups <= 4 && p.NT == a.Cout && (p.NT / a.groups) % 16 == 0),
             PAUT_ERR_UNSUPPORTED, "conv_tc: shared-input groups need a single channel block");
  PAUT_CHECK(!p.skip_lo || (a.Cin / p.CB) % 2 == 0, PAUT_ERR_INVALID, "conv_tc: skip_lo needs an even number of channel blocks");
  p.relu = a.relu ? 1 : 0; p.res = static_cast<const __nv_bfloat16*>(a.res); p.ldr = a.ldr;
  p.out = static_cast<__nv_bfloat16*>(a.out); p.ldc = a.ldc; p.coff = a.coff; p.pool = a.pool_partial;
  p.wrows = 128 + (a.taps - 1) * a.dil;
  p.a_stage_bytes = ((p.CB / 8) * p.wrows * 16 + 127) & ~127;
  p.w_bytes = a.taps * a.Cin * p.NT * 2;                   // all weights of one N tile
  if (p.grouped) {
    p.w_bytes = 0;
    for (int g = 0; g < a.groups; ++g) p.w_bytes += a.gtaps[g] * p.CB * (p.NT / a.groups) * 2;
  }
  p.num_tiles = (p.R + 127) / 128;
  const size_t smem = (size_t)p.w_bytes + (size_t)CT_STAGES * p.a_stage_bytes;
  PAUT_CHECK((int)smem <= c.smem_optin, PAUT_ERR_UNSUPPORTED, "conv_tc: weights + ring do not fit shared memory");
  // four compiled epilogue variants: (out), (out + residual), (pool), (out + pool [+ residual, null-checked])
  const bool v_out = p.out != nullptr, v_pool = p.pool != nullptr, v_res = p.res != nullptr;
  PAUT_CHECK(v_out || v_pool, PAUT_ERR_INVALID, "conv_tc: neither an output nor a pooled output was requested");
  PAUT_CHECK(!v_res || v_out, PAUT_ERR_UNSUPPORTED, "conv_tc: residual without an output tensor");
  void (*kern)(ConvTcArgs) = v_pool ? (v_out ? k_conv_tc<true, true, true> : k_conv_tc<false, true, false>)
                                    : (v_res ? k_conv_tc<true, false, true> : k_conv_tc<true, false, false>);
  smem_optin(c, kern);
  const int ntn = a.Cout / p.NT;
  int grid = (c.num_sms / ntn) * ntn;                       // every N tile gets the same number of CTAs
  if ((int64_t)grid > p.num_tiles * ntn) grid = (int)(p.num_tiles * ntn);
  static const bool debug = std::getenv("PAUT_CONV_DEBUG") != nullptr;
  p.dbg = nullptr;
  if (debug) { PAUT_CUDA(cudaMalloc(&p.dbg, 24 * sizeof(unsigned long long))); PAUT_CUDA(cudaMemset(p.dbg, 0, 24 * 8)); }
  kern<<<grid, CT_THREADS, smem, c.stream>>>(p);
  if (debug) {
    unsigned long long h[24];
    PAUT_CUDA(cudaMemcpy(h, p.dbg, sizeof(h), cudaMemcpyDeviceToHost));
    cudaFree(p.dbg);
    fprintf(stderr, "[conv probe] Cin %d Cout %d taps %d | loader: wait_empty %llu issue %llu per stage (%llu stages) | "
                    "mma: wait_acc_empty(total) %llu, per stage wait_full %llu issue %llu (%llu stages) | "
                    "epilogue: per tile wait_full %llu process %llu of which tmem_ld %llu (%llu tiles)\n",
            a.Cin, a.Cout, a.taps, h[0] / (h[2] + !h[2]), h[1] / (h[2] + !h[2]), h[2], h[8], h[9] / (h[11] + !h[11]),
            h[10] / (h[11] + !h[11]), h[11], h[16] / (h[18] + !h[18]), h[17] / (h[18] + !h[18]), h[19] / (h[18] + !h[18]), h[18]);
  }
  c.launched("conv_tc");
  if (a.pool_partial && a.pool_out) {
    const int64_t n = a.A * a.Cout;
    k_pool_finish<<<(unsigned)((n + 255) / 256), 256, 0, c.stream>>>(a.pool_partial, a.pool_out, a.ldp, a.poff, a.A,
                                                                    a.Cout, p.L, p.Lp, p.H0, p.nseg);
    c.launched("pool_finish");
  }
}

void op_unflatten(Ctx& c, const void* flat, int64_t A, int L, int halo, int C, float* out) {
  if (c.dry) return;
  PAUT_CHECK(C % 8 == 0, PAUT_ERR_UNSUPPORTED, "unflatten: C must be a multiple of 8");
  int64_t blocks = (A * L * (C / 8) + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  k_unflatten<<<(unsigned)blocks, 256, 0, c.stream>>>(static_cast<const __nv_bfloat16*>(flat), A, L, L + halo, halo, C, out);
  c.launched("unflatten");
}

void op_stem_flat(Ctx& c, const void* x, int x_dtype, int64_t A, int S, const float* w, const float* shift, int k,
                  int Cout, bool relu, void* out, int ldc, int coff, int halo) {
  if (c.dry) return;
  PAUT_CHECK(Cout % 8 == 0 && ldc % 8 == 0 && coff % 8 == 0, PAUT_ERR_UNSUPPORTED, "stem_flat: channels must be multiples of 8");
  const int64_t R = (int64_t)flat_rows(A, S, halo);
  const int c8n = Cout / 8;
  if (256 % c8n == 0 && R < (int64_t(1) << 31) - 4096 && (k == 3 || k == 5 || k == 7 || k == 11)) {
    const int rpb = (256 / c8n) * STEM_RPT;
    int64_t blocks = (R + rpb - 1) / rpb;
    if (blocks > (int64_t)c.num_sms * 16) blocks = (int64_t)c.num_sms * 16;
    __nv_bfloat16* o = static_cast<__nv_bfloat16*>(out);
#define PAUT_STEM(KK)                                                                                          \
  k_stem_flat_t<KK><<<(unsigned)blocks, 256, 0, c.stream>>>(x, x_dtype, (int)A, S, w, shift, Cout, relu ? 1 : 0, o, \
                                                            ldc, coff, S + halo, halo, (int)R)
    if (k == 3) PAUT_STEM(3);
    else if (k == 5) PAUT_STEM(5);
    else if (k == 7) PAUT_STEM(7);
    else PAUT_STEM(11);
#undef PAUT_STEM
  } else {
    const int64_t total = R * (Cout / 8);
    int64_t blocks = (total + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    k_stem_flat<<<(unsigned)blocks, 256, 0, c.stream>>>(x, x_dtype, A, S, w, shift, k, Cout, relu ? 1 : 0,
                                                        static_cast<__nv_bfloat16*>(out), ldc, coff, S + halo, halo, R);
  }
  c.launched("stem_flat");
}

}  // namespace paut
