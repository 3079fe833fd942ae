// tcgen05 tensor-core GEMM:  C[M,N] = epilogue(A[M,K] * W[N,K]^T)   bf16 operands, fp32 accumulate in TMEM.
//
// Persistent, warp-specialised (544 threads, one CTA per SM):
//   * a CTA owns ONE N tile (NT <= 256 columns) whose packed weights ([K/8 chunks][NT rows][16 B], K-major, no
//     swizzle) stay resident in shared memory, and walks a strided set of 128-row M tiles;
//     (when K * NT bf16 does not fit, e.g. the 2048-deep FFN output layer, the weights' K blocks ride the ring
//     next to the activation blocks instead -- `stream_b` -- which keeps the N tile wide: re-reading the
//     activations once per narrow N tile costs more than re-reading L2-resident weights once per M tile);
//   * 8 loader warps stream the fp32 activations of a [128 x 64] block with fully coalesced 128-bit loads (two
//     rows per warp instruction), convert to bf16 and store the canonical K-major operand into a 4-stage ring
//     (chunk stride 2048+16 B so that the 8-byte stores of a half-warp cover all banks); the loads of the next
//     block are issued before the current one is converted, so two blocks per SM are always in flight;
//   * one elected thread of the issuer warp runs the K loop (tcgen05.mma 128 x NT x 16), committing every block
//     to its ring slot's "empty" barrier and the last block of a tile to the accumulator's "full" barrier;
//   * 8 epilogue warps read the accumulator (two TMEM buffers: the MMAs of tile i+1 overlap the epilogue of
//     tile i) with tcgen05.ld, transpose 32x32 blocks through a swizzled per-warp shared-memory tile and, with
//     8 lanes per row, apply bias / activation / residual / row table and write fp32 rows as full 128-byte
//     lines (a row-per-lane store touches 32 lines per instruction and made the epilogue 16x longer than the MMAs).
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "tc_common.cuh"

namespace paut {

using namespace tc;

namespace {

constexpr int BM = 128;
constexpr int BK = 64;                       // K elements per ring stage (8 chunks of 8)
constexpr int GT_STAGES = 4;
constexpr int GT_EPI_WARPS = 8, GT_LOAD_WARPS = 8;
constexpr int GT_THREADS = 32 * (GT_EPI_WARPS + 1 + GT_LOAD_WARPS);
constexpr int GT_ALBO = 2048 + 16;           // chunk stride of the A operand (+16 B: bank skew)
constexpr int GT_ASTAGE = 8 * GT_ALBO;
constexpr int GT_TSM_BYTES = GT_EPI_WARPS * 4096;   // per-warp 32x32 fp32 transposition tiles
constexpr size_t GT_DYN_SMEM = 232448 - 2304;   // opt-in limit minus the static barriers / bias

__device__ __forceinline__ float tc_act(float v, int act) {
  switch (act) {
    case ACT_RELU: return fmaxf(v, 0.f);
    case ACT_GELU: return 0.5f * v * (1.f + erff(v * 0.70710678118654752440f));
    case ACT_SIGMOID: return 1.f / (1.f + expf(-v));
    case ACT_SOFTPLUS: return v > 20.f ? v : log1pf(expf(v));
    case ACT_TANH_HALF: return tanhf(v) * 0.5f + 0.5f;
    default: return v;
  }
}

struct GemmTcArgs {
  const float* A;                  // [M, lda] fp32
  int lda;
  const __nv_bfloat16* Wp;         // packed: [N/NT tiles][Kp/8 chunks][NT rows][8]
  const float* bias;
  int64_t M;
  int K, Kp, N, NT;                // Kp = K rounded up to 16
  float* C;
  int ldc, coff;
  int act;
  float act_eps;
  const float* res;
  int ldr;
  const float* table;
  int table_mod;
  uint32_t tmem_cols;              // power of two >= 2 * NT
  int stream_b;                    // weights too large to stay resident at this N tile: their K blocks ride the ring
};

__global__ void __launch_bounds__(GT_THREADS, 1) k_gemm_tc(GemmTcArgs p) {
  extern __shared__ __align__(128) unsigned char smem[];
  // all mbarriers in ONE array behind a pinned base address (as separate variables every use re-derived its shared-window
  // address with S2R + LEA): full[GT_STAGES] | empty[GT_STAGES] | acc_full[2] | acc_empty[2]
  __shared__ __align__(8) uint64_t bars[2 * GT_STAGES + 4];
  constexpr int FULL = 0, EMPTY = GT_STAGES, ACC_FULL = 2 * GT_STAGES, ACC_EMPTY = 2 * GT_STAGES + 2;
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(16) float bias_s[256];

  const int tid = threadIdx.x, lane = tid & 31, warp = uniform_warp_id();
  const int NT = p.NT;
  const int ntn = p.N / NT;
  const int nt = blockIdx.x % ntn;                                  // this CTA's N tile
  const int64_t tile0 = blockIdx.x / ntn, tstep = gridDim.x / ntn;
  const int64_t num_tiles = (p.M + BM - 1) / BM;
  const int chunks_total = p.Kp / 8;
  const int nkb = (p.Kp + BK - 1) / BK;
  const int w_bytes = p.stream_b ? 0 : chunks_total * NT * 16;
  const uint32_t stage_bytes = GT_ASTAGE + (p.stream_b ? (uint32_t)NT * 128 : 0u);   // A block (+ B block when streaming)
  unsigned char* Wres = smem;                                       // [chunks_total][NT][16 B]
  unsigned char* Aring = smem + ((w_bytes + 127) & ~127);           // [GT_STAGES][8 chunks][GT_ALBO] (+ [8][NT][16 B])
  unsigned char* tsm = Aring + GT_STAGES * stage_bytes;              // [GT_EPI_WARPS][32 rows][128 B], swizzled
  const uint32_t acc_stride = p.tmem_cols / 2;

  if (warp == 0) tmem_alloc(&tmem_slot, p.tmem_cols);
  if (tid == 0) {
    for (int s = 0; s < GT_STAGES; ++s) { mbar_init(&bars[FULL + s], GT_LOAD_WARPS); mbar_init(&bars[EMPTY + s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&bars[ACC_FULL + a], 1); mbar_init(&bars[ACC_EMPTY + a], GT_EPI_WARPS * 32); }
    fence_mbar_init();
  }
  {  // resident weights of this N tile: one contiguous block of the packed layout
    const uint4* src = reinterpret_cast<const uint4*>(p.Wp) + (size_t)nt * (chunks_total * NT);
    uint4* dst = reinterpret_cast<uint4*>(Wres);
    for (int i = tid; i < w_bytes / 16; i += GT_THREADS) dst[i] = __ldg(src + i);
  }
  for (int i = tid; i < NT; i += GT_THREADS) bias_s[i] = p.bias ? __ldg(p.bias + nt * NT + i) : 0.f;
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  uint32_t bar_u = smem_u32(bars);
  asm volatile("" : "+r"(bar_u));
  auto BA = [&](int idx) { return bar_u + (uint32_t)idx * 8u; };

  if (warp > GT_EPI_WARPS) {
    // ================= loaders =================
    const int lw = warp - (GT_EPI_WARPS + 1);                       // 0..7: rows lw*16 .. lw*16+15 of the tile
    const int j = lane & 15;                                        // K quad of the block: k = k0 + 4j .. 4j+3
    const int r0 = lw * 16 + (lane >> 4);                           // rows r0, r0+2, ..., r0+14
    const uint32_t ring0 = smem_u32(Aring);
    const uint32_t dst0 = ring0 + (uint32_t)(j >> 1) * GT_ALBO + (uint32_t)(j & 1) * 8 + (uint32_t)r0 * 16;
    auto load = [&](int64_t tile, int kb, float4 (&v)[8]) {
      const int k = kb * BK + 4 * j;
      const bool kok = k + 4 <= p.K;
      const int64_t m0 = tile * BM + r0;
      const float* src = p.A + m0 * p.lda + k;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        v[i] = (kok && m0 + 2 * i < p.M) ? __ldg(reinterpret_cast<const float4*>(src + (int64_t)(2 * i) * p.lda))
                                         : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    auto store = [&](int stage, const float4 (&v)[8]) {
      const uint32_t d = dst0 + (uint32_t)stage * stage_bytes;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        __nv_bfloat162 h0 = __floats2bfloat162_rn(v[i].x, v[i].y), h1 = __floats2bfloat162_rn(v[i].z, v[i].w);
        asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(d + (uint32_t)i * 32), "r"(*reinterpret_cast<uint32_t*>(&h0)),
                     "r"(*reinterpret_cast<uint32_t*>(&h1))
                     : "memory");
      }
    };
    // weight block of K block kb -> the B half of a ring stage (cp.async, 16 B per request; contiguous in the packed layout)
    const int lt = tid - (GT_EPI_WARPS + 1) * 32;                   // 0..255
    const unsigned char* wsrc = reinterpret_cast<const unsigned char*>(p.Wp) + (size_t)nt * chunks_total * NT * 16;
    auto load_b = [&](int kb, int stage) {
      const int nch = min(8, chunks_total - kb * 8);
      const uint32_t d = ring0 + (uint32_t)stage * stage_bytes + GT_ASTAGE;
      const unsigned char* src = wsrc + (size_t)kb * 8 * NT * 16;
      for (int i = lt; i < nch * NT; i += GT_LOAD_WARPS * 32)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d + (uint32_t)i * 16), "l"(src + (size_t)i * 16) : "memory");
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    int64_t tile = tile0;
    int kb = 0, stage = 0;
    uint32_t use = 0;
    float4 cur[8], nxt[8];
    bool have = tile < num_tiles;
    if (have) {
      if (p.stream_b) load_b(kb, stage);
      load(tile, kb, cur);
    }
    while (have) {
      int64_t tile_n = tile;
      int kb_n = kb + 1;
      if (kb_n == nkb) { kb_n = 0; tile_n += tstep; }
      const bool have_n = tile_n < num_tiles;
      int stage_n = stage + 1;
      uint32_t use_n = use;
      if (stage_n == GT_STAGES) { stage_n = 0; ++use_n; }
      if (p.stream_b) {
        // the next block's weights are requested now (its slot must be free), so that they land while this
        // block's activations are converted; this block's slot was waited for one iteration ago
        if (have_n) {
          if (use_n > 0) wait_a(BA(EMPTY + (stage_n)), (use_n - 1) & 1);
          load_b(kb_n, stage_n);
          load(tile_n, kb_n, nxt);
        }
        store(stage, cur);
        if (have_n) asm volatile("cp.async.wait_group 1;" ::: "memory");
        else asm volatile("cp.async.wait_group 0;" ::: "memory");
      } else {
        // next block in flight while this one is stored.  (Two stages in flight per thread -- each register set refilled
        // right after its stage is stored -- was measured in round 2: no change, 10.90 -> 10.86 ms on the two-stage model's
        // linears; after the residual prefetch below the loaders are not what bounds this kernel.)
        if (have_n) load(tile_n, kb_n, nxt);
        if (use > 0) wait_a(BA(EMPTY + (stage)), (use - 1) & 1);       // MMAs that read this slot are done
        store(stage, cur);
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) arrive_a(BA(FULL + (stage)));
      stage = stage_n; use = use_n;
#pragma unroll
      for (int i = 0; i < 8; ++i) cur[i] = nxt[i];
      tile = tile_n; kb = kb_n; have = have_n;
    }
  } else if (warp == GT_EPI_WARPS) {
    // ================= MMA issuer (one elected lane) =================
    const bool leader = elect_one();
    const uint32_t idesc = make_idesc_bf16(BM, NT);
    const uint32_t a_ring = smem_u32(Aring), w_addr = smem_u32(Wres);
    const uint64_t a_ks = (uint64_t)((2 * GT_ALBO) >> 4), b_ks = (uint64_t)((2 * NT * 16) >> 4);
    int stage = 0, it = 0;
    uint32_t use = 0;
    for (int64_t tile = tile0; tile < num_tiles; tile += tstep, ++it) {
      const int acc = it & 1;
      if (it >= 2) wait_a(BA(ACC_EMPTY + (acc)), ((it >> 1) - 1) & 1);   // epilogue drained this accumulator
      for (int kb = 0; kb < nkb; ++kb) {
        wait_a(BA(FULL + (stage)), use & 1);
        if (leader) {
          tc_fence_after();
          const int nch = min(8, chunks_total - kb * 8);             // chunks of this block (even)
          uint64_t ad = make_desc(a_ring + (uint32_t)stage * stage_bytes, GT_ALBO, 128);
          uint64_t bd = p.stream_b ? make_desc(a_ring + (uint32_t)stage * stage_bytes + GT_ASTAGE, NT * 16, 128)
                                   : make_desc(w_addr + (uint32_t)(kb * 8 * NT) * 16, NT * 16, 128);
          const uint32_t d = tmem + acc * acc_stride;
          for (int ks = 0; ks < nch / 2; ++ks) {
            mma_bf16_ss(d, ad, bd, idesc, (kb | ks) ? 1u : 0u);
            ad += a_ks;
            bd += b_ks;
          }
          commit_a(BA(EMPTY + (stage)));
          if (kb == nkb - 1) commit_a(BA(ACC_FULL + (acc)));
        }
        __syncwarp();
        if (++stage == GT_STAGES) { stage = 0; ++use; }
      }
    }
  } else {
    // ================= epilogue: warp w owns TMEM lanes 32*(w%4).. and the 32-column passes w/4, w/4+2, ... ====
    const int q = warp & 3, half = warp >> 2;
    const int rr = lane >> 3, cg = lane & 7;               // transposed domain: rows rr + 4i, columns 4cg..4cg+3
    const uint32_t tw = smem_u32(tsm) + (uint32_t)warp * 4096;
    uint32_t st_addr[8], ld_addr[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      st_addr[c] = tw + (uint32_t)lane * 128 + (uint32_t)((c ^ (lane & 7)) * 16);
      const int rl = rr + 4 * c;
      ld_addr[c] = tw + (uint32_t)rl * 128 + (uint32_t)((cg ^ (rl & 7)) * 16);
    }
    const int npass = NT > 32 * half ? (NT - 32 * half + 63) / 64 : 0;
    int it = 0;
    for (int64_t tile = tile0; tile < num_tiles; tile += tstep, ++it) {
      const int acc = it & 1;
      wait_a(BA(ACC_FULL + (acc)), (it >> 1) & 1);
      tc_fence_after();
      if (npass == 0) {                                      // narrow N tile: only keep the barrier phases in step
        tc_fence_before();
        arrive_a(BA(ACC_EMPTY + (acc)));
      }
      const int64_t m0 = tile * BM + q * 32 + rr;
      const uint32_t t_row = tmem + ((uint32_t)(q * 32) << 16) + acc * acc_stride;
      for (int k = 0; k < npass; ++k) {
        const int c0 = 32 * half + 64 * k;
        // the residual rows of this pass are requested FIRST: their latency then hides behind the TMEM load and the
        // transposition (requested at the point of use, each of the eight adds waited for its own load -- ncu: 30 % of the
        // kernel's stall samples on the 512 -> 128 + residual layer, profiles/r02/3j_ncu_gemm_tc_two_stage.txt)
        float4 rres[8];
        const bool has_res = p.res != nullptr && c0 + 4 * cg < NT;
        if (has_res) {
          const float* rp = p.res + (int64_t)(nt * NT + c0 + 4 * cg);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int64_t m = m0 + 4 * i;
            rres[i] = m < p.M ? __ldg(reinterpret_cast<const float4*>(rp + m * p.ldr)) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
        uint32_t t32[32];
        tmem_ld32(t_row + c0, t32);
        if (k == npass - 1) {                                // accumulator drained: the MMAs of tile it+2 may start
          tc_fence_before();
          arrive_a(BA(ACC_EMPTY + (acc)));
        }
#pragma unroll
        for (int c = 0; c < 8; ++c)
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(st_addr[c]), "r"(t32[4 * c]), "r"(t32[4 * c + 1]),
                       "r"(t32[4 * c + 2]), "r"(t32[4 * c + 3])
                       : "memory");
        __syncwarp();
        const int col = c0 + 4 * cg;                         // this lane's 4 columns inside the N tile
        if (col < NT) {
          const int n = nt * NT + col;
          const float4 b4 = *reinterpret_cast<const float4*>(bias_s + col);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int64_t m = m0 + 4 * i;
            float4 t;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(t.x), "=f"(t.y), "=f"(t.z), "=f"(t.w) : "r"(ld_addr[i]));
            if (m < p.M) {
              t.x = tc_act(t.x + b4.x, p.act) + p.act_eps; t.y = tc_act(t.y + b4.y, p.act) + p.act_eps;
              t.z = tc_act(t.z + b4.z, p.act) + p.act_eps; t.w = tc_act(t.w + b4.w, p.act) + p.act_eps;
              if (has_res) { t.x += rres[i].x; t.y += rres[i].y; t.z += rres[i].z; t.w += rres[i].w; }
              if (p.table) {
                const float4 r = __ldg(reinterpret_cast<const float4*>(p.table + (m % p.table_mod) * p.N + n));
                t.x += r.x; t.y += r.y; t.z += r.z; t.w += r.w;
              }
              *reinterpret_cast<float4*>(p.C + m * p.ldc + p.coff + n) = t;
            }
          }
        }
        __syncwarp();                                        // the tile is rewritten by the next pass
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, p.tmem_cols);
}

}  // namespace

// N tile and mode.  Resident mode: the largest multiple of 16 that divides N, is <= 256 and whose weights (Kp * NT
// bf16) fit the shared memory beside the activation ring.  If that forces more N tiles than the widest legal tile
// would need, the weights are streamed through the ring instead (stream = 1) and the tile stays wide.
int tc_pick_ntile(int N, int K, int* stream) {
  if (stream) *stream = 0;
  if (N % 16 != 0) return 0;
  const size_t Kp = (size_t)(K + 15) / 16 * 16;
  int wide = 0, resident = 0;
  for (int nt = 256; nt >= 16; nt -= 16) {
    if (N % nt != 0) continue;
    if (!wide && (size_t)GT_STAGES * (GT_ASTAGE + (size_t)nt * 128) + GT_TSM_BYTES <= GT_DYN_SMEM) wide = nt;
    if (!resident && ((Kp * nt * 2 + 127) & ~(size_t)127) + (size_t)GT_STAGES * GT_ASTAGE + GT_TSM_BYTES <= GT_DYN_SMEM) resident = nt;
  }
  if (stream && wide > resident) {
    *stream = 1;
    return wide;
  }
  return resident;
}

// Host-side packing of an nn.Linear weight [N][K] (fp32) into the chunked K-major bf16 layout.
void tc_pack_weight(const float* W, int N, int K, int NT, std::vector<uint16_t>& out, int* Kp_out) {
  const int Kp = (K + 15) / 16 * 16;
  const int chunks = Kp / 8;
  out.assign((size_t)N * Kp, 0);
  auto bf16 = [](float f) -> uint16_t {
    uint32_t u;
    memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);   // NaN
    const uint32_t r = 0x7fffu + ((u >> 16) & 1u);                                 // round to nearest even
    return (uint16_t)((u + r) >> 16);
  };
  for (int n = 0; n < N; ++n) {
    const int tile = n / NT, r = n % NT;
    for (int k = 0; k < K; ++k) {
      const int ch = k / 8, e = k % 8;
      out[(((size_t)tile * chunks + ch) * NT + r) * 8 + e] = bf16(W[(size_t)n * K + k]);
    }
  }
  *Kp_out = Kp;
}

bool linear_tc_supported(const LinArgs& a) {
  return a.Wp != nullptr && a.NT >= 16 && a.NT <= 256 && a.N % a.NT == 0 && a.lda % 4 == 0 && a.ldc % 4 == 0 && a.coff % 4 == 0 &&
         (a.res == nullptr || a.ldr % 4 == 0) && a.K % 4 == 0;
}

void op_linear_tc(Ctx& c, const LinArgs& a) {
  if (c.dry) return;
  PAUT_CHECK(linear_tc_supported(a), PAUT_ERR_UNSUPPORTED, "linear_tc: unsupported shape");
  GemmTcArgs g;
  g.A = a.A; g.lda = a.lda; g.Wp = static_cast<const __nv_bfloat16*>(a.Wp); g.bias = a.bias; g.M = a.M; g.K = a.K;
  g.Kp = (a.K + 15) / 16 * 16; g.N = a.N; g.NT = a.NT; g.C = a.C; g.ldc = a.ldc; g.coff = a.coff; g.act = a.act;
  g.act_eps = a.act_eps; g.res = a.res; g.ldr = a.ldr; g.table = a.table; g.table_mod = a.table_mod;
  g.tmem_cols = 64;                                        // the epilogue reads 32-column blocks: >= 32 columns per accumulator
  while ((int)g.tmem_cols < 2 * a.NT) g.tmem_cols <<= 1;
  g.stream_b = a.stream_b ? 1 : 0;
  const size_t smem = g.stream_b ? (size_t)GT_STAGES * (GT_ASTAGE + (size_t)a.NT * 128) + GT_TSM_BYTES
                                 : (((size_t)g.Kp * a.NT * 2 + 127) & ~(size_t)127) + (size_t)GT_STAGES * GT_ASTAGE + GT_TSM_BYTES;
  PAUT_CHECK(smem <= GT_DYN_SMEM && (int)smem <= c.smem_optin, PAUT_ERR_UNSUPPORTED,
             "linear_tc: resident weights do not fit shared memory (N tile too wide for this K)");
  smem_optin(c, k_gemm_tc);
  const int ntn = a.N / a.NT;
  const int64_t tiles = (a.M + BM - 1) / BM;
  int64_t grid = (int64_t)(c.num_sms / ntn) * ntn;          // every N tile gets the same number of CTAs
  if (grid < ntn) grid = ntn;
  if (grid > tiles * ntn) grid = tiles * ntn;
  k_gemm_tc<<<(unsigned)grid, GT_THREADS, smem, c.stream>>>(g);
  c.launched("linear_tc");
}

}  // namespace paut
