// tcgen05 tensor-core GEMM:  C[M,N] = epilogue(A[M,K] * W[N,K]^T)   bf16 operands, fp32 accumulate in TMEM.
//
// One CTA = one 128-row M tile x one N tile (NT <= 256 columns).  K is consumed in blocks of 64:
// all four warps stage the A block (fp32 or bf16 activations -> bf16, canonical no-swizzle layout) and the
// pre-packed weight block into a 2-stage shared-memory ring, one elected thread issues the four K=16
// tcgen05.mma of the block and commits them to the stage's mbarrier; the ring lets the loads of block k+1
// overlap the MMAs of block k.  The accumulator ([128 x NT] fp32) lives in TMEM; the epilogue reads it back
// with tcgen05.ld (warp w <-> TMEM lanes 32w..32w+31 = rows), applies bias / activation / residual /
// row-table and writes fp32.
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "tc_common.cuh"

namespace paut {

using namespace tc;

namespace {

constexpr int BM = 128;
constexpr int BK = 64;             // K elements per pipeline stage (8 chunks of 8)
constexpr int STAGES = 2;

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
  int swap_lbo_sbo;                // debug: exchange the two descriptor strides (PAUT_TC_SWAP=1)
};

__global__ void __launch_bounds__(128) k_gemm_tc(GemmTcArgs p) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t mma_done[STAGES];
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x, lane = tid & 31, warp = uniform_warp_id();
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int ntile = blockIdx.y;
  const int NT = p.NT;
  const uint32_t a_stage_bytes = BM * BK * 2;                 // 16 KB
  const uint32_t b_stage_bytes = (uint32_t)NT * BK * 2;
  unsigned char* As = smem;                                   // [STAGES][8 chunks][128 rows][16 B]
  unsigned char* Bs = smem + STAGES * a_stage_bytes;          // [STAGES][8 chunks][NT rows][16 B]

  uint32_t ncols = 32;
  while ((int)ncols < NT) ncols <<= 1;
  if (warp == 0) tmem_alloc(&tmem_slot, ncols);
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) mbar_init(&mma_done[s], 1);
    fence_mbar_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_slot;
  const uint32_t idesc = make_idesc_bf16(BM, NT);

  const int nkb = (p.Kp + BK - 1) / BK;
  const int chunks_total = p.Kp / 8;
  const unsigned char* wp_tile = reinterpret_cast<const unsigned char*>(p.Wp) + (size_t)ntile * chunks_total * NT * 16;

  for (int kb = 0; kb < nkb; ++kb) {
    const int s = kb & 1;
    if (kb >= STAGES) mbar_wait(&mma_done[s], ((kb / STAGES) - 1) & 1);     // MMAs that read this stage are done
    const int k0 = kb * BK;
    const int nchunks = min(8, chunks_total - kb * 8);                       // chunks in this block (even)
    // ---- A block: warp w stages chunks w and w+4; lane = row within a group of 32 rows
    unsigned char* a_dst = As + s * a_stage_bytes;
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
      const int ch = warp + cc * 4;
      if (ch < nchunks) {
        const int k = k0 + ch * 8;
#pragma unroll
        for (int rg = 0; rg < 4; ++rg) {
          const int r = rg * 32 + lane;
          const int64_t m = m0 + r;
          float4 x0 = make_float4(0.f, 0.f, 0.f, 0.f), x1 = x0;
          if (m < p.M) {
            const float* src = p.A + m * p.lda + k;
            if (k + 4 <= p.K) x0 = __ldg(reinterpret_cast<const float4*>(src));
            if (k + 8 <= p.K) x1 = __ldg(reinterpret_cast<const float4*>(src + 4));
          }
          __nv_bfloat162 h0 = __floats2bfloat162_rn(x0.x, x0.y), h1 = __floats2bfloat162_rn(x0.z, x0.w);
          __nv_bfloat162 h2 = __floats2bfloat162_rn(x1.x, x1.y), h3 = __floats2bfloat162_rn(x1.z, x1.w);
          uint4 pk;
          pk.x = *reinterpret_cast<uint32_t*>(&h0);
          pk.y = *reinterpret_cast<uint32_t*>(&h1);
          pk.z = *reinterpret_cast<uint32_t*>(&h2);
          pk.w = *reinterpret_cast<uint32_t*>(&h3);
          *reinterpret_cast<uint4*>(a_dst + ch * (BM * 16) + r * 16) = pk;
        }
      }
    }
    // ---- B block: contiguous nchunks * NT * 16 bytes of the packed weights
    {
      const uint4* src = reinterpret_cast<const uint4*>(wp_tile + (size_t)kb * 8 * NT * 16);
      uint4* dst = reinterpret_cast<uint4*>(Bs + s * b_stage_bytes);
      const int n16 = nchunks * NT;
      for (int i = tid; i < n16; i += 128) dst[i] = __ldg(src + i);
    }
    fence_async_smem();
    __syncthreads();
    if (warp == 0 && elect_one()) {
      tc_fence_after();
      const uint32_t a_addr = smem_u32(As) + s * a_stage_bytes;
      const uint32_t b_addr = smem_u32(Bs) + s * b_stage_bytes;
      for (int ks = 0; ks < nchunks / 2; ++ks) {
        const uint64_t ad = p.swap_lbo_sbo ? make_desc(a_addr + ks * 2 * (BM * 16), 128, BM * 16)
                                           : make_desc(a_addr + ks * 2 * (BM * 16), BM * 16, 128);
        const uint64_t bd = p.swap_lbo_sbo ? make_desc(b_addr + ks * 2 * (NT * 16), 128, NT * 16)
                                           : make_desc(b_addr + ks * 2 * (NT * 16), NT * 16, 128);
        mma_bf16_ss(tmem_d, ad, bd, idesc, (kb > 0 || ks > 0) ? 1u : 0u);
      }
      mma_commit(&mma_done[s]);
    }
  }
  // ---- wait for the last commit (it covers every MMA issued before it)
  {
    const int last = nkb - 1;
    mbar_wait(&mma_done[last & 1], (last / STAGES) & 1);
  }
  tc_fence_after();

  // ---- epilogue: warp w owns rows 32w..32w+31 (= TMEM lanes)
  const int64_t m = m0 + warp * 32 + lane;
  const uint32_t t_row = tmem_d + ((uint32_t)(warp * 32) << 16);
  const int nbase = ntile * NT;
  for (int c0 = 0; c0 < NT; c0 += 16) {
    float v[16];
    tmem_ld16(t_row + c0, v);
    if (m < p.M) {
      const int n = nbase + c0;
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = tc_act(v[j] + (p.bias ? __ldg(p.bias + n + j) : 0.f), p.act) + p.act_eps;
      if (p.res) {
        const float4* r4 = reinterpret_cast<const float4*>(p.res + m * p.ldr + n);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 r = __ldg(r4 + j);
          v[4 * j] += r.x; v[4 * j + 1] += r.y; v[4 * j + 2] += r.z; v[4 * j + 3] += r.w;
        }
      }
      if (p.table) {
        const float4* r4 = reinterpret_cast<const float4*>(p.table + (m % p.table_mod) * p.N + n);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 r = __ldg(r4 + j);
          v[4 * j] += r.x; v[4 * j + 1] += r.y; v[4 * j + 2] += r.z; v[4 * j + 3] += r.w;
        }
      }
      float4* dst = reinterpret_cast<float4*>(p.C + m * p.ldc + p.coff + n);
#pragma unroll
      for (int j = 0; j < 4; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_d, ncols);
}

}  // namespace

// N tile: the largest multiple of 16 that divides N and is <= 256
int tc_pick_ntile(int N) {
  if (N % 16 != 0) return 0;
  for (int nt = 256; nt >= 16; nt -= 16)
    if (N % nt == 0) return nt;
  return 0;
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
  return a.Wp != nullptr && a.NT >= 16 && a.lda % 4 == 0 && a.ldc % 4 == 0 && a.coff % 4 == 0 &&
         (a.res == nullptr || a.ldr % 4 == 0) && a.K % 4 == 0;
}

void op_linear_tc(Ctx& c, const LinArgs& a) {
  if (c.dry) return;
  PAUT_CHECK(linear_tc_supported(a), PAUT_ERR_UNSUPPORTED, "linear_tc: unsupported shape");
  GemmTcArgs g;
  g.A = a.A; g.lda = a.lda; g.Wp = static_cast<const __nv_bfloat16*>(a.Wp); g.bias = a.bias; g.M = a.M; g.K = a.K;
  g.Kp = (a.K + 15) / 16 * 16; g.N = a.N; g.NT = a.NT; g.C = a.C; g.ldc = a.ldc; g.coff = a.coff; g.act = a.act;
  g.act_eps = a.act_eps; g.res = a.res; g.ldr = a.ldr; g.table = a.table; g.table_mod = a.table_mod;
  static const int swap = std::getenv("PAUT_TC_SWAP") ? atoi(std::getenv("PAUT_TC_SWAP")) : 0;
  g.swap_lbo_sbo = swap;
  const size_t smem = (size_t)STAGES * (BM * BK * 2 + (size_t)a.NT * BK * 2);
  if (smem > c.gemm_tc_smem_configured) {
    PAUT_CUDA(cudaFuncSetAttribute(k_gemm_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    c.gemm_tc_smem_configured = smem;
  }
  dim3 grid((unsigned)((a.M + BM - 1) / BM), a.N / a.NT);
  k_gemm_tc<<<grid, 128, smem, c.stream>>>(g);
  c.launched("linear_tc");
}

}  // namespace paut
