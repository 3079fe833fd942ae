// Debug / micro-benchmark entry points (not on any product path): tcgen05.mma cost per shape and the
// overlapping-chunk descriptor experiment that the conv kernels' tap addressing relies on.
#include "common.cuh"
#include "tc_common.cuh"

namespace paut {
using namespace tc;
namespace {

// mode 0: issue `reps` MMAs of shape 128 x N x 16 (operands: whatever is in shared memory), report cycles
// mode 2: A operand in TENSOR MEMORY (tcgen05.st -> TS-form MMA), B = identity: D[r][n] must equal A[r][n]
// mode 3: cycles per TS-form MMA 128 x N x 16
// mode 1: overlapped A view. A rows are 16 B apart, chunk 1 of row r = row r+1 (LBO = lbo bytes); B = identity
//         (16 x 16), so D[r][n] must equal n < 8 ? buf[r][n] : buf[r + lbo/16][n - 8].
__global__ void __launch_bounds__(128) k_debug_mma(int mode, int N, int reps, int lbo, int a_from_far, float* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, lane = tid & 31, warp = uniform_warp_id();
  unsigned char* A = smem;                    // up to 64 KB
  unsigned char* B = smem + 65536;            // up to 64 KB
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  // fill: A row r (16 B = 8 bf16) = r*8 + e scaled small; B = identity in the canonical [2 chunks][N rows][16B] layout
  for (int i = tid; i < 65536 / 2; i += 128) {
    const int r = i / 8, e = i % 8;
    reinterpret_cast<__nv_bfloat16*>(A)[i] = __float2bfloat16_rn((float)((r * 8 + e) % 251));
  }
  for (int i = tid; i < 65536 / 2; i += 128) reinterpret_cast<__nv_bfloat16*>(B)[i] = __float2bfloat16_rn(0.f);
  __syncthreads();
  if (mode == 1) {
    // B[n][k] = (n == k): chunk c (k = 8c..8c+7), row n at B + c*(N*16) + n*16
    for (int n = tid; n < 16; n += 128) {
      const int c = n / 8, e = n % 8;
      reinterpret_cast<__nv_bfloat16*>(B + c * (N * 16) + n * 16)[e] = __float2bfloat16_rn(1.f);
    }
  }
  if (mode == 2) {
    for (int n = tid; n < 16; n += 128) {
      const int c = n / 8, e = n % 8;
      reinterpret_cast<__nv_bfloat16*>(B + c * (N * 16) + n * 16)[e] = __float2bfloat16_rn(1.f);
    }
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  long long t0 = 0, t1 = 0;
  if (mode == 2 || mode == 3) {
    // A[r][k] = (r*16 + k) % 251, two bf16 per 32-bit column (even k in the low half), 8 columns at column 256
    uint32_t a[8];
    const int r = warp * 32 + lane;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      __nv_bfloat162 h2 = __floats2bfloat162_rn((float)((r * 16 + 2 * j) % 251), (float)((r * 16 + 2 * j + 1) % 251));
      a[j] = *reinterpret_cast<uint32_t*>(&h2);
    }
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(
                     tmem + ((uint32_t)(warp * 32) << 16) + 256),
                 "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7])
                 : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 0) {
      if (elect_one()) {
        const uint32_t idesc = make_idesc_bf16(128, N);
        const uint64_t bd = make_desc(smem_u32(B), N * 16, 128);
        t0 = clock64();
        for (int i = 0; i < reps; ++i)
          asm volatile(
              "{\n"
              ".reg .pred p;\n"
              "setp.ne.b32 p, %4, 0;\n"
              "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
              "}\n" ::"r"(tmem + (a_from_far ? (i & 1) * 128 : 0)),
              "r"(tmem + 256), "l"(bd), "r"(idesc), "r"(mode == 2 ? 0u : 1u)
              : "memory");
        mma_commit(&bar);
      }
      __syncwarp();
    }
  } else if (mode >= 6 && mode <= 9) {
    // descriptors that CHANGE between consecutive MMAs (a_from_far = mask, i & mask selects the variant; the descriptor
    // low words are precomputed so that the issuing thread does one AND + one add per MMA):
    //   mode 6: A start advances by 16 B (convolution taps);  mode 7: A start advances by 8 KB (different tiles);
    //   mode 8: B start advances by 2 KB;  mode 9: both A (16 B) and B (2 KB) advance
    if (warp == 0) {
      if (elect_one()) {
        const uint32_t idesc = make_idesc_bf16(128, N);
        const uint32_t hi = (uint32_t)(128 >> 4) | (1u << 14);
        const uint32_t a_lo = ((smem_u32(A) >> 4) & 0x3FFFu) | ((uint32_t)(lbo >> 4) << 16);
        const uint32_t b_lo = ((smem_u32(B) >> 4) & 0x3FFFu) | ((uint32_t)N << 16);
        const uint32_t a_step = mode == 6 || mode == 9 ? 1u : (mode == 7 ? 512u : 0u);
        const uint32_t b_step = mode == 8 || mode == 9 ? 128u : 0u;
        t0 = clock64();
        for (int i = 0; i < reps; ++i) {
          const uint32_t v = (uint32_t)i & (uint32_t)a_from_far;
          mma_bf16_ss2(tmem, a_lo + v * a_step, hi, b_lo + v * b_step, hi, idesc, 1u);
        }
        mma_commit(&bar);
      }
      __syncwarp();
    }
  } else if (mode == 4 || mode == 5) {
    // mode 4: cycles per SS-form MMA whose A descriptor starts `a_from_far` rows (16 B each) into the tile, chunk
    //         stride `lbo` bytes -- the convolution taps' shifted views (is a start that is not 128-byte aligned slower?)
    // mode 5: the same, cycling through the shifts 0 .. a_from_far (a convolution's tap loop)
    if (warp == 0) {
      if (elect_one()) {
        const uint32_t idesc = make_idesc_bf16(128, N);
        const uint64_t bd = make_desc(smem_u32(B), N * 16, 128);
        t0 = clock64();
        for (int i = 0; i < reps; ++i) {
          const int sh = mode == 4 ? a_from_far : (a_from_far > 0 ? i % (a_from_far + 1) : 0);
          mma_bf16_ss(tmem, make_desc(smem_u32(A) + sh * 16, lbo, 128), bd, idesc, 1u);
        }
        mma_commit(&bar);
      }
      __syncwarp();
    }
  } else
  if (warp == 0) {
    if (elect_one()) {
      const uint32_t idesc = make_idesc_bf16(128, N);
      const uint64_t ad = make_desc(smem_u32(A), mode == 1 ? lbo : 2048, 128);
      const uint64_t bd = make_desc(smem_u32(B), N * 16, 128);
      t0 = clock64();
      for (int i = 0; i < reps; ++i) mma_bf16_ss(tmem + (a_from_far ? (i & 1) * 256 : 0), ad, bd, idesc, mode == 1 ? 0u : 1u);
      mma_commit(&bar);
    }
    __syncwarp();
  }
  mbar_wait(&bar, 0);
  if (tid == 0) { t1 = clock64(); }
  tc_fence_after();
  if (mode == 0 || mode == 3 || mode >= 4) {
    if (tid == 0) out[0] = (float)(t1 - t0) / (float)reps;
  } else {
    float v[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16), v);
    for (int j = 0; j < 16; ++j) out[(warp * 32 + lane) * 16 + j] = v[j];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}
}  // namespace

void op_debug_mma(Ctx& c, int mode, int N, int reps, int lbo, int alt, float* out_dev) {
  PAUT_CUDA(cudaFuncSetAttribute(k_debug_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072));
  k_debug_mma<<<1, 128, 131072, c.stream>>>(mode, N, reps, lbo, alt, out_dev);
  c.launched("debug_mma");
}
}  // namespace paut
