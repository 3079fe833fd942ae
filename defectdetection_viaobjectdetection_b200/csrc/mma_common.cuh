// warp-level mma.sync (m16n8k16, bf16 -> fp32) helpers shared by the attention and set-stage kernels
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace paut {
namespace mma {

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
// D (+)= A[16x16, row] * B[16x8, col]; fragment layout (g = lane>>2, t = lane&3):
//   a0 = A[g][2t,2t+1]  a1 = A[g+8][2t..]  a2 = A[g][2t+8..]  a3 = A[g+8][2t+8..]
//   b0 = B[k=2t,2t+1][n=g]  b1 = B[k=2t+8..][n=g]      c0,c1 = C[g][2t,2t+1]  c2,c3 = C[g+8][2t..]
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// same product with fp16 operands
__device__ __forceinline__ void mma_f16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// 2^(lo), 2^(hi) as a packed fp16 pair with ONE special-function instruction (half the MUFU work of two f32
// exponentials); the result is directly one register of an fp16 A fragment
__device__ __forceinline__ uint32_t exp2_pair_f16(float lo, float hi) {
  const __half2 x = __floats2half2_rn(lo, hi);
  uint32_t y;
  asm("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(*reinterpret_cast<const uint32_t*>(&x)));
  return y;
}
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// two adjacent 8-column accumulator tiles are exactly one 16-wide A fragment of the next product
__device__ __forceinline__ void c_to_a(const float (&c_lo)[4], const float (&c_hi)[4], uint32_t (&a)[4]) {
  a[0] = pack_bf16(c_lo[0], c_lo[1]);
  a[1] = pack_bf16(c_lo[2], c_lo[3]);
  a[2] = pack_bf16(c_hi[0], c_hi[1]);
  a[3] = pack_bf16(c_hi[2], c_hi[3]);
}

}  // namespace mma
}  // namespace paut
