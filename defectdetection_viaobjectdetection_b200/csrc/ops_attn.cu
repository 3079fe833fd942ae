// bf16-mode attention: softmax(Q K^T / sqrt(hd)) V per set with bf16 operands on the tensor cores
// (warp-level mma.sync m16n8k16, fp32 accumulate) and an fp32 exp2 softmax reduced with warp shuffles.
// The kernel is MUFU/CUDA-core bound (one exp per score), not tensor-pipe bound: head_dim is 16 or 32,
// so the QK^T and PV products are a few k-steps each and the accumulators stay in registers.
//
// CTA = one set (loops over heads), 4 warps; warp = 16 query rows at a time.  K and V^T of the current
// head live in shared memory as bf16; scores never leave registers (flash-style online softmax over
// key blocks of 64), so nothing of the N x N score matrix touches HBM unless the caller asks for the
// head-averaged weights (enhanced_model.py:197,246), which needs the set to fit one key block (N <= 64).
#include "common.cuh"

namespace paut {

namespace {

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

constexpr int KB = 64;          // keys per block (8 n-tiles)
constexpr int NT = KB / 8;

template <int HD, bool WEIGHTS>
__global__ void __launch_bounds__(128) k_attn_mma(const float* __restrict__ q, int ldq, const float* __restrict__ k,
                                                  int ldk, const float* __restrict__ v, int ldv,
                                                  float* __restrict__ out, int ldo, int Nq, int Nk, int H,
                                                  int kv_shift, float* __restrict__ avgw) {
  constexpr int KS = HD / 16;                 // k-steps of the QK^T product
  constexpr int DT = HD / 8;                  // n-tiles of the PV product
  extern __shared__ __align__(16) unsigned char smraw[];
  const int Nkp = (Nk + KB - 1) / KB * KB;    // keys padded to whole blocks
  const int ks_stride = HD + 8;               // bf16 elements per K row (padding breaks bank conflicts)
  const int vt_stride = Nkp + 8;
  __nv_bfloat16* Ks = reinterpret_cast<__nv_bfloat16*>(smraw);            // [Nkp][HD+8]
  __nv_bfloat16* Vt = Ks + (size_t)Nkp * ks_stride;                         // [HD][Nkp+8]
  float* Aw = reinterpret_cast<float*>(Vt + (size_t)HD * vt_stride);       // [Nq16][KB] (WEIGHTS only)

  const int64_t b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const float qscale = rsqrtf((float)HD) * 1.4426950408889634f;            // 1/sqrt(hd) * log2(e)
  const float invH = 1.f / (float)H;
  const int Nq16 = (Nq + 15) / 16 * 16;

  if (WEIGHTS)
    for (int i = tid; i < Nq16 * KB; i += 128) Aw[i] = 0.f;

  for (int h = 0; h < H; ++h) {
    __syncthreads();                                                         // previous head fully consumed
    // ---- stage K (row-major) and V^T of this head as bf16
    for (int i = tid; i < Nkp * (HD / 4); i += 128) {
      const int j = i / (HD / 4), c4 = i - j * (HD / 4);
      float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
      if (j < Nk) {
        const int js = kv_shift ? min(j + 1, Nk - 1) : j;
        const int64_t row = b * Nk + js;
        kv = __ldg(reinterpret_cast<const float4*>(k + row * ldk + h * HD + c4 * 4));
        vv = __ldg(reinterpret_cast<const float4*>(v + row * ldv + h * HD + c4 * 4));
      }
      uint2 pk = make_uint2(pack_bf16(kv.x, kv.y), pack_bf16(kv.z, kv.w));
      *reinterpret_cast<uint2*>(Ks + (size_t)j * ks_stride + c4 * 4) = pk;
      Vt[(size_t)(c4 * 4 + 0) * vt_stride + j] = __float2bfloat16_rn(vv.x);
      Vt[(size_t)(c4 * 4 + 1) * vt_stride + j] = __float2bfloat16_rn(vv.y);
      Vt[(size_t)(c4 * 4 + 2) * vt_stride + j] = __float2bfloat16_rn(vv.z);
      Vt[(size_t)(c4 * 4 + 3) * vt_stride + j] = __float2bfloat16_rn(vv.w);
    }
    __syncthreads();

    for (int r0 = warp * 16; r0 < Nq; r0 += 64) {
      // ---- Q fragments (pre-scaled by log2(e)/sqrt(hd)), straight from global
      uint32_t qa[KS][4];
      const int row_lo = r0 + g, row_hi = r0 + g + 8;
#pragma unroll
      for (int kk = 0; kk < KS; ++kk) {
        float2 x0 = make_float2(0.f, 0.f), x1 = x0, x2 = x0, x3 = x0;
        if (row_lo < Nq) {
          const float* p = q + (b * Nq + row_lo) * (int64_t)ldq + h * HD + kk * 16 + 2 * t;
          x0 = __ldg(reinterpret_cast<const float2*>(p));
          x2 = __ldg(reinterpret_cast<const float2*>(p + 8));
        }
        if (row_hi < Nq) {
          const float* p = q + (b * Nq + row_hi) * (int64_t)ldq + h * HD + kk * 16 + 2 * t;
          x1 = __ldg(reinterpret_cast<const float2*>(p));
          x3 = __ldg(reinterpret_cast<const float2*>(p + 8));
        }
        qa[kk][0] = pack_bf16(x0.x * qscale, x0.y * qscale);
        qa[kk][1] = pack_bf16(x1.x * qscale, x1.y * qscale);
        qa[kk][2] = pack_bf16(x2.x * qscale, x2.y * qscale);
        qa[kk][3] = pack_bf16(x3.x * qscale, x3.y * qscale);
      }
      float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
      float o[DT][4];
#pragma unroll
      for (int d = 0; d < DT; ++d) o[d][0] = o[d][1] = o[d][2] = o[d][3] = 0.f;

      for (int kb = 0; kb < Nkp; kb += KB) {
        float s[NT][4];
#pragma unroll
        for (int j = 0; j < NT; ++j) {
          s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
          const __nv_bfloat16* kr = Ks + (size_t)(kb + j * 8 + g) * ks_stride + 2 * t;
#pragma unroll
          for (int kk = 0; kk < KS; ++kk)
            mma_bf16_16816(s[j], qa[kk], *reinterpret_cast<const uint32_t*>(kr + kk * 16),
                           *reinterpret_cast<const uint32_t*>(kr + kk * 16 + 8));
        }
        if (kb + KB > Nk) {                                                   // mask the padded keys
#pragma unroll
          for (int j = 0; j < NT; ++j) {
            const int c0 = kb + j * 8 + 2 * t;
            if (c0 >= Nk) s[j][0] = s[j][2] = -INFINITY;
            if (c0 + 1 >= Nk) s[j][1] = s[j][3] = -INFINITY;
          }
        }
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
          mx0 = fmaxf(mx0, fmaxf(s[j][0], s[j][1]));
          mx1 = fmaxf(mx1, fmaxf(s[j][2], s[j][3]));
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
        const float c0 = fast_exp2(m0 - mn0), c1 = fast_exp2(m1 - mn1);       // 0 on the first block
        m0 = mn0;
        m1 = mn1;
        l0 *= c0;
        l1 *= c1;
#pragma unroll
        for (int d = 0; d < DT; ++d) { o[d][0] *= c0; o[d][1] *= c0; o[d][2] *= c1; o[d][3] *= c1; }
#pragma unroll
        for (int j = 0; j < NT; ++j) {
          s[j][0] = fast_exp2(s[j][0] - mn0);
          s[j][1] = fast_exp2(s[j][1] - mn0);
          s[j][2] = fast_exp2(s[j][2] - mn1);
          s[j][3] = fast_exp2(s[j][3] - mn1);
          l0 += s[j][0] + s[j][1];
          l1 += s[j][2] + s[j][3];
        }
        if (WEIGHTS) {                                                        // single key block: l is final
          float t0 = l0, t1 = l1;
          t0 += __shfl_xor_sync(0xffffffffu, t0, 1);
          t0 += __shfl_xor_sync(0xffffffffu, t0, 2);
          t1 += __shfl_xor_sync(0xffffffffu, t1, 1);
          t1 += __shfl_xor_sync(0xffffffffu, t1, 2);
          const float w0 = invH / t0, w1 = invH / t1;
#pragma unroll
          for (int j = 0; j < NT; ++j) {
            float* a0 = Aw + (size_t)(r0 + g) * KB + j * 8 + 2 * t;
            float* a1 = Aw + (size_t)(r0 + g + 8) * KB + j * 8 + 2 * t;
            a0[0] += s[j][0] * w0;
            a0[1] += s[j][1] * w0;
            a1[0] += s[j][2] * w1;
            a1[1] += s[j][3] * w1;
          }
        }
        // ---- O += P V : the score fragments of two adjacent n-tiles are exactly one A fragment
#pragma unroll
        for (int ks = 0; ks < NT / 2; ++ks) {
          uint32_t pa[4];
          pa[0] = pack_bf16(s[2 * ks][0], s[2 * ks][1]);
          pa[1] = pack_bf16(s[2 * ks][2], s[2 * ks][3]);
          pa[2] = pack_bf16(s[2 * ks + 1][0], s[2 * ks + 1][1]);
          pa[3] = pack_bf16(s[2 * ks + 1][2], s[2 * ks + 1][3]);
#pragma unroll
          for (int d = 0; d < DT; ++d) {
            const __nv_bfloat16* vr = Vt + (size_t)(d * 8 + g) * vt_stride + kb + ks * 16 + 2 * t;
            mma_bf16_16816(o[d], pa, *reinterpret_cast<const uint32_t*>(vr),
                           *reinterpret_cast<const uint32_t*>(vr + 8));
          }
        }
      }
      l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
      l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
      l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
      l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
      const float i0 = 1.f / l0, i1 = 1.f / l1;
#pragma unroll
      for (int d = 0; d < DT; ++d) {
        const int col = h * HD + d * 8 + 2 * t;
        if (row_lo < Nq)
          *reinterpret_cast<float2*>(out + (b * Nq + row_lo) * (int64_t)ldo + col) = make_float2(o[d][0] * i0, o[d][1] * i0);
        if (row_hi < Nq)
          *reinterpret_cast<float2*>(out + (b * Nq + row_hi) * (int64_t)ldo + col) = make_float2(o[d][2] * i1, o[d][3] * i1);
      }
    }
  }
  if (WEIGHTS) {
    __syncthreads();
    for (int i = tid; i < Nq * Nk; i += 128) {
      const int r = i / Nk, j = i - r * Nk;
      avgw[(b * Nq + r) * (int64_t)Nk + j] = Aw[(size_t)r * KB + j];
    }
  }
}

template <int HD, bool W>
void launch_attn(Ctx& c, const float* q, int ldq, const float* k, int ldk, const float* v, int ldv, float* out, int ldo,
                 int64_t B, int Nq, int Nk, int H, bool kv_shift, float* avgw) {
  const int Nkp = (Nk + KB - 1) / KB * KB;
  const int Nq16 = (Nq + 15) / 16 * 16;
  size_t smem = sizeof(__nv_bfloat16) * ((size_t)Nkp * (HD + 8) + (size_t)HD * (Nkp + 8));
  smem = (smem + 15) & ~size_t(15);
  if (W) smem += sizeof(float) * (size_t)Nq16 * KB;
  PAUT_CHECK((int)smem <= c.smem_optin, PAUT_ERR_UNSUPPORTED, "attention: set too long for shared memory");
  if (smem > 48 * 1024) smem_optin(c, k_attn_mma<HD, W>);
  PAUT_CHECK(B < (int64_t(1) << 31), PAUT_ERR_INVALID, "attention: too many sets");
  k_attn_mma<HD, W><<<(unsigned)B, 128, smem, c.stream>>>(q, ldq, k, ldk, v, ldv, out, ldo, Nq, Nk, H,
                                                          kv_shift ? 1 : 0, avgw);
  c.launched("attention_mma");
}

}  // namespace

bool attention_bf16_supported(int Nq, int Nk, int hd, bool want_weights) {
  if (hd != 16 && hd != 32) return false;
  if (want_weights && (Nk > KB || Nq > 64)) return false;
  return Nk <= 2048;
}

void op_attention_bf16(Ctx& c, const float* q, int ldq, const float* k, int ldk, const float* v, int ldv, float* out,
                       int ldo, int64_t B, int Nq, int Nk, int H, int hd, bool kv_shift, float* avgw) {
  if (c.dry) return;
  PAUT_CHECK(attention_bf16_supported(Nq, Nk, hd, avgw != nullptr), PAUT_ERR_UNSUPPORTED,
             "attention_bf16: unsupported shape");
  PAUT_CHECK(ldq % 2 == 0 && ldk % 4 == 0 && ldv % 4 == 0 && ldo % 2 == 0, PAUT_ERR_UNSUPPORTED,
             "attention_bf16: leading dimensions must be even (q, out) / multiples of 4 (k, v)");
  if (hd == 16) {
    if (avgw) launch_attn<16, true>(c, q, ldq, k, ldk, v, ldv, out, ldo, B, Nq, Nk, H, kv_shift, avgw);
    else launch_attn<16, false>(c, q, ldq, k, ldk, v, ldv, out, ldo, B, Nq, Nk, H, kv_shift, avgw);
  } else {
    if (avgw) launch_attn<32, true>(c, q, ldq, k, ldk, v, ldv, out, ldo, B, Nq, Nk, H, kv_shift, avgw);
    else launch_attn<32, false>(c, q, ldq, k, ldk, v, ldv, out, ldo, B, Nq, Nk, H, kv_shift, avgw);
  }
}

}  // namespace paut
