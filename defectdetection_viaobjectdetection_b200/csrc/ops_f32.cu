// fp32 CUDA-core operators: the exact-arithmetic path (PAUT_PRECISION_FP32) and the per-set
// sequence stage of every model.  All activations are channels-last, fp32.
#include "common.cuh"

namespace paut {

// ------------------------------------------------------------------------------------------ helpers
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }
__device__ __forceinline__ float apply_act(float v, int act) {
  switch (act) {
    case ACT_RELU: return fmaxf(v, 0.f);
    case ACT_GELU: return 0.5f * v * (1.f + erff(v * 0.70710678118654752440f));   // exact-erf GELU
    case ACT_SIGMOID: return sigmoidf_(v);
    case ACT_SOFTPLUS: return v > 20.f ? v : log1pf(expf(v));                     // beta 1, threshold 20
    case ACT_TANH_HALF: return tanhf(v) * 0.5f + 0.5f;
    default: return v;
  }
}
__device__ __forceinline__ float ld_any(const void* p, int dtype, int64_t i) {
  return dtype == PAUT_BF16 ? __bfloat162float(static_cast<const __nv_bfloat16*>(p)[i])
                            : static_cast<const float*>(p)[i];
}
static inline int grid_for(int64_t n, int block, int cap = 148 * 32) {
  int64_t g = (n + block - 1) / block;
  if (g < 1) g = 1;
  return (int)(g > cap ? cap : g);
}

// ------------------------------------------------------------------------------------------ casts
__global__ void k_to_f32(const void* __restrict__ x, int dtype, float* __restrict__ out, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = ld_any(x, dtype, i);
}
void op_to_f32(Ctx& c, const void* x, int x_dtype, float* out, int64_t n) {
  if (c.dry) return;
  k_to_f32<<<grid_for(n, 256), 256, 0, c.stream>>>(x, x_dtype, out, n);
  c.launched("to_f32");
}

__global__ void k_to_bf16(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int64_t n4) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    reinterpret_cast<uint2*>(out)[i] = make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
  }
}
void op_to_bf16(Ctx& c, const float* x, void* out, int64_t n) {
  if (c.dry) return;
  PAUT_CHECK(n % 4 == 0, PAUT_ERR_UNSUPPORTED, "to_bf16: element count must be a multiple of 4");
  k_to_bf16<<<grid_for(n / 4, 256), 256, 0, c.stream>>>(x, static_cast<__nv_bfloat16*>(out), n / 4);
  c.launched("to_bf16");
}

// [B,S,N] -> [B,N,S] through a 32x32 shared tile (coalesced on both sides)
__global__ void k_transpose_sn(const void* __restrict__ x, int dtype, float* __restrict__ out, int S, int N) {
  __shared__ float tile[32][33];
  const int64_t b = blockIdx.z;
  const int s0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    int s = s0 + r, n = n0 + threadIdx.x;
    if (s < S && n < N) tile[r][threadIdx.x] = ld_any(x, dtype, (b * S + s) * (int64_t)N + n);
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    int n = n0 + r, s = s0 + threadIdx.x;
    if (s < S && n < N) out[(b * N + n) * (int64_t)S + s] = tile[threadIdx.x][r];
  }
}
void op_transpose_sn(Ctx& c, const void* x, int x_dtype, float* out, int64_t B, int S, int N) {
  if (c.dry) return;
  for (int64_t b0 = 0; b0 < B; b0 += 65535) {
    int64_t nb = B - b0 < 65535 ? B - b0 : 65535;
    dim3 grid((N + 31) / 32, (S + 31) / 32, (unsigned)nb);
    size_t esz = x_dtype == PAUT_BF16 ? 2 : 4;
    k_transpose_sn<<<grid, dim3(32, 8), 0, c.stream>>>(static_cast<const char*>(x) + b0 * S * N * esz, x_dtype,
                                                      out + b0 * S * N, S, N);
    c.launched("transpose_sn");
  }
}

// ------------------------------------------------------------------------------------------ stem conv
// One thread = one position x 4 output channels.  Weights [k][Cout] stay in L1.
__global__ void k_stem_conv(const float* __restrict__ x, int64_t A, int S, const float* __restrict__ w,
                            const float* __restrict__ shift, int k, int Cout, int relu, float* __restrict__ out) {
  const int c4n = Cout >> 2;
  const int64_t total = A * S * c4n;
  const int half = k >> 1;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int c4 = (int)(idx % c4n);
    const int64_t pos = idx / c4n;
    const int l = (int)(pos % S);
    const float* xr = x + (pos - l);
    float4 acc = *reinterpret_cast<const float4*>(shift + c4 * 4);
    for (int t = 0; t < k; ++t) {
      const int li = l + t - half;
      if (li >= 0 && li < S) {
        const float xv = __ldg(xr + li);
        const float4 wv = __ldg(reinterpret_cast<const float4*>(w + t * Cout + c4 * 4));
        acc.x = fmaf(xv, wv.x, acc.x);
        acc.y = fmaf(xv, wv.y, acc.y);
        acc.z = fmaf(xv, wv.z, acc.z);
        acc.w = fmaf(xv, wv.w, acc.w);
      }
    }
    if (relu) {
      acc.x = fmaxf(acc.x, 0.f); acc.y = fmaxf(acc.y, 0.f); acc.z = fmaxf(acc.z, 0.f); acc.w = fmaxf(acc.w, 0.f);
    }
    *reinterpret_cast<float4*>(out + pos * Cout + c4 * 4) = acc;
  }
}
void op_stem_conv(Ctx& c, const float* x, int64_t A, int S, const float* w, const float* shift, int k, int Cout,
                  bool relu, float* out) {
  if (c.dry) return;
  PAUT_CHECK(Cout % 4 == 0, PAUT_ERR_UNSUPPORTED, "stem conv: Cout must be a multiple of 4");
  k_stem_conv<<<grid_for(A * S * (Cout / 4), 256), 256, 0, c.stream>>>(x, A, S, w, shift, k, Cout, relu ? 1 : 0, out);
  c.launched("stem_conv");
}

// ------------------------------------------------------------------------------------------ conv (implicit GEMM)
// CTA = one A-scan x BN output channels, loops over Lout in tiles of 64 positions, so the pooled mean is
// reduced in a fixed order (no atomics: results are bit-identical however the volume is sharded).
template <int BN>
__global__ void __launch_bounds__(256) k_conv_f32(ConvArgs p) {
  constexpr int BM = 64, BK = 16;
  constexpr int TXN = BN / 4;          // threads along n
  constexpr int TYN = 256 / TXN;       // threads along m
  constexpr int TM = BM / TYN;         // rows per thread (4 for BN=64, 2 for BN=32)
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN];
  __shared__ float Ps[TYN][BN];

  const int64_t a = blockIdx.x;
  const int n0 = blockIdx.y * BN;
  const int tid = threadIdx.x;
  const int tx = tid % TXN, ty = tid / TXN;
  const float* in = p.in + a * (int64_t)p.Lin * p.Cin;
  const int a_row = tid >> 2, a_c4 = tid & 3;             // A tile load: 64 rows x 4 float4
  const int b_row = tid / (BN / 4), b_c4 = tid % (BN / 4);  // B tile load: 16 rows x BN/4 float4

  float4 shift4 = *reinterpret_cast<const float4*>(p.shift + n0 + tx * 4);
  float psum[4] = {0.f, 0.f, 0.f, 0.f};

  for (int l0 = 0; l0 < p.Lout; l0 += BM) {
    float acc[TM][4];
#pragma unroll
    for (int i = 0; i < TM; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;

    for (int t = 0; t < p.taps; ++t) {
      const int lo = l0 + a_row;
      const int li = lo * p.stride + t * p.dil - p.pad;
      const bool a_ok = (lo < p.Lout) && (li >= 0) && (li < p.Lin);
      for (int c0 = 0; c0 < p.Cin; c0 += BK) {
        float4 av = make_float4(0.f, 0.f, 0.f, 0.f);
        if (a_ok) av = __ldg(reinterpret_cast<const float4*>(in + (int64_t)li * p.Cin + c0 + a_c4 * 4));
        float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (b_row < BK)
          bv = __ldg(reinterpret_cast<const float4*>(p.w + ((int64_t)t * p.Cin + c0 + b_row) * p.Cout + n0 + b_c4 * 4));
        __syncthreads();
        As[a_c4 * 4 + 0][a_row] = av.x;
        As[a_c4 * 4 + 1][a_row] = av.y;
        As[a_c4 * 4 + 2][a_row] = av.z;
        As[a_c4 * 4 + 3][a_row] = av.w;
        if (b_row < BK) *reinterpret_cast<float4*>(&Bs[b_row][b_c4 * 4]) = bv;
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
          const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
          float av_[TM];
#pragma unroll
          for (int i = 0; i < TM; ++i) av_[i] = As[kk][ty * TM + i];
#pragma unroll
          for (int i = 0; i < TM; ++i) {
            acc[i][0] = fmaf(av_[i], b.x, acc[i][0]);
            acc[i][1] = fmaf(av_[i], b.y, acc[i][1]);
            acc[i][2] = fmaf(av_[i], b.z, acc[i][2]);
            acc[i][3] = fmaf(av_[i], b.w, acc[i][3]);
          }
        }
      }
    }
    // epilogue for this tile of positions
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      const int l = l0 + ty * TM + i;
      if (l < p.Lout) {
        float4 v = make_float4(acc[i][0] + shift4.x, acc[i][1] + shift4.y, acc[i][2] + shift4.z, acc[i][3] + shift4.w);
        const int64_t row = a * p.Lout + l;
        if (p.res) {
          const float4 r = __ldg(reinterpret_cast<const float4*>(p.res + row * p.ldr + n0 + tx * 4));
          v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
        }
        if (p.relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
        if (p.out) *reinterpret_cast<float4*>(p.out + row * p.ldc + p.coff + n0 + tx * 4) = v;
        psum[0] += v.x; psum[1] += v.y; psum[2] += v.z; psum[3] += v.w;
      }
    }
  }
  if (p.pool) {
    __syncthreads();
    *reinterpret_cast<float4*>(&Ps[ty][tx * 4]) = make_float4(psum[0], psum[1], psum[2], psum[3]);
    __syncthreads();
    if (tid < BN) {
      float s = 0.f;
      for (int r = 0; r < TYN; ++r) s += Ps[r][tid];
      p.pool[a * p.ldp + p.poff + n0 + tid] = s / (float)p.Lout;
    }
  }
}
void op_conv(Ctx& c, const ConvArgs& a) {
  if (c.dry) return;
  PAUT_CHECK(a.Cin % 16 == 0, PAUT_ERR_UNSUPPORTED, "conv: Cin must be a multiple of 16");
  PAUT_CHECK(a.Cout % 32 == 0, PAUT_ERR_UNSUPPORTED, "conv: Cout must be a multiple of 32");
  PAUT_CHECK(a.A > 0 && a.A < (int64_t(1) << 31), PAUT_ERR_INVALID, "conv: bad A");
  if (a.Cout % 64 == 0) {
    k_conv_f32<64><<<dim3((unsigned)a.A, a.Cout / 64), 256, 0, c.stream>>>(a);
  } else {
    k_conv_f32<32><<<dim3((unsigned)a.A, a.Cout / 32), 256, 0, c.stream>>>(a);
  }
  c.launched("conv_f32");
}

// ------------------------------------------------------------------------------------------ linear
// C = act(A W^T + b) [+ res] [+ table[m % mod]]   64x64x16 tiles, 4x4 per thread.
__global__ void __launch_bounds__(256) k_linear_f32(LinArgs p) {
  constexpr int BM = 64, BN = 64, BK = 16;
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int a_row = tid >> 2, a_c4 = tid & 3;
  const int b_row = tid >> 4, b_c4 = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
  const int64_t am = m0 + a_row;
  for (int k0 = 0; k0 < p.K; k0 += BK) {
    float4 av = make_float4(0.f, 0.f, 0.f, 0.f);
    const int ak = k0 + a_c4 * 4;
    if (am < p.M && ak < p.K) av = __ldg(reinterpret_cast<const float4*>(p.A + am * p.lda + ak));
    float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
    const int bk = k0 + b_row, bn = n0 + b_c4 * 4;
    if (bk < p.K && bn < p.N) bv = __ldg(reinterpret_cast<const float4*>(p.Wt + (int64_t)bk * p.N + bn));
    __syncthreads();
    As[a_c4 * 4 + 0][a_row] = av.x;
    As[a_c4 * 4 + 1][a_row] = av.y;
    As[a_c4 * 4 + 2][a_row] = av.z;
    As[a_c4 * 4 + 3][a_row] = av.w;
    *reinterpret_cast<float4*>(&Bs[b_row][b_c4 * 4]) = bv;
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float a_[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        acc[i][0] = fmaf(a_[i], b.x, acc[i][0]);
        acc[i][1] = fmaf(a_[i], b.y, acc[i][1]);
        acc[i][2] = fmaf(a_[i], b.z, acc[i][2]);
        acc[i][3] = fmaf(a_[i], b.w, acc[i][3]);
      }
    }
  }
  const int n = n0 + tx * 4;
  if (n >= p.N) return;
  float4 b4 = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias + n)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
    float v[4] = {acc[i][0] + b4.x, acc[i][1] + b4.y, acc[i][2] + b4.z, acc[i][3] + b4.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = apply_act(v[j], p.act) + p.act_eps;
    if (p.res) {
      const float4 r = __ldg(reinterpret_cast<const float4*>(p.res + m * p.ldr + n));
      v[0] += r.x; v[1] += r.y; v[2] += r.z; v[3] += r.w;
    }
    if (p.table) {
      const float4 r = __ldg(reinterpret_cast<const float4*>(p.table + (m % p.table_mod) * p.N + n));
      v[0] += r.x; v[1] += r.y; v[2] += r.z; v[3] += r.w;
    }
    *reinterpret_cast<float4*>(p.C + m * p.ldc + p.coff + n) = make_float4(v[0], v[1], v[2], v[3]);
  }
}
// tiny N (heads): one warp per row, W in its original [N][K] layout
__global__ void k_linear_rowwarp(LinArgs p) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t m = warp; m < p.M; m += nwarps) {
    const float* a = p.A + m * p.lda;
    for (int n = 0; n < p.N; ++n) {
      const float* w = p.W + (int64_t)n * p.K;
      float s = 0.f;
      for (int k = lane; k < p.K; k += 32) s = fmaf(__ldg(a + k), __ldg(w + k), s);
      s = warp_sum(s);
      if (lane == 0) {
        float v = apply_act(s + (p.bias ? p.bias[n] : 0.f), p.act) + p.act_eps;
        if (p.res) v += p.res[m * p.ldr + n];
        if (p.table) v += p.table[(m % p.table_mod) * p.N + n];
        p.C[m * p.ldc + p.coff + n] = v;
      }
    }
  }
}
void op_linear(Ctx& c, const LinArgs& a) {
  if (c.dry) return;
  PAUT_CHECK(a.M > 0, PAUT_ERR_INVALID, "linear: M must be positive");
  if (a.N < 16 || a.N % 4 != 0) {
    PAUT_CHECK(a.W != nullptr, PAUT_ERR_STATE, "linear: small-N path needs the [N][K] weight");
    k_linear_rowwarp<<<grid_for(a.M * 32, 256), 256, 0, c.stream>>>(a);
    c.launched("linear_rowwarp");
    return;
  }
  PAUT_CHECK(a.Wt != nullptr, PAUT_ERR_STATE, "linear: tiled path needs the [K][N] weight");
  PAUT_CHECK(a.K % 4 == 0 && a.lda % 4 == 0 && a.ldc % 4 == 0 && a.coff % 4 == 0, PAUT_ERR_UNSUPPORTED,
             "linear: K, lda, ldc, coff must be multiples of 4");
  dim3 grid((unsigned)((a.M + 63) / 64), (a.N + 63) / 64);
  k_linear_f32<<<grid, 256, 0, c.stream>>>(a);
  c.launched("linear_f32");
}

// ------------------------------------------------------------------------------------------ layernorm
template <int NV>
__global__ void k_layernorm(const float* __restrict__ x, const float* __restrict__ res, const float* __restrict__ g,
                            const float* __restrict__ b, float* __restrict__ out, int64_t M, int D, int act) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t m = warp; m < M; m += nwarps) {
    float v[NV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int d = lane + i * 32;
      float t = 0.f;
      if (d < D) {
        t = x[m * D + d];
        if (res) t += res[m * D + d];
      }
      v[i] = t;
      s += t;
    }
    const float mean = warp_sum(s) / (float)D;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float t = (lane + i * 32 < D) ? v[i] - mean : 0.f;
      q = fmaf(t, t, q);
    }
    const float rstd = rsqrtf(warp_sum(q) / (float)D + 1e-5f);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int d = lane + i * 32;
      if (d < D) out[m * D + d] = apply_act((v[i] - mean) * rstd * g[d] + b[d], act);
    }
  }
}
void op_layernorm(Ctx& c, const float* x, const float* res, const float* gamma, const float* beta, float* out,
                  int64_t M, int D, int act) {
  if (c.dry) return;
  PAUT_CHECK(D <= 1024, PAUT_ERR_UNSUPPORTED, "layernorm: D must be <= 1024");
  const int grid = grid_for(M * 32, 256);
  if (D <= 64) k_layernorm<2><<<grid, 256, 0, c.stream>>>(x, res, gamma, beta, out, M, D, act);
  else if (D <= 128) k_layernorm<4><<<grid, 256, 0, c.stream>>>(x, res, gamma, beta, out, M, D, act);
  else if (D <= 256) k_layernorm<8><<<grid, 256, 0, c.stream>>>(x, res, gamma, beta, out, M, D, act);
  else k_layernorm<32><<<grid, 256, 0, c.stream>>>(x, res, gamma, beta, out, M, D, act);
  c.launched("layernorm");
}

// ------------------------------------------------------------------------------------------ attention
// CTA = (query tile of 16 rows, set).  Loops over heads; K/V of the head live in shared memory.
constexpr int ATT_QT = 16;
__global__ void __launch_bounds__(256) k_attention(const float* __restrict__ q, int ldq, const float* __restrict__ k,
                                                   int ldk, const float* __restrict__ v, int ldv,
                                                   float* __restrict__ out, int ldo, int Nq, int Nk, int H, int hd,
                                                   int kv_shift, float* __restrict__ avgw) {
  extern __shared__ float sm[];
  float* Ks = sm;                               // [Nk][hd+1]
  float* Vs = Ks + (size_t)Nk * (hd + 1);       // [Nk][hd]
  float* Qs = Vs + (size_t)Nk * hd;             // [QT][hd]
  float* Ps = Qs + ATT_QT * hd;                 // [QT][Nk]
  float* Aw = Ps + (size_t)ATT_QT * Nk;         // [QT][Nk] (only when avgw)
  const int64_t b = blockIdx.y;
  const int q0 = blockIdx.x * ATT_QT;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float scale = rsqrtf((float)hd);
  const float invH = 1.f / (float)H;
  if (avgw)
    for (int i = tid; i < ATT_QT * Nk; i += 256) Aw[i] = 0.f;
  for (int h = 0; h < H; ++h) {
    __syncthreads();
    for (int i = tid; i < Nk * hd; i += 256) {
      const int j = i / hd, d = i - j * hd;
      const int js = kv_shift ? min(j + 1, Nk - 1) : j;
      const int64_t row = b * Nk + js;
      Ks[j * (hd + 1) + d] = k[row * ldk + h * hd + d];
      Vs[j * hd + d] = v[row * ldv + h * hd + d];
    }
    for (int i = tid; i < ATT_QT * hd; i += 256) {
      const int r = i / hd, d = i - r * hd;
      Qs[i] = (q0 + r < Nq) ? q[(b * Nq + q0 + r) * (int64_t)ldq + h * hd + d] * scale : 0.f;
    }
    __syncthreads();
    for (int r = warp; r < ATT_QT; r += 8) {
      if (q0 + r >= Nq) continue;                 // warp-uniform
      float* prow = Ps + (size_t)r * Nk;
      float mx = -INFINITY;
      for (int j = lane; j < Nk; j += 32) {
        float s = 0.f;
        for (int d = 0; d < hd; ++d) s = fmaf(Qs[r * hd + d], Ks[j * (hd + 1) + d], s);
        prow[j] = s;
        mx = fmaxf(mx, s);
      }
      mx = warp_max(mx);
      float sum = 0.f;
      for (int j = lane; j < Nk; j += 32) {
        const float e = expf(prow[j] - mx);
        prow[j] = e;
        sum += e;
      }
      const float inv = 1.f / warp_sum(sum);
      __syncwarp();
      for (int j = lane; j < Nk; j += 32) {
        const float pj = prow[j] * inv;
        prow[j] = pj;
        if (avgw) Aw[(size_t)r * Nk + j] += pj * invH;
      }
      __syncwarp();
      for (int d = lane; d < hd; d += 32) {
        float o = 0.f;
        for (int j = 0; j < Nk; ++j) o = fmaf(prow[j], Vs[j * hd + d], o);
        out[(b * Nq + q0 + r) * (int64_t)ldo + h * hd + d] = o;
      }
    }
  }
  if (avgw) {
    __syncthreads();
    for (int i = tid; i < ATT_QT * Nk; i += 256) {
      const int r = i / Nk, j = i - r * Nk;
      if (q0 + r < Nq) avgw[(b * Nq + q0 + r) * (int64_t)Nk + j] = Aw[i];
    }
  }
}
void op_attention(Ctx& c, const float* q, int ldq, const float* k, int ldk, const float* v, int ldv, float* out,
                  int ldo, int64_t B, int Nq, int Nk, int H, int hd, bool kv_shift, float* avgw) {
  if (c.dry) return;
  size_t smem = sizeof(float) * ((size_t)Nk * (hd + 1) + (size_t)Nk * hd + ATT_QT * hd + (size_t)ATT_QT * Nk +
                                 (avgw ? (size_t)ATT_QT * Nk : 0));
  PAUT_CHECK((int)smem <= c.smem_optin, PAUT_ERR_UNSUPPORTED, "attention: set too long for shared memory");
  smem_optin(c, k_attention);
  for (int64_t b0 = 0; b0 < B; b0 += 65535) {
    const int64_t nb = B - b0 < 65535 ? B - b0 : 65535;
    dim3 grid((Nq + ATT_QT - 1) / ATT_QT, (unsigned)nb);
    k_attention<<<grid, 256, smem, c.stream>>>(q + b0 * Nq * ldq, ldq, k + b0 * Nk * ldk, ldk, v + b0 * Nk * ldv, ldv,
                                               out + b0 * Nq * ldo, ldo, Nq, Nk, H, hd, kv_shift ? 1 : 0,
                                               avgw ? avgw + b0 * Nq * Nk : nullptr);
    c.launched("attention");
  }
}

// ------------------------------------------------------------------------------------------ depthwise conv over the set axis
__global__ void k_dwconv_seq(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                             float* __restrict__ out, int64_t B, int N, int D, int k) {
  const int64_t total = B * N * D;
  const int half = k >> 1;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int d = (int)(idx % D);
    const int64_t row = idx / D;
    const int i = (int)(row % N);
    float s = bias[d];
    for (int t = 0; t < k; ++t) {
      const int ii = i + t - half;
      if (ii >= 0 && ii < N) s = fmaf(w[d * k + t], x[(row + (ii - i)) * D + d], s);
    }
    out[idx] = s;
  }
}
void op_dwconv_seq(Ctx& c, const float* x, const float* w, const float* bias, float* out, int64_t B, int N, int D,
                   int k) {
  if (c.dry) return;
  k_dwconv_seq<<<grid_for(B * N * D, 256), 256, 0, c.stream>>>(x, w, bias, out, B, N, D, k);
  c.launched("dwconv_seq");
}

// ------------------------------------------------------------------------------------------ GRU / LSTM
// CTA = (8 sets, direction).  Thread j owns gate row j: W_hh^T column j streamed from L1/L2 each step,
// the 8 hidden states are broadcast from shared memory.
constexpr int RNN_SB = 8;
template <int G>
__global__ void k_rnn_bidir(const float* __restrict__ gi, const float* __restrict__ whh_t,
                            const float* __restrict__ bhh, float* __restrict__ out, int64_t B, int T, int H) {
  extern __shared__ __align__(16) float sm[];
  float* hs = sm;                       // [H][SB] (transposed: see below)
  float* cs = hs + RNN_SB * H;          // [SB][H] (LSTM cell)
  float* gh = cs + RNN_SB * H;          // [SB][G*H]
  const int dir = blockIdx.y;
  const int64_t b0 = (int64_t)blockIdx.x * RNN_SB;
  const int GH = G * H;
  const int j = threadIdx.x;            // blockDim.x == GH
  const float* wt = whh_t + (size_t)dir * H * GH;
  const float bj = bhh[dir * GH + j];
  // hs is stored [H][SB]: the 8 hidden states a thread needs for one k are two 128-bit broadcast loads
  for (int i = j; i < RNN_SB * H; i += GH) { hs[i] = 0.f; cs[i] = 0.f; }
  __syncthreads();
  for (int step = 0; step < T; ++step) {
    const int t = dir == 0 ? step : T - 1 - step;
    float acc[RNN_SB];
#pragma unroll
    for (int s = 0; s < RNN_SB; ++s) acc[s] = bj;
#pragma unroll 4
    for (int k = 0; k < H; ++k) {
      const float w = __ldg(wt + (size_t)k * GH + j);
      const float4 h0 = *reinterpret_cast<const float4*>(hs + k * RNN_SB);
      const float4 h1 = *reinterpret_cast<const float4*>(hs + k * RNN_SB + 4);
      acc[0] = fmaf(w, h0.x, acc[0]); acc[1] = fmaf(w, h0.y, acc[1]); acc[2] = fmaf(w, h0.z, acc[2]); acc[3] = fmaf(w, h0.w, acc[3]);
      acc[4] = fmaf(w, h1.x, acc[4]); acc[5] = fmaf(w, h1.y, acc[5]); acc[6] = fmaf(w, h1.z, acc[6]); acc[7] = fmaf(w, h1.w, acc[7]);
    }
#pragma unroll
    for (int s = 0; s < RNN_SB; ++s) gh[s * GH + j] = acc[s];
    __syncthreads();
    for (int i = j; i < RNN_SB * H; i += GH) {
      const int s = i / H, u = i - s * H;
      const int64_t b = b0 + s;
      if (b < B) {
        const float* g_in = gi + ((b * T + t) * 2 + dir) * (int64_t)GH;
        const float* g_h = gh + s * GH;
        const float hprev = hs[u * RNN_SB + s];
        float hnew;
        if (G == 3) {   // GRU: r | z | n
          const float r = sigmoidf_(g_in[u] + g_h[u]);
          const float z = sigmoidf_(g_in[H + u] + g_h[H + u]);
          const float n = tanhf(g_in[2 * H + u] + r * g_h[2 * H + u]);
          hnew = (1.f - z) * n + z * hprev;
        } else {        // LSTM: i | f | g | o
          const float ig = sigmoidf_(g_in[u] + g_h[u]);
          const float fg = sigmoidf_(g_in[H + u] + g_h[H + u]);
          const float gg = tanhf(g_in[2 * H + u] + g_h[2 * H + u]);
          const float og = sigmoidf_(g_in[3 * H + u] + g_h[3 * H + u]);
          const float cn = fg * cs[i] + ig * gg;
          cs[i] = cn;
          hnew = og * tanhf(cn);
        }
        hs[u * RNN_SB + s] = hnew;
        out[(b * T + t) * (int64_t)(2 * H) + dir * H + u] = hnew;
      }
    }
    __syncthreads();
  }
}
void op_rnn_bidir(Ctx& c, const float* gi, const float* whh_t, const float* bhh, float* out, int64_t B, int T, int H,
                  int G) {
  if (c.dry) return;
  PAUT_CHECK(G * H <= 1024, PAUT_ERR_UNSUPPORTED, "rnn: gates*hidden must be <= 1024");
  const size_t smem = sizeof(float) * (size_t)RNN_SB * H * (2 + G);
  dim3 grid((unsigned)((B + RNN_SB - 1) / RNN_SB), 2);
  if (G == 3) {
    k_rnn_bidir<3><<<grid, G * H, smem, c.stream>>>(gi, whh_t, bhh, out, B, T, H);
  } else {
    PAUT_CHECK(G == 4, PAUT_ERR_INVALID, "rnn: G must be 3 or 4");
    k_rnn_bidir<4><<<grid, G * H, smem, c.stream>>>(gi, whh_t, bhh, out, B, T, H);
  }
  c.launched("rnn_bidir");
}

// ------------------------------------------------------------------------------------------ small row-wise ops
__global__ void k_softmax_seq(const float* __restrict__ x, float* __restrict__ out, int64_t B, int N) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t b = warp; b < B; b += nwarps) {
    float mx = -INFINITY;
    for (int i = lane; i < N; i += 32) mx = fmaxf(mx, x[b * N + i]);
    mx = warp_max(mx);
    float s = 0.f;
    for (int i = lane; i < N; i += 32) s += expf(x[b * N + i] - mx);
    const float inv = 1.f / warp_sum(s);
    for (int i = lane; i < N; i += 32) out[b * N + i] = expf(x[b * N + i] - mx) * inv;
  }
}
void op_softmax_seq(Ctx& c, const float* x, float* out, int64_t B, int N) {
  if (c.dry) return;
  k_softmax_seq<<<grid_for(B * 32, 256), 256, 0, c.stream>>>(x, out, B, N);
  c.launched("softmax_seq");
}

__global__ void k_rowscale_add(const float* __restrict__ a, const float* __restrict__ s, const float* __restrict__ b,
                               float* __restrict__ out, int64_t M, int D) {
  const int64_t total = M * D;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = a[i] * s[i / D];
    out[i] = b ? v + b[i] : v;
  }
}
void op_rowscale_add(Ctx& c, const float* a, const float* s, const float* b, float* out, int64_t M, int D) {
  if (c.dry) return;
  k_rowscale_add<<<grid_for(M * D, 256), 256, 0, c.stream>>>(a, s, b, out, M, D);
  c.launched("rowscale_add");
}

__global__ void k_rowdot(const float* __restrict__ x, const float* __restrict__ q, float* __restrict__ out, int64_t M,
                         int D) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t m = warp; m < M; m += nwarps) {
    float s = 0.f;
    for (int d = lane; d < D; d += 32) s += x[m * D + d] * q[d];     // torch.sum(keys * query): product then sum
    s = warp_sum(s);
    if (lane == 0) out[m] = s;
  }
}
void op_rowdot(Ctx& c, const float* x, const float* q, float* out, int64_t M, int D) {
  if (c.dry) return;
  k_rowdot<<<grid_for(M * 32, 256), 256, 0, c.stream>>>(x, q, out, M, D);
  c.launched("rowdot");
}

__global__ void k_add_table(const float* __restrict__ x, const float* __restrict__ table, int mod,
                            float* __restrict__ out, int64_t M, int D) {
  const int64_t total = M * D;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = i / D;
    out[i] = x[i] + table[(m % mod) * D + (i - m * D)];
  }
}
void op_add_table(Ctx& c, const float* x, const float* table, int mod, float* out, int64_t M, int D) {
  if (c.dry) return;
  k_add_table<<<grid_for(M * D, 256), 256, 0, c.stream>>>(x, table, mod, out, M, D);
  c.launched("add_table");
}

__global__ void k_copy_cols(const float* __restrict__ src, int lds, float* __restrict__ dst, int ldd, int coff,
                            int64_t M, int D) {
  const int64_t total = M * D;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = i / D;
    const int d = (int)(i - m * D);
    dst[m * ldd + coff + d] = src[m * lds + d];
  }
}
void op_copy_cols(Ctx& c, const float* src, int lds, float* dst, int ldd, int coff, int64_t M, int D) {
  if (c.dry) return;
  k_copy_cols<<<grid_for(M * D, 256), 256, 0, c.stream>>>(src, lds, dst, ldd, coff, M, D);
  c.launched("copy_cols");
}

// ------------------------------------------------------------------------------------------ MSC front end
// Persistent CTA over A-scans.  x row + conv1 output (8 x S) + conv2 output (16 x S, only when the background
// extractor needs the neighbours) live in shared memory.  conv2 and the background filter are register-tiled:
// a thread owns 4 consecutive positions x all 16 channels (64 accumulators), so one 128-bit weight load feeds 16
// FMAs and the kernel is bound by the FMA pipe instead of by shared-memory loads.  Accumulation order per output
// (bias, then input channel outer / tap inner; channel sum in channel order) is the reference order.
__global__ void __launch_bounds__(128) k_msc_front(const float* __restrict__ x, int S, int64_t A,
                                                   const float* __restrict__ w1, const float* __restrict__ b1,
                                                   const float* __restrict__ w2, const float* __restrict__ b2,
                                                   const float* __restrict__ wbg, const float* __restrict__ bbg,
                                                   float* __restrict__ f) {
  extern __shared__ __align__(16) float sm[];
  const int S2 = S + 2 + ((4 - (S + 2) % 4) % 4);      // padded row length (multiple of 4 floats)
  const int SP = S + 12;                               // conv2 rows: zero halo of 5 on the left, 7 on the right
  float* xs = sm;                      // [S2]      zero halo of 1
  float* a1 = xs + S2;                 // [8][S2]   zero halo of 1
  float* a2 = a1 + 8 * S2;             // [16][SP]  (only with background)
  __shared__ __align__(16) float W1[8 * 3], B1[8], W2t[24 * 16], B2[16], WB[16 * 12], BB[16];
  const int tid = threadIdx.x;
  for (int i = tid; i < 24; i += 128) W1[i] = w1[i];
  for (int i = tid; i < 8; i += 128) B1[i] = b1[i];
  for (int i = tid; i < 384; i += 128) {               // w2 [co][ci][t] -> W2t [ci*3+t][co]
    const int co = i / 24, r = i - co * 24;
    W2t[r * 16 + co] = w2[i];
  }
  for (int i = tid; i < 16; i += 128) B2[i] = b2[i];
  if (wbg) {
    for (int i = tid; i < 16 * 12; i += 128) { const int co = i / 12, t = i - co * 12; WB[i] = t < 11 ? wbg[co * 11 + t] : 0.f; }
    for (int i = tid; i < 16; i += 128) BB[i] = bbg[i];
    for (int i = tid; i < 16 * SP; i += 128) a2[i] = 0.f;           // halos stay zero: only [5, 5+S) is rewritten
  }
  for (int i = tid; i < 9 * S2; i += 128) sm[i] = 0.f;
  const int l0 = 4 * tid;
  for (int64_t a = blockIdx.x; a < A; a += gridDim.x) {
    __syncthreads();
    for (int i = tid; i < S; i += 128) xs[i + 1] = x[a * S + i];
    __syncthreads();
    for (int i = tid; i < 8 * S; i += 128) {
      const int ch = i / S, l = i - ch * S;
      float v = B1[ch];
      v = fmaf(W1[ch * 3 + 0], xs[l], v);
      v = fmaf(W1[ch * 3 + 1], xs[l + 1], v);
      v = fmaf(W1[ch * 3 + 2], xs[l + 2], v);
      a1[ch * S2 + l + 1] = fmaxf(v, 0.f);
    }
    __syncthreads();
    float acc[4][16];
    if (l0 < S) {
#pragma unroll
      for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int co = 0; co < 16; ++co) acc[p][co] = B2[co];
#pragma unroll
      for (int ci = 0; ci < 8; ++ci) {
        float av[6];                                    // positions l0-1 .. l0+4
        const float4 v4 = *reinterpret_cast<const float4*>(a1 + ci * S2 + l0);
        const float2 v2 = *reinterpret_cast<const float2*>(a1 + ci * S2 + l0 + 4);
        av[0] = v4.x; av[1] = v4.y; av[2] = v4.z; av[3] = v4.w; av[4] = v2.x; av[5] = v2.y;
#pragma unroll
        for (int t = 0; t < 3; ++t) {
          float w[16];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 w4 = *reinterpret_cast<const float4*>(W2t + (ci * 3 + t) * 16 + 4 * j);
            w[4 * j] = w4.x; w[4 * j + 1] = w4.y; w[4 * j + 2] = w4.z; w[4 * j + 3] = w4.w;
          }
#pragma unroll
          for (int p = 0; p < 4; ++p)
#pragma unroll
            for (int co = 0; co < 16; ++co) acc[p][co] = fmaf(w[co], av[p + t], acc[p][co]);
        }
      }
    }
    if (!wbg) {
      if (l0 < S) {
        float s[4];
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          s[p] = 0.f;
#pragma unroll
          for (int co = 0; co < 16; ++co) s[p] += fmaxf(acc[p][co], 0.f);
        }
        *reinterpret_cast<float4*>(f + a * S + l0) =
            make_float4(s[0] * (1.f / 16.f), s[1] * (1.f / 16.f), s[2] * (1.f / 16.f), s[3] * (1.f / 16.f));
      }
    } else {
      if (l0 < S) {
#pragma unroll
        for (int co = 0; co < 16; ++co)
#pragma unroll
          for (int p = 0; p < 4; ++p) a2[co * SP + 5 + l0 + p] = fmaxf(acc[p][co], 0.f);
      }
      __syncthreads();
      if (l0 < S) {
        float s[4] = {0.f, 0.f, 0.f, 0.f};
        for (int co = 0; co < 16; ++co) {
          float ar[16];                                 // ar[j] = position l0 + j - 5
          const float* base = a2 + co * SP + l0;        // (5 + l0 - 5); l0 % 4 == 0 and SP % 4 == 0: 16-byte aligned
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 v4 = *reinterpret_cast<const float4*>(base + 4 * j);
            ar[4 * j] = v4.x; ar[4 * j + 1] = v4.y; ar[4 * j + 2] = v4.z; ar[4 * j + 3] = v4.w;
          }
          float wb[12];
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            const float4 w4 = *reinterpret_cast<const float4*>(WB + co * 12 + 4 * j);
            wb[4 * j] = w4.x; wb[4 * j + 1] = w4.y; wb[4 * j + 2] = w4.z; wb[4 * j + 3] = w4.w;
          }
#pragma unroll
          for (int p = 0; p < 4; ++p) {
            float bg = BB[co];
#pragma unroll
            for (int t = 0; t < 11; ++t) bg = fmaf(wb[t], ar[p + t], bg);
            s[p] += ar[p + 5] - bg;
          }
        }
        *reinterpret_cast<float4*>(f + a * S + l0) =
            make_float4(s[0] * (1.f / 16.f), s[1] * (1.f / 16.f), s[2] * (1.f / 16.f), s[3] * (1.f / 16.f));
      }
    }
  }
}
void op_msc_front(Ctx& c, const float* x, int64_t A, int S, const float* w1, const float* b1, const float* w2,
                  const float* b2, const float* wbg, const float* bbg, float* f) {
  if (c.dry) return;
  PAUT_CHECK(S % 4 == 0 && S <= 512, PAUT_ERR_UNSUPPORTED, "msc front: signal length must be a multiple of 4, at most 512");
  const int S2 = S + 2 + ((4 - (S + 2) % 4) % 4);
  const size_t smem = sizeof(float) * ((size_t)S2 * 9 + (wbg ? (size_t)16 * (S + 12) : 0));
  PAUT_CHECK(smem <= 48 * 1024, PAUT_ERR_UNSUPPORTED, "msc front: signal too long");
  int64_t grid = (int64_t)c.num_sms * 6;
  if (grid > A) grid = A;
  k_msc_front<<<(unsigned)grid, 128, smem, c.stream>>>(x, S, A, w1, b1, w2, b2, wbg, bbg, f);
  c.launched("msc_front");
}

__global__ void k_msc_head(const float* __restrict__ o, int64_t M, float* prob, float* start, float* end) {
  for (int64_t m = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; m < M; m += (int64_t)gridDim.x * blockDim.x) {
    if (prob) prob[m] = sigmoidf_(o[m * 3 + 0]);
    if (start) start[m] = tanhf(o[m * 3 + 1]) * 0.5f + 0.5f;
    if (end) end[m] = tanhf(o[m * 3 + 2]) * 0.5f + 0.5f;
  }
}
void op_msc_head(Ctx& c, const float* o, int64_t M, float* prob, float* start, float* end) {
  if (c.dry) return;
  k_msc_head<<<grid_for(M, 256), 256, 0, c.stream>>>(o, M, prob, start, end);
  c.launched("msc_head");
}

__global__ void k_add_anomaly(float* logits, const float* __restrict__ anomaly, int64_t M, int C) {
  const int64_t total = M * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cc = (int)(i % C);
    if (cc >= 1) logits[i] += anomaly[i / C];
  }
}
void op_add_anomaly(Ctx& c, float* logits, const float* anomaly, int64_t M, int C) {
  if (c.dry) return;
  if (C <= 1) return;
  k_add_anomaly<<<grid_for(M * C, 256), 256, 0, c.stream>>>(logits, anomaly, M, C);
  c.launched("add_anomaly");
}

__global__ void k_two_stage_final(const float* __restrict__ logits, float* probs, float* pos, int64_t M) {
  for (int64_t m = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; m < M; m += (int64_t)gridDim.x * blockDim.x) {
    const float l0 = logits[m * 2], l1 = logits[m * 2 + 1];
    const float mx = fmaxf(l0, l1);
    const float e0 = expf(l0 - mx), e1 = expf(l1 - mx);
    const float sum = e0 + e1;
    const float p0 = e0 / sum, p1 = e1 / sum;
    if (probs) { probs[m * 2] = p0; probs[m * 2 + 1] = p1; }
    if (pos) { pos[m * 2] *= p1; pos[m * 2 + 1] *= p1; }
  }
}
void op_two_stage_final(Ctx& c, const float* logits, float* probs, float* pos, int64_t M) {
  if (c.dry) return;
  k_two_stage_final<<<grid_for(M, 256), 256, 0, c.stream>>>(logits, probs, pos, M);
  c.launched("two_stage_final");
}

}  // namespace paut
