// Operators of the SURVEY section-8 "next" rows (f2 improved / hybrid / complex models, f3 legacy-MSC
// difference matrix, f4 detection-level metrics).  All of them are HBM-bound elementwise / reduction /
// integer work: coalesced loads, shared-memory staging, warp ballots; no tensor cores.
#include "common.cuh"

namespace paut {

namespace {
inline unsigned grid_cap(int64_t n, int threads, int64_t cap = 148 * 16) {
  int64_t b = (n + threads - 1) / threads;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (unsigned)b;
}
__device__ __forceinline__ float ldf(const float* p) { return *p; }
__device__ __forceinline__ float ldf(const __nv_bfloat16* p) { return __bfloat162float(*p); }
// sum of the 4 fp32 / 8 bf16 values of one 128-bit load
__device__ __forceinline__ float sum128(const uint4* p, const float*) {
  const uint4 u = __ldg(p);
  return (__uint_as_float(u.x) + __uint_as_float(u.y)) + (__uint_as_float(u.z) + __uint_as_float(u.w));
}
__device__ __forceinline__ float sum128(const uint4* p, const __nv_bfloat16*) {
  const uint4 u = __ldg(p);
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) s += __uint_as_float(w[i] << 16) + __uint_as_float(w[i] & 0xffff0000u);
  return s;
}
}  // namespace

// ------------------------------------------------------------------------------------------ f2: front ends
// ImprovedMultiSignalClassifier (improved_model.py:126-133): bg = depthwise Conv1d(C, C, k, pad k/2, groups C);
// x = x - bg; f = mean over channels.  in [A, L, C] channels-last fp32, w [C][k], bias [C] -> f [A, L].
// Folded: f[l] = (1/C) sum_c sum_o w''_c[o] x_c[l + o - 7] - mean(bias) with w'' = delta - w centred in a 15-wide
// window.  One CTA per A-scan: the tile is staged TRANSPOSED in shared memory ([C][L + halo], zero halos, so the
// taps need no bounds checks); a thread owns 4 consecutive positions of one channel group, reads its 20-sample
// window as five 128-bit loads and does 60 FMAs per channel (the first version did one LDS per FMA and was
// shared-memory bound); the channel groups' partial sums meet in shared memory.
constexpr int BG_HALO = 8;         // left halo (>= k/2, multiple of 4)
constexpr int BG_MAXCG = 8;
template <typename T>
__global__ void __launch_bounds__(256) k_bgsub_chanmean(const T* __restrict__ in, int L, int C, int k, int CG, int Lrow,
                                                         int H0, const float* __restrict__ w,
                                                         const float* __restrict__ bias, float* __restrict__ f) {
  extern __shared__ __align__(16) float sm[];
  const int Lp = L + 20;               // 8 left + 12 right zero samples; Lp % 32 = 4 * odd for L % 32 == 0
  float* tile = sm;                    // [C][Lp]
  float* ws = tile + (size_t)C * Lp;   // [C][16]: centred folded taps, ws[c][15] = 0
  float* part = ws + C * 16;           // [CG][L]
  __shared__ float meanb;
  const int64_t a = blockIdx.x;
  const T* src = in + ((int64_t)H0 + a * Lrow) * C;     // rows row(a, l) = H0 + a*Lrow + l (dense: H0 = 0, Lrow = L)
  for (int i = threadIdx.x; i < C * 20; i += blockDim.x) {
    const int c = i / 20, j = i % 20;
    tile[c * Lp + (j < BG_HALO ? j : L + j)] = 0.f;
  }
  constexpr int EPV = 16 / (int)sizeof(T);                 // 128-bit global loads: EPV channels of one row each
  if (C % EPV == 0) {
    const int vpr = C / EPV;
    const uint4* src4 = reinterpret_cast<const uint4*>(src);
    for (int i = threadIdx.x; i < L * vpr; i += blockDim.x) {
      const int l = i / vpr, c0 = (i - l * vpr) * EPV;
      const uint4 u = __ldg(src4 + i);
      const uint32_t wv[4] = {u.x, u.y, u.z, u.w};
      float* dst = tile + c0 * Lp + BG_HALO + l;
      if (sizeof(T) == 4) {
#pragma unroll
        for (int j = 0; j < 4; ++j) dst[j * Lp] = __uint_as_float(wv[j]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          dst[(2 * j) * Lp] = __uint_as_float(wv[j] << 16);
          dst[(2 * j + 1) * Lp] = __uint_as_float(wv[j] & 0xffff0000u);
        }
      }
    }
  } else {
    for (int i = threadIdx.x; i < L * C; i += blockDim.x) tile[(i % C) * Lp + BG_HALO + i / C] = ldf(src + i);
  }
  const int half = k >> 1;
  for (int i = threadIdx.x; i < C * 16; i += blockDim.x) {
    const int c = i >> 4, o = i & 15, t = o - 7 + half;          // o = t - half + 7
    float v = 0.f;
    if (o < 15 && t >= 0 && t < k) v = (t == half ? 1.f : 0.f) - w[c * k + t];
    ws[i] = v;
  }
  if (threadIdx.x < 32) {
    float b = 0.f;
    for (int c = threadIdx.x; c < C; c += 32) b += bias[c];
    for (int o = 16; o > 0; o >>= 1) b += __shfl_xor_sync(0xffffffffu, b, o);
    if (threadIdx.x == 0) meanb = b / (float)C;
  }
  __syncthreads();
  const int PG = L >> 2;
  for (int item = threadIdx.x; item < PG * CG; item += blockDim.x) {
    const int pg = item % PG, cg = item / PG;
    const int l0 = pg * 4;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int c = cg; c < C; c += CG) {
      float v[20], wv[16];
      const float4* tp = reinterpret_cast<const float4*>(tile + c * Lp + l0);      // samples l0 - 8 .. l0 + 11
#pragma unroll
      for (int i = 0; i < 5; ++i) {
        const float4 t4 = tp[i];
        v[4 * i] = t4.x; v[4 * i + 1] = t4.y; v[4 * i + 2] = t4.z; v[4 * i + 3] = t4.w;
      }
      const float4* wp = reinterpret_cast<const float4*>(ws + c * 16);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 t4 = wp[i];
        wv[4 * i] = t4.x; wv[4 * i + 1] = t4.y; wv[4 * i + 2] = t4.z; wv[4 * i + 3] = t4.w;
      }
#pragma unroll
      for (int o = 0; o < 15; ++o) {
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[j] = fmaf(wv[o], v[j + o + 1], acc[j]);   // sample l0 + j + o - 7
      }
    }
    *reinterpret_cast<float4*>(part + cg * L + l0) = make_float4(acc[0], acc[1], acc[2], acc[3]);
  }
  __syncthreads();
  const float inv = 1.f / (float)C;
  for (int l = threadIdx.x; l < L; l += blockDim.x) {
    float s = 0.f;
    for (int cg = 0; cg < CG; ++cg) s += part[cg * L + l];
    f[a * L + l] = s * inv - meanb;
  }
}

void op_bgsub_chanmean(Ctx& c, const void* in, int in_dtype, int64_t A, int L, int C, int Lrow, int H0, int k,
                       const float* w, const float* bias, float* f) {
  if (c.dry) return;
  PAUT_CHECK(L % 4 == 0 && k % 2 == 1 && k <= 15, PAUT_ERR_UNSUPPORTED,
             "bgsub_chanmean: L must be a multiple of 4 and the kernel size odd and <= 15");
  int CG = 256 / (L / 4);
  CG = CG < 1 ? 1 : (CG > BG_MAXCG ? BG_MAXCG : CG);
  if (CG > C) CG = C;
  const size_t smem = ((size_t)C * (L + 20) + (size_t)C * 16 + (size_t)CG * L) * sizeof(float);
  PAUT_CHECK(smem <= (size_t)c.smem_optin, PAUT_ERR_UNSUPPORTED, "bgsub_chanmean: A-scan tile exceeds shared memory");
  if (smem > 48 * 1024) {
    smem_optin(c, k_bgsub_chanmean<float>);
    smem_optin(c, k_bgsub_chanmean<__nv_bfloat16>);
  }
  if (in_dtype == PAUT_F32)
    k_bgsub_chanmean<float><<<(unsigned)A, 256, smem, c.stream>>>(static_cast<const float*>(in), L, C, k, CG, Lrow, H0, w,
                                                                   bias, f);
  else
    k_bgsub_chanmean<__nv_bfloat16><<<(unsigned)A, 256, smem, c.stream>>>(static_cast<const __nv_bfloat16*>(in), L, C, k,
                                                                           CG, Lrow, H0, w, bias, f);
  c.launched("bgsub_chanmean");
}

// HybridBinaryModel (hybrid_binary.py:141-150) and ComplexDetectionModel (complex_detection_model.py:71-75):
// pooling / interpolation along L and the mean over channels are all linear, so the channel mean is taken first
// (one row of C values -> one scalar, coalesced) and the 1-D resampling runs on the resulting length-L signal.
//   mode 0: AvgPool1d(kernel = stride = pk) -> F.interpolate(size = P, mode='linear', align_corners=False)
//   mode 1: AdaptiveAvgPool1d(P)
// Input rows: row(a, l) = H0 + a*Lp + l of a [rows, C] buffer (dense fp32: H0 = 0, Lp = L; bf16 flat rows of the
// tcgen05 conv: H0 = halo, Lp = L + halo).  One CTA per A-scan.
template <typename T>
__global__ void __launch_bounds__(256) k_chanmean_resample(const T* __restrict__ in, int L, int C, int Lp, int H0,
                                                            int mode, int pk, int P, float* __restrict__ out) {
  extern __shared__ float sm[];
  float* m = sm;           // [L] channel means
  float* p = sm + L;       // [L / pk] pooled (mode 0)
  const int64_t a = blockIdx.x;
  const T* src = in + ((int64_t)H0 + a * Lp) * C;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const float inv = 1.f / (float)C;
  constexpr int EPV = 16 / (int)sizeof(T);            // elements per 128-bit load
  const int V = C / EPV;                              // 128-bit loads per row
  if (C % EPV == 0 && V <= 32 && (V & (V - 1)) == 0) {
    // V lanes share a row (one 128-bit load each), 32 / V rows per warp instruction, log2(V) shuffles per row group
    const int rpw = 32 / V, sub = lane / V, part = lane % V;
    for (int l0 = warp * rpw; l0 < L; l0 += nw * rpw) {
      const int l = l0 + sub;
      float s = 0.f;
      if (l < L) s = sum128(reinterpret_cast<const uint4*>(src + (int64_t)l * C) + part, static_cast<const T*>(nullptr));
      for (int o = V >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (part == 0 && l < L) m[l] = s * inv;
    }
  } else {
    for (int l = warp; l < L; l += nw) {
      float s = 0.f;
      for (int ch = lane; ch < C; ch += 32) s += ldf(src + (int64_t)l * C + ch);
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) m[l] = s * inv;
    }
  }
  __syncthreads();
  if (mode == 0) {
    const int Lq = L / pk;
    const float ipk = 1.f / (float)pk;
    for (int j = threadIdx.x; j < Lq; j += blockDim.x) {
      float s = 0.f;
      for (int t = 0; t < pk; ++t) s += m[j * pk + t];
      p[j] = s * ipk;
    }
    __syncthreads();
    const float scale = (float)Lq / (float)P;                 // area_pixel_compute_scale, align_corners=False
    for (int j = threadIdx.x; j < P; j += blockDim.x) {
      float srcx = scale * ((float)j + 0.5f) - 0.5f;
      if (srcx < 0.f) srcx = 0.f;
      const int i0 = (int)srcx;
      const int i1 = i0 + (i0 < Lq - 1 ? 1 : 0);
      const float l1 = srcx - (float)i0, l0 = 1.f - l1;
      out[a * P + j] = l0 * p[i0] + l1 * p[i1];
    }
  } else {
    for (int j = threadIdx.x; j < P; j += blockDim.x) {
      const int s0 = (int)(((int64_t)j * L) / P);
      const int e0 = (int)((((int64_t)j + 1) * L + P - 1) / P);
      float s = 0.f;
      for (int l = s0; l < e0; ++l) s += m[l];
      out[a * P + j] = s / (float)(e0 - s0);
    }
  }
}

void op_chanmean_resample(Ctx& c, const void* in, int in_dtype, int64_t A, int L, int C, int Lp, int H0, int mode,
                          int pk, int P, float* out) {
  if (c.dry) return;
  PAUT_CHECK(pk >= 1 && L / pk >= 1 && P >= 1, PAUT_ERR_INVALID, "chanmean_resample: bad pooling geometry");
  const size_t smem = ((size_t)L + (size_t)L / pk + 1) * sizeof(float);
  PAUT_CHECK(smem <= 48 * 1024, PAUT_ERR_UNSUPPORTED, "chanmean_resample: signal too long");
  if (in_dtype == PAUT_F32)
    k_chanmean_resample<float><<<(unsigned)A, 256, smem, c.stream>>>(static_cast<const float*>(in), L, C, Lp, H0, mode,
                                                                      pk, P, out);
  else
    k_chanmean_resample<__nv_bfloat16><<<(unsigned)A, 256, smem, c.stream>>>(static_cast<const __nv_bfloat16*>(in), L,
                                                                              C, Lp, H0, mode, pk, P, out);
  c.launched("chanmean_resample");
}

// hybrid_binary.py:147-149: seq_concat = cat([seq, seq - seq.mean(dim=1)], -1).  seq [B, N, D] -> out [B, N, 2D].
// One CTA per set, one thread per feature column (coalesced rows), fixed summation order over the set axis.
__global__ void k_seqmean_concat(const float* __restrict__ seq, int N, int D, float* __restrict__ out) {
  const int64_t b = blockIdx.x;
  const float* s = seq + b * (int64_t)N * D;
  float* o = out + b * (int64_t)N * 2 * D;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float sum = 0.f;
    for (int i = 0; i < N; ++i) sum += s[(int64_t)i * D + d];
    const float mean = sum / (float)N;
    for (int i = 0; i < N; ++i) {
      const float v = s[(int64_t)i * D + d];
      o[(int64_t)i * 2 * D + d] = v;
      o[(int64_t)i * 2 * D + D + d] = v - mean;
    }
  }
}
void op_seqmean_concat(Ctx& c, const float* seq, int64_t B, int N, int D, float* out) {
  if (c.dry) return;
  k_seqmean_concat<<<(unsigned)B, 128, 0, c.stream>>>(seq, N, D, out);
  c.launched("seqmean_concat");
}

// improved_model.py:147-156: prob = sigmoid(o0); start = clamp(o1, 0, 1); end = clamp(o2, 0, 1).  o [M, 3].
__global__ void k_improved_head(const float* __restrict__ o, int64_t M, float* prob, float* start, float* end) {
  for (int64_t m = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; m < M; m += (int64_t)gridDim.x * blockDim.x) {
    if (prob) prob[m] = 1.f / (1.f + expf(-o[m * 3]));
    if (start) start[m] = fminf(fmaxf(o[m * 3 + 1], 0.f), 1.f);
    if (end) end[m] = fminf(fmaxf(o[m * 3 + 2], 0.f), 1.f);
  }
}
void op_improved_head(Ctx& c, const float* o, int64_t M, float* prob, float* start, float* end) {
  if (c.dry) return;
  k_improved_head<<<grid_cap(M, 256), 256, 0, c.stream>>>(o, M, prob, start, end);
  c.launched("improved_head");
}

// ------------------------------------------------------------------------------------------ f1: all-zero-run drop
// dataset_preparation.py:205 skips a run whose signals are all zero (np.all(signals == 0)).  flags[g] = 1 iff
// any sample of volume[g] is non-zero (-0.0 counts as zero, NaN as non-zero, like the numpy comparison).
// One CTA per group, 128-bit loads, block-uniform early exit (__syncthreads_or).
template <typename T>
__global__ void __launch_bounds__(256) k_group_nonzero(const T* __restrict__ vol, int64_t per_group,
                                                        int32_t* __restrict__ flags) {
  int found = 0;
  constexpr int EPV = 16 / (int)sizeof(T);
  const uint4* src = reinterpret_cast<const uint4*>(vol + blockIdx.x * per_group);
  const int64_t nvec = per_group / EPV;
  constexpr uint32_t M = sizeof(T) == 4 ? 0x7fffffffu : 0x7fff7fffu;      // drop the sign bit(s): -0.0 == 0
  for (int64_t i0 = 0; i0 < nvec; i0 += 256 * 8) {
    uint32_t acc = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int64_t i = i0 + j * 256 + threadIdx.x;
      if (i < nvec) {
        const uint4 u = __ldg(src + i);
        acc |= (u.x & M) | (u.y & M) | (u.z & M) | (u.w & M);
      }
    }
    found = __syncthreads_or(acc != 0);               // block-uniform: every thread leaves the loop together
    if (found) break;
  }
  if (threadIdx.x == 0) flags[blockIdx.x] = found;
}
void op_group_nonzero(Ctx& c, const void* vol, int dtype, int64_t G, int64_t per_group, int32_t* flags) {
  if (c.dry) return;
  const int epv = dtype == PAUT_F32 ? 4 : 8;
  PAUT_CHECK(per_group % epv == 0, PAUT_ERR_UNSUPPORTED, "group_nonzero: group size must be a multiple of 16 bytes");
  if (dtype == PAUT_F32)
    k_group_nonzero<float><<<(unsigned)G, 256, 0, c.stream>>>(static_cast<const float*>(vol), per_group, flags);
  else
    k_group_nonzero<__nv_bfloat16><<<(unsigned)G, 256, 0, c.stream>>>(static_cast<const __nv_bfloat16*>(vol), per_group, flags);
  c.launched("group_nonzero");
}

// ------------------------------------------------------------------------------------------ f3: difference matrix
// teststtt.py:54-69.  Per set: reference signal = mean of the A-scans with pred < thr (fp64, ascending order, like
// np.mean(axis=0) over the float64 rows), difference row = |signal - reference| where pred >= thr, zeros elsewhere.
// Thread = one sample column: the loop over the set axis reads coalesced rows.
template <typename T>
__global__ void k_reference_signal(const T* __restrict__ x, const float* __restrict__ prob, int N, int S, double thr,
                                   double* __restrict__ ref64, float* __restrict__ ref, int32_t* __restrict__ healthy) {
  const int64_t b = blockIdx.y;
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  const T* xs = x + b * (int64_t)N * S;
  const float* pb = prob + b * N;
  double sum = 0.0;
  int cnt = 0;
  for (int i = 0; i < N; ++i) {
    if ((double)pb[i] < thr) {
      sum += (double)ldf(xs + (int64_t)i * S + s);
      ++cnt;
    }
  }
  const double mean = cnt > 0 ? sum / (double)cnt : 0.0;
  ref64[b * S + s] = mean;
  if (ref) ref[b * S + s] = (float)mean;
  if (s == 0 && healthy) healthy[b] = cnt;
}
template <typename T>
__global__ void k_difference_matrix(const T* __restrict__ x, const float* __restrict__ prob,
                                    const double* __restrict__ ref64, const int32_t* __restrict__ healthy, int64_t B,
                                    int N, int S, double thr, float* __restrict__ diff) {
  const int64_t total = B * N * S;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int s = (int)(idx % S);
    const int64_t row = idx / S;
    const int64_t b = row / N;
    float d = 0.f;
    if (healthy[b] > 0 && (double)prob[row] >= thr) d = (float)fabs((double)ldf(x + idx) - ref64[b * S + s]);
    diff[idx] = d;
  }
}

void op_difference_matrix(Ctx& c, const void* x, int x_dtype, const float* prob, int64_t B, int N, int S, double thr,
                          float* ref, float* diff, int32_t* healthy) {
  double* ref64 = static_cast<double*>(c.alloc(sizeof(double) * (size_t)B * S));
  int32_t* hc = healthy ? healthy : static_cast<int32_t*>(c.alloc(sizeof(int32_t) * (size_t)B));
  if (c.dry) return;
  dim3 grid((S + 127) / 128, (unsigned)B);
  if (x_dtype == PAUT_F32)
    k_reference_signal<float><<<grid, 128, 0, c.stream>>>(static_cast<const float*>(x), prob, N, S, thr, ref64, ref, hc);
  else
    k_reference_signal<__nv_bfloat16><<<grid, 128, 0, c.stream>>>(static_cast<const __nv_bfloat16*>(x), prob, N, S, thr,
                                                                  ref64, ref, hc);
  c.launched("reference_signal");
  if (!diff) return;
  const unsigned g = grid_cap(B * N * S, 256, 148 * 32);
  if (x_dtype == PAUT_F32)
    k_difference_matrix<float><<<g, 256, 0, c.stream>>>(static_cast<const float*>(x), prob, ref64, hc, B, N, S, thr, diff);
  else
    k_difference_matrix<__nv_bfloat16><<<g, 256, 0, c.stream>>>(static_cast<const __nv_bfloat16*>(x), prob, ref64, hc, B,
                                                                N, S, thr, diff);
  c.launched("difference_matrix");
}

// ------------------------------------------------------------------------------------------ f4: detection metrics
// 1-D IoU in the reference's arithmetic: both positions are numpy float32 scalars (predict() returns float32
// arrays, the targets are tensor.cpu().numpy()), so every operation below is an fp32 operation.
__device__ __forceinline__ bool iou_f32(float ps, float pe, float ts, float te, float* iou) {
  const float inter = fmaxf(0.f, __fsub_rn(fminf(pe, te), fmaxf(ps, ts)));
  const float uni = __fsub_rn(fmaxf(pe, te), fminf(ps, ts));
  if (!(uni > 0.f)) return false;
  *iou = __fdiv_rn(inter, uni);
  return true;
}

struct SetMetrics {
  int32_t tp, fp, fn, pad;
  double sum_iou, sum_err;
};

// One warp per set.  det is in (set, position) order (paut_postprocess); the set's range is found by binary search.
//   rule 0 (two_stage_train.py:284-375): a prediction matches the not-yet-matched target at the SAME position with
//          the best IoU (at most one target per position in the dense layout), TP iff that IoU > thr.
//   rule 1 (train.py:279-361): greedy first match in target order among not-yet-matched targets of the same class
//          with IoU > 0.5 (the threshold is hard-coded there; thr is passed by the caller).
__global__ void __launch_bounds__(128) k_metrics_match(int rule, const paut_detection* __restrict__ det,
                                                        const int32_t* __restrict__ count_dev, int64_t B, int N,
                                                        const int32_t* __restrict__ tlabel,
                                                        const float* __restrict__ tpos, float thr,
                                                        SetMetrics* __restrict__ per_set) {
  const int lane = threadIdx.x & 31;
  const int64_t b = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (b >= B) return;
  const int total = *count_dev;
  // lower_bound of set_index >= b and >= b+1 (all lanes compute the same thing)
  auto lower = [&](int64_t key) {
    int lo = 0, hi = total;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (det[mid].set_index < key) lo = mid + 1; else hi = mid;
    }
    return lo;
  };
  const int p0 = lower(b), p1 = lower(b + 1);
  const int32_t* lab = tlabel + b * N;
  const float* tp_ = tpos + b * (int64_t)N * 2;
  int ntargets = 0;
  for (int i = lane; i < N; i += 32) ntargets += lab[i] > 0 ? 1 : 0;
  for (int o = 16; o > 0; o >>= 1) ntargets += __shfl_xor_sync(0xffffffffu, ntargets, o);
  int tp = 0, fp = 0, matched = 0;
  double sum_iou = 0.0, sum_err = 0.0;
  if (rule == 0) {
    // predictions are independent (unique positions): lanes take predictions, warp-reduce at the end;
    // the fp64 sums are accumulated in prediction order by lane 0 through a ballot loop to keep them deterministic
    for (int q0 = p0; q0 < p1; q0 += 32) {
      const int q = q0 + lane;
      bool hit = false;
      float iou = 0.f, err = 0.f;
      if (q < p1) {
        const paut_detection r = det[q];
        const int i = r.position;
        if (i >= 0 && i < N && lab[i] > 0) {
          const float ts = tp_[i * 2], te = tp_[i * 2 + 1];
          float v;
          if (iou_f32(r.start, r.end, ts, te, &v) && v > 0.f && v > thr) {
            hit = true;
            iou = v;
            err = __fmul_rn(__fadd_rn(fabsf(__fsub_rn(r.start, ts)), fabsf(__fsub_rn(r.end, te))), 0.5f);
          }
        }
      }
      const unsigned mask = __ballot_sync(0xffffffffu, hit);
      const unsigned valid = __ballot_sync(0xffffffffu, q < p1);
      tp += __popc(mask);
      fp += __popc(valid & ~mask);
      for (unsigned mm = mask; mm; mm &= mm - 1) {
        const int src = __ffs(mm) - 1;
        sum_iou += (double)__shfl_sync(0xffffffffu, iou, src);
        sum_err += (double)__shfl_sync(0xffffffffu, err, src);
      }
    }
    matched = tp;
  } else {
    // sequential greedy over predictions; lane l owns targets l, l+32, ... (bit c of `done` = target c*32+l matched)
    unsigned done = 0;
    const int chunks = (N + 31) >> 5;
    for (int q = p0; q < p1; ++q) {
      const paut_detection r = det[q];
      bool found = false;
      for (int ch = 0; ch < chunks && !found; ++ch) {
        const int t = ch * 32 + lane;
        bool ok = false;
        float v = 0.f;
        if (t < N && lab[t] > 0 && !((done >> ch) & 1u) && lab[t] == r.cls)
          ok = iou_f32(r.start, r.end, tp_[t * 2], tp_[t * 2 + 1], &v) && v > thr;
        const unsigned mask = __ballot_sync(0xffffffffu, ok);
        if (mask) {
          const int src = __ffs(mask) - 1;
          if (lane == src) done |= 1u << ch;
          sum_iou += (double)__shfl_sync(0xffffffffu, v, src);
          found = true;
        }
      }
      if (found) { ++tp; ++matched; } else { ++fp; }
    }
  }
  if (lane == 0) {
    SetMetrics m;
    m.tp = tp; m.fp = fp; m.fn = ntargets - matched; m.pad = 0;
    m.sum_iou = sum_iou; m.sum_err = sum_err;
    per_set[b] = m;
  }
}

// fixed-order reduction of the per-set partials (single CTA): integer counts exact, fp64 sums reproducible
__global__ void __launch_bounds__(1024) k_metrics_reduce(const SetMetrics* __restrict__ per_set, int64_t B,
                                                          paut_metrics* __restrict__ out) {
  __shared__ long long s_tp[1024], s_fp[1024], s_fn[1024];
  __shared__ double s_iou[1024], s_err[1024];
  long long tp = 0, fp = 0, fn = 0;
  double si = 0.0, se = 0.0;
  for (int64_t b = threadIdx.x; b < B; b += 1024) {
    const SetMetrics m = per_set[b];
    tp += m.tp; fp += m.fp; fn += m.fn; si += m.sum_iou; se += m.sum_err;
  }
  s_tp[threadIdx.x] = tp; s_fp[threadIdx.x] = fp; s_fn[threadIdx.x] = fn; s_iou[threadIdx.x] = si; s_err[threadIdx.x] = se;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      s_tp[threadIdx.x] += s_tp[threadIdx.x + o]; s_fp[threadIdx.x] += s_fp[threadIdx.x + o];
      s_fn[threadIdx.x] += s_fn[threadIdx.x + o]; s_iou[threadIdx.x] += s_iou[threadIdx.x + o];
      s_err[threadIdx.x] += s_err[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    out->tp = s_tp[0]; out->fp = s_fp[0]; out->fn = s_fn[0]; out->tn = 0;
    out->sum_iou = s_iou[0]; out->sum_position_error = s_err[0];
  }
}

void op_metrics_match(Ctx& c, int rule, const paut_detection* det, const int32_t* count_dev, int64_t B, int N,
                      const int32_t* tlabel, const float* tpos, double thr, paut_metrics* out) {
  SetMetrics* per_set = static_cast<SetMetrics*>(c.alloc(sizeof(SetMetrics) * (size_t)B));
  if (c.dry) return;
  PAUT_CHECK(rule == 0 || rule == 1, PAUT_ERR_INVALID, "metrics_match: rule must be 0 or 1");
  PAUT_CHECK(rule == 0 || N <= 1024, PAUT_ERR_UNSUPPORTED, "metrics_match rule 1: at most 1024 A-scans per set");
  const int64_t blocks = (B * 32 + 127) / 128;
  k_metrics_match<<<(unsigned)blocks, 128, 0, c.stream>>>(rule, det, count_dev, B, N, tlabel, tpos, (float)thr, per_set);
  c.launched("metrics_match");
  k_metrics_reduce<<<1, 1024, 0, c.stream>>>(per_set, B, out);
  c.launched("metrics_reduce");
}

// acc_metrics_hybrid_binary_dynamic_.py:73-94: preds = probs >= thr (fp32 tensor comparison: the Python scalar is
// rounded to fp32), y = labels > 0.5; confusion counts.  ge = 0 selects the strict '>' of test_detection.py:77.
__global__ void __launch_bounds__(256) k_metrics_confusion(const float* __restrict__ prob,
                                                            const float* __restrict__ label, int64_t M, float thr,
                                                            int ge, unsigned long long* __restrict__ counts) {
  unsigned tp = 0, fp = 0, tn = 0, fn = 0;
  for (int64_t m = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; m < M; m += (int64_t)gridDim.x * blockDim.x) {
    const float p = prob[m];
    const bool pred = ge ? (p >= thr) : (p > thr);
    const bool y = label[m] > 0.5f;
    tp += (pred && y); fp += (pred && !y); tn += (!pred && !y); fn += (!pred && y);
  }
  for (int o = 16; o > 0; o >>= 1) {
    tp += __shfl_xor_sync(0xffffffffu, tp, o); fp += __shfl_xor_sync(0xffffffffu, fp, o);
    tn += __shfl_xor_sync(0xffffffffu, tn, o); fn += __shfl_xor_sync(0xffffffffu, fn, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (tp) atomicAdd(counts + 0, (unsigned long long)tp);
    if (fp) atomicAdd(counts + 1, (unsigned long long)fp);
    if (fn) atomicAdd(counts + 2, (unsigned long long)fn);
    if (tn) atomicAdd(counts + 3, (unsigned long long)tn);
  }
}
__global__ void k_metrics_confusion_out(const unsigned long long* __restrict__ counts, paut_metrics* __restrict__ out) {
  out->tp = (int64_t)counts[0]; out->fp = (int64_t)counts[1]; out->fn = (int64_t)counts[2]; out->tn = (int64_t)counts[3];
  out->sum_iou = 0.0; out->sum_position_error = 0.0;
}

void op_metrics_confusion(Ctx& c, const float* prob, const float* label, int64_t M, double thr, int ge,
                          paut_metrics* out) {
  unsigned long long* counts = static_cast<unsigned long long*>(c.alloc(4 * sizeof(unsigned long long)));
  if (c.dry) return;
  PAUT_CUDA(cudaMemsetAsync(counts, 0, 4 * sizeof(unsigned long long), c.stream));
  k_metrics_confusion<<<grid_cap(M, 256, 148 * 8), 256, 0, c.stream>>>(prob, label, M, (float)thr, ge, counts);
  c.launched("metrics_confusion");
  k_metrics_confusion_out<<<1, 1, 0, c.stream>>>(counts, out);
  c.launched("metrics_confusion_out");
}

}  // namespace paut
