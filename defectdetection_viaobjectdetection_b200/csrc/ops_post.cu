// Post-processing (the reference predict() loops + integer index conversion) and window gather.
// Both are HBM-bound integer/byte work: coalesced loads, no tensor cores.
#include "common.cuh"

namespace paut {

// Per-A-scan decision, restating model.py:446-473, enhanced_model.py:766-805,
// two_stage_model.py:477-499, model_pred.py:82-85.  fp32 softmax/argmax, then the .item() promotion:
// confidence arithmetic and comparisons in fp64.
struct Decision {
  bool keep;
  int cls;
  float score, unc, anomaly, start, end;
  double conf;
};

__device__ __forceinline__ Decision decide(const PostArgs& p, int64_t m) {
  Decision d;
  d.unc = 0.f;
  d.anomaly = 0.f;
  d.cls = 1;
  if (p.kind == PAUT_MODEL_SSD || p.kind == PAUT_MODEL_ENHANCED) {
    const float* lg = p.score_src + m * p.C;
    float mx = lg[0];
    for (int c = 1; c < p.C; ++c) mx = fmaxf(mx, lg[c]);
    float sum = 0.f;
    for (int c = 0; c < p.C; ++c) sum += expf(lg[c] - mx);
    int best = 0;
    float bestp = -1.f;
    for (int c = 0; c < p.C; ++c) {
      const float pc = expf(lg[c] - mx) / sum;
      if (pc > bestp) { bestp = pc; best = c; }      // first maximum, like torch.argmax
    }
    d.cls = best;
    d.score = bestp;
    d.anomaly = p.anomaly[m];
    if (p.kind == PAUT_MODEL_ENHANCED) {
      d.unc = p.unc[m * p.C + best];
      d.conf = (double)d.score / (1.0 + (double)d.unc);
    } else {
      d.conf = (double)d.score;
    }
    d.keep = (d.conf > p.threshold) || (best > 0 && (double)d.anomaly > p.threshold);
    d.start = p.pos[m * 2];
    d.end = p.pos[m * 2 + 1];
  } else if (p.kind == PAUT_MODEL_TWO_STAGE) {
    d.score = p.score_src[m * 2 + 1];
    d.unc = p.unc[m * 2 + 1];
    d.conf = (double)d.score / (1.0 + (double)d.unc);
    d.keep = d.conf > p.threshold;
    d.start = p.pos[m * 2];
    d.end = p.pos[m * 2 + 1];
  } else {  // MSC family
    d.score = p.score_src[m];
    d.conf = (double)d.score;
    switch (p.cmp) {                                   // include/paut.h "Keep rule"
      case 1: d.keep = d.conf >= p.threshold; break;
      case 2: d.keep = d.score >= (float)p.threshold; break;
      case 3: d.keep = d.score > (float)p.threshold; break;
      default: d.keep = d.conf > p.threshold; break;
    }
    d.start = p.start ? p.start[m] : 0.f;
    d.end = p.end ? p.end[m] : 0.f;
  }
  return d;
}

// pass 1: per-set kept count
__global__ void k_post_count(PostArgs p, int32_t* __restrict__ counts) {
  __shared__ int wsum[8];
  const int64_t b = blockIdx.x;
  int local = 0;
  for (int i = threadIdx.x; i < p.N; i += blockDim.x) local += decide(p, b * p.N + i).keep ? 1 : 0;
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    int s = 0;
    for (int w = 0; w < (blockDim.x >> 5); ++w) s += wsum[w];
    counts[b] = s;
  }
}

// pass 2: exclusive scan of the per-set counts (single CTA, sequential over chunks of 1024)
__global__ void k_post_scan(const int32_t* __restrict__ counts, int32_t* __restrict__ offsets, int64_t B,
                            int32_t* __restrict__ total) {
  __shared__ int buf[1024];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int64_t base = 0; base < B; base += 1024) {
    const int64_t i = base + threadIdx.x;
    const int v = i < B ? counts[i] : 0;
    buf[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      const int t = threadIdx.x >= o ? buf[threadIdx.x - o] : 0;
      __syncthreads();
      buf[threadIdx.x] += t;
      __syncthreads();
    }
    if (i < B) offsets[i] = carry + buf[threadIdx.x] - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry += buf[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry;
}

// pass 3: ordered emission.  One warp per set: ballot gives each kept A-scan its rank in (b, i) order.
__global__ void k_post_emit(PostArgs p, const int32_t* __restrict__ offsets, paut_detection* __restrict__ det) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const float fS = (float)p.S;
  for (int64_t b = warp; b < p.B; b += nwarps) {
    int base = offsets[b];
    for (int i0 = 0; i0 < p.N; i0 += 32) {
      const int i = i0 + lane;
      Decision d;
      d.keep = false;
      if (i < p.N) d = decide(p, b * p.N + i);
      const unsigned mask = __ballot_sync(0xffffffffu, d.keep);
      if (d.keep) {
        paut_detection r;
        r.set_index = (int32_t)b;
        r.position = i;
        r.cls = d.cls;
        r.start_index = __float2int_rz(__fmul_rn(d.start, fS));   // trunc(RN_fp32(start * S))
        r.end_index = __float2int_rz(__fmul_rn(d.end, fS));
        r.start = d.start;
        r.end = d.end;
        r.score = d.score;
        r.uncertainty = d.unc;
        r.anomaly = d.anomaly;
        r.confidence = d.conf;
        det[base + __popc(mask & ((1u << lane) - 1u))] = r;
      }
      base += __popc(mask);
    }
  }
}

void op_postprocess(Ctx& c, const PostArgs& a, paut_detection* det, int32_t* count_dev) {
  if (c.dry) return;
  PAUT_CHECK(a.B > 0 && a.N > 0, PAUT_ERR_INVALID, "postprocess: empty batch");
  PAUT_CHECK(a.B < (int64_t(1) << 31), PAUT_ERR_INVALID, "postprocess: too many sets");
  int32_t* counts = static_cast<int32_t*>(c.alloc(sizeof(int32_t) * a.B));
  int32_t* offsets = static_cast<int32_t*>(c.alloc(sizeof(int32_t) * a.B));
  k_post_count<<<(unsigned)a.B, 64, 0, c.stream>>>(a, counts);
  c.launched("post_count");
  k_post_scan<<<1, 1024, 0, c.stream>>>(counts, offsets, a.B, count_dev);
  c.launched("post_scan");
  int64_t blocks = (a.B * 32 + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  k_post_emit<<<(unsigned)blocks, 256, 0, c.stream>>>(a, offsets, det);
  c.launched("post_emit");
}

// ------------------------------------------------------------------------------------------ window gather
// sets[w, r, :] = volume[g, start + r, :] for r < valid else 0, with dtype conversion.  4 samples / thread.
template <typename TS, typename TD>
__global__ void k_window_gather(const TS* __restrict__ vol, int64_t n, int S, const int32_t* __restrict__ table,
                                int64_t W, int L, TD* __restrict__ sets) {
  const int S4 = S >> 2;
  const int64_t total = W * L * S4;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int s4 = (int)(idx % S4);
    const int64_t row = idx / S4;
    const int r = (int)(row % L);
    const int64_t w = row / L;
    const int g = table[w * 3], start = table[w * 3 + 1], valid = table[w * 3 + 2];
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (r < valid) {
      const TS* src = vol + ((int64_t)g * n + start + r) * S + s4 * 4;
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = (float)src[j];
    }
    TD* dst = sets + row * S + s4 * 4;
#pragma unroll
    for (int j = 0; j < 4; ++j) dst[j] = (TD)v[j];
  }
}
void op_window_gather(Ctx& c, const void* volume, int src_dtype, int64_t G, int64_t n, int S, const int32_t* table,
                      int64_t W, int L, void* sets, int dst_dtype) {
  if (c.dry) return;
  (void)G;
  PAUT_CHECK(S % 4 == 0, PAUT_ERR_UNSUPPORTED, "window gather: S must be a multiple of 4");
  if (W == 0) return;
  const int64_t total = W * L * (S / 4);
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  const unsigned g = (unsigned)blocks;
  using bf = __nv_bfloat16;
  if (src_dtype == PAUT_F32 && dst_dtype == PAUT_F32)
    k_window_gather<float, float><<<g, 256, 0, c.stream>>>((const float*)volume, n, S, table, W, L, (float*)sets);
  else if (src_dtype == PAUT_F32 && dst_dtype == PAUT_BF16)
    k_window_gather<float, bf><<<g, 256, 0, c.stream>>>((const float*)volume, n, S, table, W, L, (bf*)sets);
  else if (src_dtype == PAUT_BF16 && dst_dtype == PAUT_F32)
    k_window_gather<bf, float><<<g, 256, 0, c.stream>>>((const bf*)volume, n, S, table, W, L, (float*)sets);
  else if (src_dtype == PAUT_BF16 && dst_dtype == PAUT_BF16)
    k_window_gather<bf, bf><<<g, 256, 0, c.stream>>>((const bf*)volume, n, S, table, W, L, (bf*)sets);
  else
    throw Error(PAUT_ERR_INVALID, "window gather: dtype must be F32 or BF16");
  c.launched("window_gather");
}

}  // namespace paut
