// extern "C" surface of libpaut.so (declared in include/paut.h).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <mutex>
#include <new>
#include <set>

#include "model.cuh"

namespace paut {

void Ctx::reserve(size_t bytes) {
  if (bytes <= ws_cap) return;
  if (ws) {
    PAUT_CUDA(cudaStreamSynchronize(stream));
    PAUT_CUDA(cudaFree(ws));
    ws = nullptr;
    ws_cap = 0;
  }
  void* p = nullptr;
  PAUT_CUDA(cudaMalloc(&p, bytes));
  ws = static_cast<char*>(p);
  ws_cap = bytes;
}

void* Ctx::alloc(size_t bytes) {
  const size_t off = (ws_off + 255) & ~size_t(255);
  ws_off = off + bytes;
  if (!dry && ws_off > ws_cap)
    throw Error(PAUT_ERR_STATE, "internal: activation workspace overflow (request " + std::to_string(bytes) + " B at offset " +
                                    std::to_string(off) + ", capacity " + std::to_string(ws_cap) + ", limit " + std::to_string(ws_limit) + ")");
  // a dry run on a fresh context has no workspace yet: hand out addresses from a fake, well-aligned, non-null base so
  // that "is this optional buffer present" tests in the model code take the same branch as the real run
  if (dry && !ws) return reinterpret_cast<char*>(uintptr_t(1) << 40) + off;
  return ws + off;
}

void smem_optin_once(Ctx& c, const void* kernel) {
  static std::mutex mu;
  static std::set<std::pair<int, const void*>> done;
  std::lock_guard<std::mutex> lock(mu);
  if (done.count({c.device, kernel})) return;
  cudaFuncAttributes fa;
  PAUT_CUDA(cudaFuncGetAttributes(&fa, kernel));
  const int dyn = c.smem_optin - (int)fa.sharedSizeBytes;
  PAUT_CHECK(dyn > 0, PAUT_ERR_UNSUPPORTED, "kernel's static shared memory exceeds the device limit");
  PAUT_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn));
  done.insert({c.device, kernel});
}

void Ctx::launched(const char* what) {
  ++launches;
  if (profiling) {
    cudaEvent_t ev;
    if (cudaEventCreate(&ev) == cudaSuccess) {
      cudaEventRecord(ev, stream);
      prof_events.emplace_back(what, ev);
    }
  }
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    cudaGetLastError();
    throw Error(PAUT_ERR_CUDA, std::string("kernel launch failed (") + what + "): " + cudaGetErrorString(e));
  }
}

}  // namespace paut

using paut::Ctx;
using paut::Error;
using paut::Model;

struct paut_ctx {
  Ctx c;
};
struct paut_model {
  Model m;
};

static thread_local std::string g_create_error;

template <typename F>
static int guarded(Ctx* c, F&& f) {
  try {
    f();
    return PAUT_OK;
  } catch (const Error& e) {
    if (c) c->last_error = e.what(); else g_create_error = e.what();
    return e.code;
  } catch (const std::exception& e) {
    if (c) c->last_error = e.what(); else g_create_error = e.what();
    return PAUT_ERR_INVALID;
  }
}

extern "C" {

int paut_abi_version(void) { return PAUT_ABI_VERSION; }

int paut_ctx_create(int device, void* cuda_stream, paut_ctx** out) {
  return guarded(nullptr, [&] {
    PAUT_CHECK(out != nullptr, PAUT_ERR_INVALID, "ctx_create: out is null");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
      throw Error(PAUT_ERR_CUDA, std::string("no CUDA device available (libpaut has no CPU fallback): ") +
                                     cudaGetErrorString(e));
    PAUT_CHECK(device >= 0 && device < n, PAUT_ERR_INVALID, "ctx_create: bad device index");
    PAUT_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    PAUT_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
      throw Error(PAUT_ERR_UNSUPPORTED, std::string("libpaut is built for sm_100a only; device is ") + prop.name);
    paut_ctx* p = new paut_ctx();
    p->c.device = device;
    p->c.stream = static_cast<cudaStream_t>(cuda_stream);
    p->c.num_sms = prop.multiProcessorCount;
    p->c.smem_optin = (int)prop.sharedMemPerBlockOptin;
    // resident chunk budget: a tenth of the device memory, at most 16 GiB (five contexts -- one per scanner
    // lane plus the caller's -- stay under half of a B200's 180 GB); larger chunks mean fuller grids
    p->c.ws_limit = std::min<size_t>(size_t(16) << 30, std::max<size_t>(prop.totalGlobalMem / 10, size_t(256) << 20));
    *out = p;
  });
}

void paut_ctx_destroy(paut_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->c.device);
  cudaStreamSynchronize(ctx->c.stream);
  if (ctx->c.ws) cudaFree(ctx->c.ws);
  delete ctx;
}

const char* paut_last_error(const paut_ctx* ctx) { return ctx ? ctx->c.last_error.c_str() : g_create_error.c_str(); }

int paut_ctx_set_workspace_limit(paut_ctx* ctx, uint64_t bytes) {
  if (!ctx) return PAUT_ERR_INVALID;
  return guarded(&ctx->c, [&] {
    PAUT_CHECK(bytes >= (uint64_t(64) << 20), PAUT_ERR_INVALID, "workspace limit must be at least 64 MiB");
    ctx->c.ws_limit = (size_t)bytes;
  });
}

int64_t paut_ctx_launch_count(const paut_ctx* ctx) { return ctx ? ctx->c.launches : -1; }

int paut_ctx_profile_begin(paut_ctx* ctx) {
  if (!ctx) return PAUT_ERR_INVALID;
  return guarded(&ctx->c, [&] {
    Ctx& c = ctx->c;
    PAUT_CUDA(cudaSetDevice(c.device));
    for (auto& e : c.prof_events) cudaEventDestroy(e.second);
    c.prof_events.clear();
    c.profiling = true;
    cudaEvent_t ev;
    PAUT_CUDA(cudaEventCreate(&ev));
    PAUT_CUDA(cudaEventRecord(ev, c.stream));
    c.prof_events.emplace_back("", ev);
  });
}

int paut_ctx_profile_end(paut_ctx* ctx, char* buf, int64_t cap) {
  if (!ctx) return PAUT_ERR_INVALID;
  return guarded(&ctx->c, [&] {
    Ctx& c = ctx->c;
    PAUT_CHECK(c.profiling, PAUT_ERR_STATE, "profile_end without profile_begin");
    c.profiling = false;
    PAUT_CUDA(cudaStreamSynchronize(c.stream));
    std::map<std::string, std::pair<int64_t, double>> agg;   // name -> (launches, total ms)
    for (size_t i = 1; i < c.prof_events.size(); ++i) {
      float ms = 0.f;
      PAUT_CUDA(cudaEventElapsedTime(&ms, c.prof_events[i - 1].second, c.prof_events[i].second));
      auto& a = agg[c.prof_events[i].first];
      a.first += 1;
      a.second += ms;
    }
    for (auto& e : c.prof_events) cudaEventDestroy(e.second);
    c.prof_events.clear();
    std::string out;
    for (const auto& kv : agg)
      out += kv.first + " " + std::to_string(kv.second.first) + " " + std::to_string(kv.second.second) + "\n";
    if (buf && cap > 0) {
      const size_t n = std::min<size_t>(out.size(), (size_t)cap - 1);
      std::memcpy(buf, out.data(), n);
      buf[n] = 0;
    }
  });
}

int paut_model_create(paut_ctx* ctx, int model_kind, const paut_model_cfg* cfg, paut_model** out) {
  if (!ctx) return PAUT_ERR_INVALID;
  return guarded(&ctx->c, [&] {
    PAUT_CHECK(out != nullptr, PAUT_ERR_INVALID, "model_create: out is null");
    *out = nullptr;
    PAUT_CHECK(model_kind >= PAUT_MODEL_MSC && model_kind <= PAUT_MODEL_COMPLEX, PAUT_ERR_INVALID,
               "model_create: unknown model kind");
    paut_model_cfg c{};
    if (cfg) c = *cfg;
    // reference constructor defaults
    const bool seq_kind = model_kind >= PAUT_MODEL_SSD && model_kind <= PAUT_MODEL_TWO_STAGE;
    if (c.signal_length <= 0) c.signal_length = seq_kind ? 100 : 320;
    if (c.hidden_sizes[0] <= 0) {
      if (model_kind == PAUT_MODEL_HYBRID) { c.hidden_sizes[0] = 256; c.hidden_sizes[1] = 128; c.hidden_sizes[2] = 48; }
      else { c.hidden_sizes[0] = 128; c.hidden_sizes[1] = 64; c.hidden_sizes[2] = 32; }
    }
    if (c.num_classes <= 0) c.num_classes = 2;
    switch (model_kind) {
      case PAUT_MODEL_MSC:
      case PAUT_MODEL_MSC_N:
        if (c.num_heads <= 0) c.num_heads = 4;
        break;
      case PAUT_MODEL_CONV1D_MSC:
        c.num_heads = 4; c.d_model = 128; c.num_layers = 4; c.dim_feedforward = 2048;
        break;
      case PAUT_MODEL_SSD:
        if (c.d_model <= 0) c.d_model = 128;
        if (c.num_heads <= 0) c.num_heads = 8;
        if (c.num_layers <= 0) c.num_layers = 4;
        if (c.dim_feedforward <= 0) c.dim_feedforward = 512;
        break;
      case PAUT_MODEL_ENHANCED:
        if (c.d_model <= 0) c.d_model = 256;
        if (c.num_heads <= 0) c.num_heads = 8;
        if (c.num_layers <= 0) c.num_layers = 6;
        if (c.dim_feedforward <= 0) c.dim_feedforward = 1024;
        break;
      case PAUT_MODEL_TWO_STAGE:
        if (c.d_model <= 0) c.d_model = 128;
        c.num_heads = 8; c.num_layers = 4; c.dim_feedforward = 512; c.num_classes = 2;
        break;
      case PAUT_MODEL_MSC_LEGACY:
        c.num_heads = 4;
        break;
      case PAUT_MODEL_IMPROVED:
      case PAUT_MODEL_HYBRID:
        if (c.num_heads <= 0) c.num_heads = 8;
        if (c.num_layers <= 0) c.num_layers = 4;
        break;
      case PAUT_MODEL_COMPLEX:
        if (c.d_model <= 0) c.d_model = 64;
        if (c.num_heads <= 0) c.num_heads = 8;
        if (c.num_layers <= 0) c.num_layers = 4;
        c.dim_feedforward = 2 * c.d_model;
        break;
    }
    PAUT_CHECK(c.num_classes <= 16, PAUT_ERR_UNSUPPORTED, "num_classes must be <= 16");
    PAUT_CHECK(c.precision == PAUT_PRECISION_FP32 || c.precision == PAUT_PRECISION_BF16, PAUT_ERR_INVALID,
               "precision must be PAUT_PRECISION_FP32 or PAUT_PRECISION_BF16");
    if (seq_kind || model_kind == PAUT_MODEL_COMPLEX) {
      PAUT_CHECK(c.d_model % 32 == 0 && c.d_model <= 1024, PAUT_ERR_UNSUPPORTED,
                 "d_model must be a multiple of 32 and <= 1024");
      PAUT_CHECK(c.dim_feedforward % 4 == 0, PAUT_ERR_UNSUPPORTED, "dim_feedforward must be a multiple of 4");
    } else if (model_kind != PAUT_MODEL_CONV1D_MSC) {
      for (int i = 0; i < 3; ++i)
        PAUT_CHECK(c.hidden_sizes[i] % 4 == 0 && c.hidden_sizes[i] >= 16 && c.hidden_sizes[i] <= 1024,
                   PAUT_ERR_UNSUPPORTED, "hidden_sizes must be multiples of 4 in [16, 1024]");
    }
    paut_model* m = new paut_model();
    m->m.ctx = &ctx->c;
    m->m.kind = model_kind;
    m->m.cfg = c;
    try {
      m->m.build_spec();
    } catch (...) {
      delete m;
      throw;
    }
    *out = m;
  });
}

void paut_model_destroy(paut_model* m) {
  if (!m) return;
  cudaSetDevice(m->m.ctx->device);
  cudaStreamSynchronize(m->m.ctx->stream);
  delete m;
}

int paut_model_set_tensor(paut_model* m, const char* key, const void* ptr, int dtype, const int64_t* shape, int ndim) {
  if (!m) return PAUT_ERR_INVALID;
  return guarded(m->m.ctx, [&] { m->m.set_tensor(key, ptr, dtype, shape, ndim); });
}

int paut_model_finalize(paut_model* m) {
  if (!m) return PAUT_ERR_INVALID;
  return guarded(m->m.ctx, [&] { m->m.finalize(); });
}

int paut_model_num_keys(const paut_model* m) { return m ? (int)m->m.spec.size() : PAUT_ERR_INVALID; }

int paut_model_key(const paut_model* m, int i, const char** key, int64_t* shape4, int* ndim) {
  if (!m || i < 0 || i >= (int)m->m.spec.size() || !key || !shape4 || !ndim) return PAUT_ERR_INVALID;
  const auto& k = m->m.spec[i];
  *key = k.key.c_str();
  *ndim = (int)k.shape.size();
  for (int j = 0; j < 4; ++j) shape4[j] = j < (int)k.shape.size() ? k.shape[j] : 1;
  return PAUT_OK;
}

int paut_forward(paut_model* m, const void* x, int x_dtype, int64_t B, int64_t N, int64_t S, const paut_outputs* out) {
  if (!m) return PAUT_ERR_INVALID;
  return guarded(m->m.ctx, [&] {
    PAUT_CHECK(out != nullptr, PAUT_ERR_INVALID, "forward: outputs is null");
    m->m.forward(x, x_dtype, B, N, S, *out);
  });
}

int paut_postprocess(paut_model* m, const paut_outputs* outs, int64_t B, int64_t N, int64_t S, double threshold,
                     paut_detection* det, int32_t* count_dev) {
  if (!m) return PAUT_ERR_INVALID;
  return guarded(m->m.ctx, [&] {
    PAUT_CHECK(outs != nullptr, PAUT_ERR_INVALID, "postprocess: outputs is null");
    m->m.postprocess(*outs, B, N, S, threshold, det, count_dev);
  });
}

int paut_window_gather(paut_ctx* ctx, const void* volume, int src_dtype, int64_t G, int64_t n, int64_t S,
                       const int32_t* table_dev, int64_t W, int64_t L, void* sets_dev, int dst_dtype) {
  if (!ctx) return PAUT_ERR_INVALID;
  return guarded(&ctx->c, [&] {
    PAUT_CHECK(volume && sets_dev && (table_dev || W == 0), PAUT_ERR_INVALID, "window_gather: null pointer");
    PAUT_CHECK(G >= 0 && n > 0 && S > 0 && W >= 0 && L > 0, PAUT_ERR_INVALID, "window_gather: bad sizes");
    PAUT_CUDA(cudaSetDevice(ctx->c.device));
    paut::op_window_gather(ctx->c, volume, src_dtype, G, n, (int)S, table_dev, W, (int)L, sets_dev, dst_dtype);
  });
}

int paut_group_nonzero(paut_ctx* ctx, const void* volume, int dtype, int64_t G, int64_t n, int64_t S, int32_t* flags_dev) {
  if (!ctx) return PAUT_ERR_INVALID;
  return guarded(&ctx->c, [&] {
    PAUT_CHECK(volume && flags_dev && G > 0 && n > 0 && S > 0, PAUT_ERR_INVALID, "group_nonzero: bad arguments");
    PAUT_CHECK(dtype == PAUT_F32 || dtype == PAUT_BF16, PAUT_ERR_INVALID, "group_nonzero: dtype must be F32 or BF16");
    PAUT_CHECK(G < (int64_t(1) << 31), PAUT_ERR_UNSUPPORTED, "group_nonzero: too many groups");
    PAUT_CUDA(cudaSetDevice(ctx->c.device));
    paut::op_group_nonzero(ctx->c, volume, dtype, G, n * S, flags_dev);
  });
}

int paut_difference_matrix(paut_ctx* ctx, const void* x, int x_dtype, const float* prob, int64_t B, int64_t N,
                           int64_t S, double threshold, float* reference, float* diff, int32_t* healthy_count) {
  if (!ctx) return PAUT_ERR_INVALID;
  return guarded(&ctx->c, [&] {
    Ctx& c = ctx->c;
    PAUT_CHECK(x && prob && B > 0 && N > 0 && S > 0, PAUT_ERR_INVALID, "difference_matrix: bad arguments");
    PAUT_CHECK(x_dtype == PAUT_F32 || x_dtype == PAUT_BF16, PAUT_ERR_INVALID, "difference_matrix: dtype must be F32 or BF16");
    PAUT_CHECK(B < 65536, PAUT_ERR_UNSUPPORTED, "difference_matrix: at most 65535 sets per call");
    PAUT_CUDA(cudaSetDevice(c.device));
    c.reset();
    c.reserve((size_t)B * S * sizeof(double) + (size_t)B * sizeof(int32_t) + 1024);
    paut::op_difference_matrix(c, x, x_dtype, prob, B, (int)N, (int)S, threshold, reference, diff, healthy_count);
  });
}

int paut_metrics_match(paut_ctx* ctx, int rule, const paut_detection* det, const int32_t* count_dev, int64_t B,
                       int64_t N, const int32_t* target_label, const float* target_pos, double iou_threshold,
                       paut_metrics* out_dev) {
  if (!ctx) return PAUT_ERR_INVALID;
  return guarded(&ctx->c, [&] {
    Ctx& c = ctx->c;
    PAUT_CHECK(det && count_dev && target_label && target_pos && out_dev && B > 0 && N > 0, PAUT_ERR_INVALID,
               "metrics_match: bad arguments");
    PAUT_CHECK(iou_threshold >= 0.0, PAUT_ERR_INVALID, "metrics_match: iou_threshold must be >= 0");
    PAUT_CHECK(B < (int64_t(1) << 26), PAUT_ERR_UNSUPPORTED, "metrics_match: too many sets");
    PAUT_CUDA(cudaSetDevice(c.device));
    c.reset();
    c.reserve((size_t)B * 32 + 1024);
    paut::op_metrics_match(c, rule, det, count_dev, B, (int)N, target_label, target_pos, iou_threshold, out_dev);
  });
}

int paut_metrics_confusion(paut_ctx* ctx, const float* prob, const float* label, int64_t M, double threshold, int ge,
                           paut_metrics* out_dev) {
  if (!ctx) return PAUT_ERR_INVALID;
  return guarded(&ctx->c, [&] {
    Ctx& c = ctx->c;
    PAUT_CHECK(prob && label && out_dev && M > 0, PAUT_ERR_INVALID, "metrics_confusion: bad arguments");
    PAUT_CUDA(cudaSetDevice(c.device));
    c.reset();
    c.reserve(4096);
    paut::op_metrics_confusion(c, prob, label, M, threshold, ge, out_dev);
  });
}

// One fused linear layer on device buffers (unit tests / stand-alone GEMM).  Weights are packed per call.
int paut_op_linear(paut_ctx* ctx, const float* A, int64_t M, int K, const float* W, const float* bias, int N,
                   float* C, int act, int impl) {
  if (!ctx) return PAUT_ERR_INVALID;
  return guarded(&ctx->c, [&] {
    Ctx& c = ctx->c;
    PAUT_CHECK(A && W && C && M > 0 && K > 0 && N > 0, PAUT_ERR_INVALID, "op_linear: bad arguments");
    PAUT_CUDA(cudaSetDevice(c.device));
    std::vector<float> w((size_t)N * K), wt((size_t)N * K), b(N, 0.f);
    PAUT_CUDA(cudaMemcpy(w.data(), W, w.size() * sizeof(float), cudaMemcpyDefault));
    if (bias) PAUT_CUDA(cudaMemcpy(b.data(), bias, N * sizeof(float), cudaMemcpyDefault));
    for (int n = 0; n < N; ++n)
      for (int k = 0; k < K; ++k) wt[(size_t)k * N + n] = w[(size_t)n * K + k];
    std::vector<void*> tmp;
    auto up = [&](const void* src, size_t bytes) {
      void* p = nullptr;
      PAUT_CUDA(cudaMalloc(&p, bytes));
      tmp.push_back(p);
      PAUT_CUDA(cudaMemcpy(p, src, bytes, cudaMemcpyHostToDevice));
      return p;
    };
    paut::LinArgs a;
    a.A = A; a.lda = K; a.M = M; a.K = K; a.N = N; a.C = C; a.ldc = N; a.act = act;
    a.W = static_cast<const float*>(up(w.data(), w.size() * 4));
    a.Wt = static_cast<const float*>(up(wt.data(), wt.size() * 4));
    a.bias = static_cast<const float*>(up(b.data(), b.size() * 4));
    try {
      if (impl == 1) {
        int stream_b = 0;
        const int nt = paut::tc_pick_ntile(N, K, &stream_b);
        PAUT_CHECK(nt > 0, PAUT_ERR_UNSUPPORTED, "op_linear: tcgen05 path needs N to be a multiple of 16");
        std::vector<uint16_t> packed;
        int Kp = 0;
        paut::tc_pack_weight(w.data(), N, K, nt, packed, &Kp);
        a.Wp = up(packed.data(), packed.size() * 2);
        a.NT = nt;
        a.stream_b = stream_b;
        paut::op_linear_tc(c, a);
      } else {
        paut::op_linear(c, a);
      }
      PAUT_CUDA(cudaStreamSynchronize(c.stream));
    } catch (...) {
      for (void* p : tmp) cudaFree(p);
      throw;
    }
    for (void* p : tmp) cudaFree(p);
  });
}

// intermediate tensors of the fused kernels (kernel-level parity tests); not part of the product path
int paut_debug_stage(paut_model* m, int stage, const void* x, int x_dtype, int64_t B, int64_t N, int64_t S, float* out_dev) {
  if (!m) return PAUT_ERR_INVALID;
  return guarded(m->m.ctx, [&] {
    PAUT_CHECK(x && out_dev && B > 0 && N > 0 && S > 0, PAUT_ERR_INVALID, "debug_stage: bad arguments");
    m->m.debug_stage(stage, x, x_dtype, B, N, S, out_dev);
    PAUT_CUDA(cudaStreamSynchronize(m->m.ctx->stream));
  });
}

// micro-benchmark / descriptor experiments (tools/mma_probe.py); not part of the product path
int paut_debug_mma(paut_ctx* ctx, int mode, int N, int reps, int lbo, int alt, float* out_dev) {
  if (!ctx) return PAUT_ERR_INVALID;
  return guarded(&ctx->c, [&] {
    PAUT_CUDA(cudaSetDevice(ctx->c.device));
    paut::op_debug_mma(ctx->c, mode, N, reps, lbo, alt, out_dev);
    PAUT_CUDA(cudaStreamSynchronize(ctx->c.stream));
  });
}

// Window tables: rule 0 = json_dataset.py:84-103 (end-anchored last window, short runs skipped),
// rule 1 = dataset_preparation.py:222-282 (zero-pad short runs, overlapping windows + tail).
int paut_window_table_host(int rule, int64_t n, int64_t L, int32_t* pairs, int cap) {
  if (n < 0 || L <= 0 || (rule != 0 && rule != 1)) return PAUT_ERR_INVALID;
  int count = 0;
  auto emit = [&](int64_t start, int64_t valid) {
    if (pairs && count < cap) {
      pairs[2 * count] = (int32_t)start;
      pairs[2 * count + 1] = (int32_t)valid;
    }
    ++count;
  };
  if (rule == 0) {
    if (n < L) return 0;
    const int64_t num = (n + L - 1) / L;
    for (int64_t i = 0; i < num; ++i) emit(i < num - 1 ? i * L : n - L, L);
  } else {
    if (n == 0) return 0;
    if (n <= L) {
      emit(0, n);
    } else {
      const int64_t total = std::max<int64_t>(1, (int64_t)std::ceil((double)n / ((double)L / 2.0)));
      const int64_t step =
          total > 1 ? std::max<int64_t>(1, (int64_t)std::floor((double)(n - L) / (double)(total - 1))) : L;
      for (int64_t s = 0; s <= n - L; s += step) emit(s, L);
      if ((n - L) % step != 0) emit(n - L, L);
    }
  }
  return count;
}

}  // extern "C"
