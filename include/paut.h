/*
 * paut.h -- C ABI of libpaut.so: B200 (sm_100a) batched inference of the PAUT
 * A-scan signal models of CSMaus/DefectDetection_viaObjectDetection.
 *
 * The reference has no FFI layer: its boundary for this path is the Python
 * nn.Module surface.  This header is the boundary a
 * Python (ctypes), C or C++ caller binds instead; each entry point names the
 * reference interface it replaces (paths relative to the reference root).
 *
 * Conventions
 *   - plain pointers and sizes only; every data pointer is a DEVICE pointer
 *     unless the name ends in _host;
 *   - one paut_ctx <-> one GPU <-> one CUDA stream; calls on a ctx are
 *     asynchronous on that stream and never synchronise, except
 *     paut_model_finalize (one-time weight packing), paut_ctx_destroy and the
 *     *_host convenience calls;
 *   - a ctx is not thread-safe; different ctxs may be used from different
 *     threads concurrently;
 *   - the library never allocates caller-visible memory: outputs are written
 *     to caller-owned buffers; its private workspace grows on demand;
 *   - every function returns PAUT_OK (0) or a negative paut_status;
 *     paut_last_error() gives the message of the last failure on that ctx.
 *   - there is NO CPU fallback: without a CUDA device every compute call
 *     fails with PAUT_ERR_CUDA.
 */
#ifndef PAUT_H_
#define PAUT_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* bumped on incompatible changes only; entry points added since (paut_debug_stage, paut_json_beam_status) keep it at 1 */
#define PAUT_ABI_VERSION 1

typedef enum {
  PAUT_OK = 0,
  PAUT_ERR_INVALID = -1,     /* bad argument / shape / unknown key            */
  PAUT_ERR_CUDA = -2,        /* CUDA runtime error (message has the detail)    */
  PAUT_ERR_STATE = -3,       /* call order (forward before finalize, ...)      */
  PAUT_ERR_MISSING = -4,     /* finalize: a state_dict tensor was never set    */
  PAUT_ERR_UNSUPPORTED = -5  /* configuration outside what the kernels cover   */
} paut_status;

/* Model kinds = the reference classes on the hot path. */
typedef enum {
  PAUT_MODEL_MSC = 0,        /* signals/multisignalNN/NN_models.py:45-128   MultiSignalClassifier   */
  PAUT_MODEL_MSC_N = 1,      /* signals/multisignalNN/NN_models.py:198-246  MultiSignalClassifier_N */
  PAUT_MODEL_CONV1D_MSC = 2, /* signals/MSC_Conv1D_training.py:50-89        DefectDetectionModel    */
  PAUT_MODEL_SSD = 3,        /* SignalSequenceDetection/model.py:230-343    SignalSequenceDetector  */
  PAUT_MODEL_ENHANCED = 4,   /* SignalSequenceDetection/enhanced_model.py:449-566                   */
  PAUT_MODEL_TWO_STAGE = 5,  /* SignalSequenceDetection/two_stage_model.py:254-312                  */
  /* SURVEY section 8 "next" rows f2 / f3 */
  PAUT_MODEL_MSC_LEGACY = 6, /* signals/resaveModelOnnx.py:7-33 (= GNN_testing_multi_v2_MAP.py:16-36): the no-conv
                                MultiSignalClassifier the repository ships trained .pth weights for             */
  PAUT_MODEL_IMPROVED = 7,   /* signals/improved_multisignal/improved_model.py:69-157 ImprovedMultiSignalClassifier */
  PAUT_MODEL_HYBRID = 8,     /* signals/improved_multisignal/detection_models/hybrid_binary.py:83-168 HybridBinaryModel */
  PAUT_MODEL_COMPLEX = 9     /* .../detection_models/complex_detection_model.py:6-96 ComplexDetectionModel       */
} paut_model_kind;

typedef enum { PAUT_F32 = 0, PAUT_BF16 = 1, PAUT_I64 = 2 } paut_dtype;

/* Arithmetic mode of the encoder (the per-A-scan conv/MLP stack).
 * FP32: fp32 CUDA-core math everywhere (logits within 1e-4 of the reference).
 * BF16: bf16 operands on the tcgen05 tensor cores, fp32 accumulate (within 1e-2). */
typedef enum { PAUT_PRECISION_FP32 = 0, PAUT_PRECISION_BF16 = 1 } paut_precision;

/* Constructor arguments of the reference classes (unused ones are ignored per kind):
 *   MSC / MSC_N : (signal_length, hidden_sizes[3], num_heads)        NN_models.py:46,199
 *   CONV1D_MSC  : (signal_length, num_signals_per_set) -- both unused by the layers; nhead fixed 4
 *   SSD         : (signal_length, d_model, num_classes, nhead, num_layers, dim_feedforward) model.py:234-243
 *   ENHANCED    : same six, enhanced_model.py:453-462 (cross-attention heads fixed at 8, :491)
 *   TWO_STAGE   : (signal_length, d_model, num_classes)  two_stage_model.py:258 (nhead 8, 4 layers, ff 512)
 *   MSC_LEGACY  : (signal_length, hidden_sizes[3])  resaveModelOnnx.py:8 (4 heads, fixed)
 *   IMPROVED    : (signal_length, hidden_sizes[3], num_heads=8, num_layers=num_transformer_layers=4) improved_model.py:70
 *   HYBRID      : (signal_length=320, hidden_sizes={256,128,48}, num_heads=8, num_layers=4)  hybrid_binary.py:88
 *   COMPLEX     : (signal_length=320, d_model=64, num_heads=8, num_layers=4)  complex_detection_model.py:11
 * A zero field selects the reference default. */
typedef struct {
  int32_t signal_length;
  int32_t hidden_sizes[3];
  int32_t num_heads;
  int32_t d_model;
  int32_t num_classes;
  int32_t num_layers;
  int32_t dim_feedforward;
  int32_t precision;       /* paut_precision */
  int32_t reserved[6];
} paut_model_cfg;

/* Output slots of paut_forward, per model kind.  Every slot is fp32, contiguous,
 * caller-owned; a NULL slot is skipped (not computed if it is a pure by-product).
 *   MSC, MSC_N  : 0 defect_prob[B,N]  1 defect_start[B,N]  2 defect_end[B,N]        NN_models.py:125-128
 *   CONV1D_MSC  : 0 defect_prob[B,N]                                            MSC_Conv1D_training.py:87
 *   SSD         : 0 class_preds[B,N,C] 1 position_preds[B,N,2] 2 anomaly_scores[B,N,1]
 *                 3 attention_weights[B,N,1]                                        model.py:337-343
 *   ENHANCED    : 0 class_preds[B,N,C] 1 class_uncertainty[B,N,C] 2 position_preds[B,N,2]
 *                 3 position_uncertainty[B,N,2] 4 anomaly_scores[B,N,1] 5 anomaly_uncertainty[B,N,1]
 *                 6 attention_weights[L,B,N,N] (layer-major) 7 context_attention[B,N]
 *                 8 cross_attention[B,N,N]                                   enhanced_model.py:556-566
 *   TWO_STAGE   : 0 defect_logits[B,N,2] 1 defect_probs[B,N,2] 2 defect_uncertainty[B,N,2]
 *                 3 position_preds[B,N,2] 4 position_uncertainty[B,N,2]      two_stage_model.py:305-312
 *   MSC_LEGACY  : 0 outputs[B,N] (sigmoid)                                       resaveModelOnnx.py:33
 *   IMPROVED    : 0 defect_prob[B,N] 1 defect_start[B,N] 2 defect_end[B,N] (clamped) improved_model.py:147-157
 *   HYBRID      : 0 defect_prob[B,N]                                             hybrid_binary.py:165-168
 *   COMPLEX     : 0 detection_prob[B,N]                                complex_detection_model.py:93-96 */
#define PAUT_MAX_OUTPUTS 12
typedef struct {
  void* slot[PAUT_MAX_OUTPUTS];
} paut_outputs;

/* One kept prediction; replaces the dict the reference predict() appends
 * (model.py:465-471, enhanced_model.py:793-803, two_stage_model.py:490-497) plus the integer
 * sample indices of predict.py:111-113 / signal_visualizer.py:409-410. */
typedef struct {
  int32_t set_index;    /* b                                                    */
  int32_t position;     /* i, index of the A-scan inside its set                */
  int32_t cls;          /* pred_class (1 for the two-stage and MSC rules)       */
  int32_t start_index;  /* trunc(RN_fp32(start * S))                            */
  int32_t end_index;    /* trunc(RN_fp32(end * S))                              */
  float start, end;     /* defect_position (fp32, as returned by forward)       */
  float score;          /* class_score / defect_prob                            */
  float uncertainty;    /* class_unc[pred_class] / defect_unc[...,1]; else 0    */
  float anomaly;        /* anomaly score (SSD, ENHANCED); else 0                */
  double confidence;    /* fp64 value compared with the threshold               */
} paut_detection;       /* 48 bytes */

typedef struct paut_ctx paut_ctx;
typedef struct paut_model paut_model;

int paut_abi_version(void);

/* Context: replaces the reference's device selection idiom (predict.py:215, training_01.py:125).
 * cuda_stream is a cudaStream_t (NULL = the legacy default stream). */
int paut_ctx_create(int device, void* cuda_stream, paut_ctx** out);
void paut_ctx_destroy(paut_ctx* ctx);
const char* paut_last_error(const paut_ctx* ctx); /* ctx may be NULL: last create failure */
/* Upper bound (bytes) of the private activation workspace a forward may use; default min(16 GiB, device memory / 10).
 * Larger batches are processed in resident chunks of whole sets. */
int paut_ctx_set_workspace_limit(paut_ctx* ctx, uint64_t bytes);

/* Model lifetime: replaces  Model(**ctor).to(device); load_state_dict(sd); eval()
 * (predict.py:14-49, model_pred.py:11-13). */
int paut_model_create(paut_ctx* ctx, int model_kind, const paut_model_cfg* cfg, paut_model** out);
void paut_model_destroy(paut_model* m);
/* One state_dict entry (key exactly as in the reference state_dict).  ptr may be host or device
 * memory; the tensor is copied.  dtype F32 (or I64 for num_batches_tracked, ignored). */
int paut_model_set_tensor(paut_model* m, const char* key, const void* ptr, int dtype,
                          const int64_t* shape, int ndim);
/* Folds BatchNorm, packs weights into kernel layouts.  Fails with PAUT_ERR_MISSING (and the
 * key name in paut_last_error) if a tensor of the contract was not set. */
int paut_model_finalize(paut_model* m);
/* Number of state_dict keys of this kind/cfg, and the i-th key + shape (contract introspection). */
int paut_model_num_keys(const paut_model* m);
int paut_model_key(const paut_model* m, int i, const char** key, int64_t* shape4, int* ndim);

/* forward(x): replaces Model.forward in eval mode under no_grad
 * (NN_models.py:108, :225; MSC_Conv1D_training.py:78; model.py:287; enhanced_model.py:502;
 * two_stage_model.py:273).  x is [B,N,S] contiguous (CONV1D_MSC: [B,S,N], as the reference),
 * x_dtype F32 or BF16.  Asynchronous on the ctx stream. */
int paut_forward(paut_model* m, const void* x, int x_dtype, int64_t B, int64_t N, int64_t S,
                 const paut_outputs* out);

/* predict() post-processing: replaces the Python loops model.py:446-473, enhanced_model.py:766-805,
 * two_stage_model.py:477-499, the MSC threshold of model_pred.py:82-85, and the index conversion
 * predict.py:111-113.  `outs` holds the forward outputs (same slots).  Records are written in the
 * reference's (b, i) order to det[0..count) (capacity B*N), *count_dev receives the count.
 * Confidence arithmetic is fp64, indices are trunc(fp32 product) -- bit-exact with the reference
 * given identical forward outputs. */
int paut_postprocess(paut_model* m, const paut_outputs* outs, int64_t B, int64_t N, int64_t S,
                     double threshold, paut_detection* det, int32_t* count_dev);

/* Keep rule of paut_postprocess for the single-probability kinds (start/end are 0 where the model has none):
 *   MSC, MSC_N, CONV1D_MSC : float64(prob) >  threshold   model_pred.py:84, evalMSC.py:91
 *   MSC_LEGACY             : float64(prob) >= threshold   teststtt.py:60,66 (tolist() -> Python floats)
 *   IMPROVED               : float64(prob) >= threshold   improved_model.py:178-181 (.item())
 *   HYBRID                 : prob >= float32(threshold)   acc_metrics_hybrid_binary_dynamic_.py:84 (tensor compare)
 *   COMPLEX                : prob >  float32(threshold)   test_detection.py:77 (tensor compare) */

/* Windowing + cast (a0): gathers fixed-length sets from a resident volume [G, n, S]
 * (json_dataset.py:84-103, dataset_preparation.py:222-282, cast json_dataset.py:112-116).
 * table_dev is int32 [W,3] = (group, start, valid_len) rows; rows >= valid_len are zero.
 * sets_dev is [W, L, S] in dst_dtype. */
int paut_window_gather(paut_ctx* ctx, const void* volume, int src_dtype, int64_t G, int64_t n, int64_t S,
                       const int32_t* table_dev, int64_t W, int64_t L, void* sets_dev, int dst_dtype);
/* The all-zero-run drop of dataset_preparation.py:205 on a resident volume [G, n, S] (F32 or BF16):
 * flags_dev[g] = 1 iff any sample of group g is non-zero (np.all(signals == 0) is False). */
int paut_group_nonzero(paut_ctx* ctx, const void* volume, int dtype, int64_t G, int64_t n, int64_t S, int32_t* flags_dev);
/* Host-side window tables for the two reference rules (rule 0 = signals/ json_dataset rule,
 * 1 = SignalSequenceDetection rule).  Writes up to cap (start, valid_len) pairs, returns the count. */
int paut_window_table_host(int rule, int64_t n, int64_t L, int32_t* pairs_host, int cap);

/* f1 -- host-side loader of the reference's on-disk volume format (improved_multisignal/README.md:67-89):
 *   {"<beam>": {"<scan>_<label>[_<start>-<end>]": [S numbers] | {"signal": [...]}, ...}, ...}
 * Restates the parsing half of JsonSignalDataset._load_all_json_files (json_dataset.py:36-81): beams in file order,
 * scans stably sorted by int(key.split('_')[0]), label 0 iff the second field is "Health", defect range from the
 * third field ("<start>-<end>", [0, 0] on any parse failure), samples as float32(float64(text)).  The windowing
 * half (json_dataset.py:84-160) is paut_window_table_host(rule 0) + paut_window_gather.  Pure host calls. */
typedef struct paut_json_volume paut_json_volume;
int paut_json_load_host(const char* path, paut_json_volume** out);
void paut_json_free(paut_json_volume* v);
const char* paut_json_last_error(void);
int paut_json_num_beams(const paut_json_volume* v);
/* signal_length is -1 when the scans of the beam differ in length */
int paut_json_beam_info(const paut_json_volume* v, int beam, const char** key, int64_t* n_scans, int64_t* signal_length);
/* Where JsonSignalDataset would raise out of its per-file try block (keeping the sequences of the EARLIER beams of the
 * file, json_dataset.py:38,158): a scan key without an integer prefix breaks the sort of every beam (:48); a key without a
 * label field breaks only beams with at least seq_length scans (:51-52,:69).  Returns the PAUT_JSON_BEAM_* bits of the beam
 * (0 = clean) or a negative paut_status; *message (may be NULL) describes the first problem.  A beam with
 * PAUT_JSON_BEAM_BAD_ORDER_KEY is left in file order.  Repeated scan keys keep the last value at the first position, like
 * Python's json. */
#define PAUT_JSON_BEAM_BAD_ORDER_KEY 1
#define PAUT_JSON_BEAM_NO_LABEL 2
int paut_json_beam_status(const paut_json_volume* v, int beam, const char** message);
/* returned by paut_json_scan_copy_host for a scan the reference skips (an object without a "signal" list: the window that
 * contains it is dropped, json_dataset.py:108-131) */
#define PAUT_JSON_SCAN_SKIPPED (-100)
/* full key ("<scan>_<label>[_<start>-<end>]") of the i-th sorted scan of a beam; NULL when out of range */
const char* paut_json_scan_key(const paut_json_volume* v, int beam, int64_t i);
/* samples of the i-th sorted scan of a beam (for beams whose scans differ in length): copies min(cap, length)
 * values into out (may be NULL) and returns the scan's length, or a negative paut_status */
int64_t paut_json_scan_copy_host(const paut_json_volume* v, int beam, int64_t i, float* out, int64_t cap);
/* sorted scans of one beam: signals float32 [n,S], labels int32 [n], defects float32 [n,2], scan_order int64 [n]
 * (the integer prefix of the key); any pointer may be NULL */
int paut_json_beam_copy_host(const paut_json_volume* v, int beam, float* signals, int32_t* labels, float* defects,
                             int64_t* scan_order);

/* f3 -- reference signal and difference matrix (signals/teststtt.py:54-69).  Per set b: reference[b,:] = mean of
 * the A-scans with prob < threshold (fp64 accumulation in set order, like np.mean(axis=0) of the float64 rows),
 * diff[b,i,:] = |x[b,i,:] - reference[b,:]| where prob >= threshold, zeros elsewhere.  A set without a healthy
 * A-scan (the reference returns None and skips it) gets healthy_count 0 and zero rows.  x is [B,N,S] F32 or BF16;
 * reference [B,S], diff [B,N,S] fp32 (either may be NULL), healthy_count int32 [B] (may be NULL). */
int paut_difference_matrix(paut_ctx* ctx, const void* x, int x_dtype, const float* prob, int64_t B, int64_t N,
                           int64_t S, double threshold, float* reference, float* diff, int32_t* healthy_count);

/* f4 -- detection-level metrics on device.  The fp64 sums are accumulated in a fixed order (reproducible);
 * precision / recall / F1 / means are three divisions the caller does on these numbers. */
typedef struct {
  int64_t tp, fp, fn, tn;          /* tn only from paut_metrics_confusion                            */
  double sum_iou;                  /* rule 1: sum of the matched IoUs (mean_iou = sum_iou / tp)      */
  double sum_position_error;       /* rule 0: sum of (|ds| + |de|) / 2 over the matches              */
} paut_metrics;
/* det/count_dev: the records paut_postprocess wrote ((set, position) order).  Targets are dense:
 * target_label int32 [B,N] (> 0 = defect, the class for rule 1), target_pos fp32 [B,N,2] = (start, end).
 *   rule 0: two_stage_train.py:284-375 calculate_metrics -- match on the same position, TP iff IoU > iou_threshold
 *   rule 1: train.py:279-361 calculate_metrics -- greedy first match in target order on equal class, IoU > iou_threshold
 *           (0.5 in the reference)
 * IoU arithmetic is fp32, as in the reference (numpy float32 scalars on both sides).  out_dev: one paut_metrics. */
int paut_metrics_match(paut_ctx* ctx, int rule, const paut_detection* det, const int32_t* count_dev, int64_t B,
                       int64_t N, const int32_t* target_label, const float* target_pos, double iou_threshold,
                       paut_metrics* out_dev);
/* acc_metrics_hybrid_binary_dynamic_.py:73-94: preds = prob >= float32(threshold) (ge = 1; ge = 0: '>',
 * test_detection.py:77), y = label > 0.5; TP / FP / FN / TN over M values. */
int paut_metrics_confusion(paut_ctx* ctx, const float* prob, const float* label, int64_t M, double threshold, int ge,
                           paut_metrics* out_dev);

/* One fused linear layer C[M,N] = act(A[M,K] W[N,K]^T + bias) on device buffers A, C (fp32, contiguous);
 * W and bias may be host or device memory and are packed on every call (a unit-test / stand-alone entry,
 * not a hot path; synchronises).  act: 0 none, 1 ReLU, 2 GELU(erf), 3 sigmoid, 4 softplus.
 * impl: 0 = fp32 CUDA cores (F.linear within fp32 rounding), 1 = bf16 operands on tcgen05, fp32 accumulate. */
int paut_op_linear(paut_ctx* ctx, const float* A, int64_t M, int K, const float* W, const float* bias, int N,
                   float* C, int act, int impl);

/* Micro-benchmark of tcgen05.mma shapes and descriptor experiments (tools/mma_probe.py); debugging aid. */
int paut_debug_mma(paut_ctx* ctx, int mode, int N, int reps, int lbo, int alt, float* out_dev);

/* Intermediate tensor of one fused kernel, for kernel-level parity tests (tests/test_gpu_fused.py); x as in
 * paut_forward, one resident chunk, synchronises.  stage 1: pooled encoder features of the two-stage model
 * (MultiScaleSignalEncoder before the projection, two_stage_model.py:102-115), out_dev fp32 [B*N, 128];
 * stages 2-5: the MSC attention block (tcgen05 / mma.sync, self / shifted keys) on an fp32 [B, N, 64] input, out_dev
 * [B*N, 64]; stage 6: the fused MSC_N front end (NN_models.py:227-234), out_dev fp32 [B*N, S]. */
int paut_debug_stage(paut_model* model, int stage, const void* x_dev, int x_dtype, int64_t B, int64_t N, int64_t S,
                     float* out_dev);

/* Instrumentation: number of kernels this ctx launched since creation (bench.py's gpu_launches). */
int64_t paut_ctx_launch_count(const paut_ctx* ctx);
/* Per-kernel device timing: between begin and end every launch on the ctx is followed by a CUDA event
 * on the ctx stream; end synchronises and writes one line per kernel name, "name launches total_ms\n",
 * into buf (bench.py's roofline numbers come from here, not from a profiler). */
int paut_ctx_profile_begin(paut_ctx* ctx);
int paut_ctx_profile_end(paut_ctx* ctx, char* buf, int64_t cap);

#ifdef __cplusplus
}
#endif
#endif /* PAUT_H_ */
