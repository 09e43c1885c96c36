/* siren_b200.h — C ABI of libsirenb200.so: the B200 (sm_100a) implementation of the per-image SIREN fit
 * hot path of varun19299/implicit-image-compression.
 *
 * The reference has no FFI: its boundary is a Python call surface.  Each entry point below
 * names the reference call it replaces (paths relative to the reference root, implicit_image/...).
 *
 * Conventions
 *   - every function returns 0 on success and a negative sirenb200_status on failure; it never throws and
 *     never exits.  sirenb200_last_error() returns a thread-local message for the last failure.
 *   - all pointers are DEVICE pointers unless the parameter name starts with `h_` (host array).
 *   - parameter tensors are fp32, row-major [out, in] weights and [out] biases, ordered as
 *     model.parameters(): w0, b0, w1, b1, ... (2*depth entries).  Pointer arrays are read on every call and
 *     never cached (the reference rebinds weight.data: pipeline/masking/core.py:279, quant/kmeans.py:71).
 *   - calls are asynchronous on `stream`; no hidden host synchronisation unless stated.
 *   - a handle is bound to the device that was current at create time and is not re-entrant.
 */
#ifndef SIREN_B200_H_
#define SIREN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* sirenb200_stream_t;
typedef struct sirenb200_plan* sirenb200_handle_t;

typedef enum {
  SIRENB200_OK = 0,
  SIRENB200_ERR_INVALID = -1,     /* bad argument / unsupported shape */
  SIRENB200_ERR_CUDA = -2,        /* CUDA runtime error (message has the details) */
  SIRENB200_ERR_NO_DEVICE = -3,   /* no sm_100 device */
  SIRENB200_ERR_STATE = -4        /* call sequence error (e.g. backward without grid) */
} sirenb200_status;

typedef enum {
  SIRENB200_PREC_FP32 = 0,  /* fp32 CUDA-core path, any shape */
  SIRENB200_PREC_F16TC = 1  /* fp16 operands / fp32 accumulate on tcgen05 tensor cores (hidden in {128,256}) */
} sirenb200_precision;

/* Siren(...) constructor arguments (models/siren.py:71-121) + the image geometry of compress.py:64-67. */
typedef struct {
  int32_t depth;            /* number of SineLayers including first and last (mlp.depth) */
  int32_t hidden;           /* effective hidden width (after small_dense_density) */
  int32_t in_features;      /* 2 */
  int32_t out_features;     /* 3 (<= 4) */
  float first_omega;        /* mlp.first_omega_0 */
  float hidden_omega;       /* mlp.hidden_omega_0 */
  int32_t outermost_linear; /* mlp.outermost_linear */
  int32_t height, width;    /* full image size H x W: the loss is a mean over H*W*out_features */
  int32_t row_begin;        /* this handle processes image rows [row_begin, row_end) (pixel sharding) */
  int32_t row_end;
  int32_t precision;        /* sirenb200_precision */
  int32_t reserved[4];      /* [0] model family: 0 = Siren, 1 = FourierNet (models/fourier.py:28-69; fp32 path; `depth`
                             *     is then FourierNet's depth, i.e. depth - 1 linear layers, ReLU between them,
                             *     sigmoid on the output); [1] FourierNet map_size; [2], [3] zero */
} sirenb200_config_t;

int sirenb200_version(void);
const char* sirenb200_last_error(void);

int sirenb200_create(const sirenb200_config_t* cfg, sirenb200_handle_t* out);
int sirenb200_destroy(sirenb200_handle_t h);
int64_t sirenb200_workspace_bytes(sirenb200_handle_t h);

/* Input coordinates (data.py:78-88 get_grid + siren.py:125-128).  Either the two 1-D linspace tables
 * (the grid is their outer product and is generated in-kernel from the pixel index), or an explicit
 * [rows*width, 2] coordinate tensor (values in [0,1], (h, w) order) for arbitrary grids. */
int sirenb200_set_grid_lut(sirenb200_handle_t h, const float* lin_h, const float* lin_w);
int sirenb200_set_grid_coords(sirenb200_handle_t h, const float* coords);

/* FourierNet only: encoding.B (models/fourier.py:16-25), device fp32 [2, map_size / 2] row-major, read on every
 * forward (it is a frozen nn.Parameter of the model).  Parameter tables of a FourierNet handle are
 * [w0, b0, w1, b1, ...] of its depth - 1 nn.Linear layers. */
int sirenb200_set_fourier_encoding(sirenb200_handle_t h, const float* B);

/* Siren.forward (models/siren.py:123-134): pred[rows, width, out_features] fp32 in [0,1] space. */
int sirenb200_forward(sirenb200_handle_t h, const float* const* h_params, float* pred,
                      sirenb200_stream_t stream);

/* train_epoch's forward + F.mse_loss + backward (utils/train_helper.py:148-161).
 * grads: h_grads[i] receives d(loss*loss_scale)/d(param i) summed over this handle's rows and divided
 * by the FULL image element count (so that ranks can be summed); stats (device, 4 floats):
 *   [0] sum of squared error over this handle's rows, [1] loss (= [0] / (H*W*out)) valid when the handle
 *   covers the whole image, [2] 1.0 if any gradient is non-finite, [3] reserved. */
int sirenb200_forward_backward(sirenb200_handle_t h, const float* const* h_params, const float* img,
                               float loss_scale, float* const* h_grads, float* stats,
                               sirenb200_stream_t stream);

/* Backward for an arbitrary upstream gradient dpred[rows, width, out] (autograd of Siren.forward); uses
 * the activations stashed by the last sirenb200_forward(_backward) call on this handle. */
int sirenb200_backward(sirenb200_handle_t h, const float* const* h_params, const float* dpred,
                       float* const* h_grads, sirenb200_stream_t stream);

/* eval_epoch's metrics (utils/train_helper.py:48-57) from pred/img of n elements:
 * out (device, 2 floats) = [mse, mse of the (x*255).int() images]. */
int sirenb200_eval_metrics(const float* pred, const float* img, int64_t n, float* out,
                           sirenb200_stream_t stream);

/* torch.optim.Adam.step for a list of tensors (train_helper.py:72-78, stepped at :177 / core.py:687),
 * fused with GradScaler's unscale + inf check (compress.py:131-135), Masking.apply_mask (core.py:272-279)
 * and optim.zero_grad (train_helper.py:145).
 *   h_mask[i] may be NULL (no mask).  step = 1-based step count after increment.
 *   grads are multiplied by inv_scale before use; if skip_flag != NULL and *skip_flag != 0 (device float,
 *   e.g. stats[2]) the update is skipped (GradScaler.step semantics) but masks are still applied. */
int sirenb200_adam_step(int32_t n_tensors, float* const* h_params, float* const* h_grads,
                        float* const* h_exp_avg, float* const* h_exp_avg_sq,
                        const float* const* h_mask, const int64_t* h_numel, float lr, float beta1,
                        float beta2, float eps, int32_t step, float inv_scale, const float* skip_flag,
                        int32_t zero_grad, sirenb200_stream_t stream);

/* CUDA-graph friendly variants: every per-step quantity lives in device memory so that ONE captured fit step
 * can be replayed.  sched_state is 8 doubles: [0] step (0-based index of the next optimiser step), [1] base lr,
 * [2] StepLR gamma, [3] StepLR period, [4] beta1, [5] beta2, [6]/[7] outputs (step size, sqrt(1-beta2^t)).
 * sirenb200_sched_step computes [6],[7] for the step about to run (train_helper.py:80-84 StepLR + Adam bias
 * correction), stores the loss of the step that just ran into loss_ring[step % ring_len] (stats[1], or
 * stats[0]*inv_count when inv_count > 0, i.e. after an all-reduce of pixel shards) and, when loss_host is not
 * NULL, also into that host-mapped (pinned, device-accessible) float[2] as {loss, steps completed} (loss first,
 * system fence, then the count), so that train_epoch's "return loss.item()" is a host-side poll of the count
 * instead of a copy kernel + cudaMemcpy + stream synchronisation - and increments [0];
 * sirenb200_adam_step_dev is sirenb200_adam_step reading the schedule from sched_state. */
int sirenb200_sched_step(double* sched_state, const float* stats, float inv_count, float* loss_ring,
                         int32_t ring_len, float* loss_host, sirenb200_stream_t stream);
int sirenb200_adam_step_dev(int32_t n_tensors, float* const* h_params, float* const* h_grads,
                            float* const* h_exp_avg, float* const* h_exp_avg_sq,
                            const float* const* h_mask, const int64_t* h_numel, float beta1, float beta2,
                            float eps, const double* sched_state, float inv_scale, const float* skip_flag,
                            int32_t zero_grad, sirenb200_stream_t stream);

/* Masking.apply_mask for one tensor (pipeline/masking/core.py:272-279): w <- w * mask (bit-exact). */
int sirenb200_apply_mask(float* w, const float* mask, int64_t n, sirenb200_stream_t stream);

/* Global-magnitude prune threshold search (pipeline/masking/funcs/prune.py:54-104) on the device.
 * sorted_mags: the |w| of all masked layers, ascending (NaNs last), n elements; nonzero_total: weights currently
 * active; tokill: ceil(prune_rate * baseline_nonzero).  state (device, 2 doubles): [0] the persistent threshold
 * (in/out), [1] the increment (in).  result (device, 1 double): the number of weights the final threshold removes.
 * The threshold trajectory is the reference's, step for step (IEEE double arithmetic, unfused products, fp32
 * rounding of the threshold in every comparison). */
int sirenb200_prune_threshold_search(const float* sorted_mags, int64_t n, int64_t nonzero_total, int64_t tokill,
                                     double tolerance, double* state, double* result, sirenb200_stream_t stream);

/* KmeansQuant.find_centroids (pipeline/quant/kmeans.py:110-150 + kmeans_helper.py:59-116) on n weights.
 * init_centers: optional [2^bits - 1] initial guess (kmeans.py:123-129 builds it with torch.linspace on the
 * weights' device; NULL = linspace(min, max) of the non-zero weights evaluated in-kernel with CUDA
 * torch.linspace's symmetric formula).
 * Outputs: centroids[2^bits] (sorted by |c|, first n_centroids valid), n_centroids (device int32),
 * labels[n] int64, w_out[n] = centroids[labels].  Synchronises `stream` internally (convergence test). */
int sirenb200_kmeans_quantize(const float* w, int64_t n, int32_t bits, int32_t iter_limit, float tol,
                              const float* init_centers, float* centroids, int32_t* n_centroids, int64_t* labels, float* w_out,
                              sirenb200_stream_t stream);

/* QAT weight fake-quant (torch per-channel weight fake-quant via quant/context.py:35-47): per-output-row
 * symmetric int8, zero point 0:
 *   scale = max(-min(lo,0) / neg_div, max(hi,0) / pos_div, eps);  q = clamp(rne(w * (1/scale)), -128, 127)
 * row_min / row_max (device, [rows], the observer's running min / max) may both be NULL (use the row's own
 * extrema).  (neg_div, pos_div) = (127.5, 127.5) is torch 1.7's observer formula (the reference's pin);
 * (128, 127) is what torch >= 1.13's fused observer kernel computes. */
int sirenb200_fakequant_per_channel(const float* w, int32_t rows, int32_t cols, const float* row_min,
                                    const float* row_max, float neg_div, float pos_div, int8_t* codes,
                                    float* scales, float* w_out, sirenb200_stream_t stream);

/* QAT activation fake-quant (quant/context.py:35-47: torch.quantization.prepare_qat puts a
 * FusedMovingAvgObsFakeQuantize — per-tensor affine quint8 with reduce_range, i.e. [0, 127], moving-average min/max
 * observer with averaging constant 0.01 — on the OUTPUT of every nn.Linear, models/siren.py:62).  fp32 handles only.
 * state: device float[depth][4] = {running min, running max, scale, zero point} per layer, owned by the caller;
 * running min / max start at +inf / -inf.  training != 0: the observers are updated from the current batch before
 * quantising (model.train()); 0: frozen (model.eval(), and what the converted int8 model computes).  With it
 * enabled, sirenb200_forward / forward_backward quantise every layer's pre-activation and the backward applies the
 * straight-through mask.  enable = 0 switches it off. */
int sirenb200_set_act_quant(sirenb200_handle_t h, float* state, int32_t enable, int32_t training,
                            float averaging_const, int32_t qmin, int32_t qmax);
/* The same operator on one tensor (torch.fused_moving_avg_obs_fake_quant): out = fake_quant(x) (may alias x),
 * mask[n] (optional) = 1 where the value was not clipped, state updated as above. */
int sirenb200_fakequant_per_tensor(const float* x, int64_t n, float* state, int32_t training, float averaging_const,
                                   int32_t qmin, int32_t qmax, float* out, uint8_t* mask, sirenb200_stream_t stream);

/* Device-side packing for the entropy coder (pipeline/entropy_coding/__init__.py:15-41,70-120): assembles — or takes
 * apart — the byte stream the reference writes with NumpyParser / zstd, i.e. the tensors of
 * linear_state_dict(model.half()) back to back.  One kernel for all items (<= 96).  Per item i: kind
 *   0 fp32 -> fp16, 1 int64 -> uint8, 2 int64 -> uint16, 3 raw bytes            (pack:   h_src[i] -> stream + offset)
 *   4 fp16 -> fp32, 5 / 6 weight = fp16 code book [at stream + h_aux[i]] gathered by uint8 / uint16 codes
 *                                                                              (unpack: stream + offset -> h_dst[i])
 * count = elements (kind 3: bytes), offsets in bytes.  Host zstd / lzma stay on the host (out of scope). */
int sirenb200_pack_stream(int32_t n_items, const void* const* h_src, void* const* h_dst, const int32_t* h_kind,
                          const int64_t* h_count, const int64_t* h_offset, const int64_t* h_aux, uint8_t* stream_buf,
                          sirenb200_stream_t stream);

/* Optional per-kernel timing for bench.py's roofline: while enabled, tagged launches of this handle are
 * bracketed by cudaEvent pairs on the launching stream.  profile_read synchronises the recorded events and
 * returns, per kind, the summed device time (ms) and the launch count into HOST arrays of n_kinds entries.
 * Kinds: 0 weight staging, 1 first layer, 2 forward GEMM, 3 last layer + loss, 4 dX GEMM, 5 dW GEMM,
 *        6 layer-0 gradient, 7 partial reduction, 8 loss finalise. */
int sirenb200_profile_enable(sirenb200_handle_t h, int32_t enable);
int sirenb200_profile_read(sirenb200_handle_t h, float* h_total_ms, int32_t* h_count, int32_t n_kinds);

/* Developer aid: copies the fused-forward timeline capture (clock64 stamps of block 0; enabled by the
 * SIRENB200_TIMELINE environment variable at create time) to a HOST array of n int64. */
int sirenb200_debug_timeline(sirenb200_handle_t h, long long* h_out, int32_t n);

/* ---- k fit steps in one call (SURVEY.md §8b "fit_steps(k) fused driver") ---------------------------------
 * The loop of compress.py:137-143 for a host language that has no per-step work of its own: k times
 *   forward_backward (gradients into h_grads, stats)  ->  [comm_allreduce of the flat gradient buffer when
 *   comm != NULL]  ->  sched_step  ->  adam_step_dev (masks applied when h_mask != NULL),
 * all enqueued on `stream` with no host synchronisation; the per-step losses land in loss_ring (and
 * loss_host).  For a sharded fit the caller lays the gradients out as views of ONE flat buffer `flat`
 * (flat_n floats: [all gradients | stats[4]], stats = flat + flat_n - 4 rounded as the caller chose) and
 * passes inv_count = 1 / (H*W*C) of the FULL image; for a single-GPU fit comm = NULL, flat = NULL,
 * inv_count = 0.  Pointer tables are HOST arrays of device pointers (as everywhere in this ABI). */
typedef struct sirenb200_comm* sirenb200_comm_t;
typedef struct {
  int32_t n_tensors;
  float* const* h_params;
  float* const* h_grads;
  float* const* h_exp_avg;
  float* const* h_exp_avg_sq;
  const float* const* h_mask; /* NULL: no masks */
  const int64_t* h_numel;
  float beta1, beta2, eps;
  double* sched_state;        /* 8 doubles, see sirenb200_sched_step */
  float* stats;               /* device float[4] */
  float* loss_ring;           /* device, may be NULL */
  int32_t ring_len;
  float* loss_host;           /* host-mapped, may be NULL */
  sirenb200_comm_t comm;      /* NULL for a single-GPU fit */
  float* flat;                /* flat gradient buffer (sharded fits) */
  int64_t flat_n;
  float inv_count;            /* > 0: loss = stats[0] * inv_count (after the exchange) */
} sirenb200_fit_t;
int sirenb200_fit_steps(sirenb200_handle_t h, int32_t k, const float* img, const sirenb200_fit_t* fit,
                        sirenb200_stream_t stream);
/* One step of the above (what a CUDA-graph capture of a step records).  On a tensor-core handle everything after
 * the last GEMM is ONE kernel: it reduces the split-K / per-CTA gradient partials into h_grads, sums them over the
 * ranks through NVLink peer memory when fit->comm is set (block b of every rank exchanges its slice with block b of
 * every peer; results are bit-identical on all ranks), finalises the loss (stats[0] = sum of squared errors over the
 * whole image, stats[1] = loss, stats[2] = non-finite flag), advances the device-side schedule and applies Adam
 * (+ masks) — train_helper.py:151-184 from `loss.backward()` onwards.  SIRENB200_STEP_END=0 (environment) keeps
 * the separate kernels. */
int sirenb200_fit_step(sirenb200_handle_t h, const float* img, const sirenb200_fit_t* fit,
                       sirenb200_stream_t stream);

/* ---- gradient exchange for pixel-sharded fits (SURVEY.md §8e "allreduce_grads") ---------------------------
 * One process per GPU of ONE node.  The reference has no distributed code; this is the exchange step of the
 * row-sharded fit: an in-place SUM of a flat fp32 buffer [all dW | all db | sum_sq_err, -, nonfinite, -] over
 * the ranks, done by one kernel over NVLink peer memory (P2P loads of every rank's buffer, summed in rank
 * order so all ranks get bit-identical results).  Setup: every rank calls comm_create (allocates its
 * peer-visible region) and comm_handle (64-byte CUDA IPC handle); the caller exchanges the handles by any
 * means (torch.distributed all_gather in the Python mirror) and passes all of them, in rank order, to
 * comm_connect.  comm_allreduce is asynchronous on `stream`, takes no per-step host state (epochs live on the
 * device) and can be captured in a CUDA graph; `data` must be 16-byte aligned with room for n rounded up to a
 * multiple of 4 floats, n <= max_floats.  A peer that never arrives makes the kernel trap after a few
 * seconds (sticky CUDA error) rather than hang. */
int sirenb200_comm_create(int32_t rank, int32_t world, int64_t max_floats, sirenb200_comm_t* out);
int sirenb200_comm_handle(sirenb200_comm_t c, void* handle_out_64_bytes);
int sirenb200_comm_connect(sirenb200_comm_t c, const void* handles_world_x_64_bytes);
int sirenb200_comm_allreduce(sirenb200_comm_t c, float* data, int64_t n, sirenb200_stream_t stream);
int sirenb200_comm_destroy(sirenb200_comm_t c);

/* Number of kernels this library has launched since load (bench.py's gpu_launches). */
int64_t sirenb200_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* SIREN_B200_H_ */
