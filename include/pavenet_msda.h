/*
 * pavenet_msda.h — C ABI of the B200-native multi-scale deformable attention
 * sampling op (forward + backward) used by PAVE-Net's spatial encoder and
 * pose-aware decoder attention.
 *
 * This is the drop-in boundary.  It replaces, entry point for entry point,
 * the two functions the reference binds into `mmcv._ext`:
 *
 *   ms_deform_attn_forward   third_party/mmcv/mmcv/ops/csrc/pytorch/ms_deform_attn.cpp:38-46
 *                            (pybind: .../csrc/pytorch/pybind.cpp:737-742,
 *                             CUDA host code: .../csrc/pytorch/cuda/ms_deform_attn_cuda.cu:209-277)
 *   ms_deform_attn_backward  third_party/mmcv/mmcv/ops/csrc/pytorch/ms_deform_attn.cpp:48-60
 *                            (pybind: .../csrc/pytorch/pybind.cpp:743-748,
 *                             CUDA host code: .../csrc/pytorch/cuda/ms_deform_attn_cuda.cu:279-351)
 *
 * which `MultiScaleDeformableAttnFunction` calls
 * (third_party/mmcv/mmcv/ops/multi_scale_deform_attn.py:47-53, 76-86).
 *
 * Conventions
 *  - Plain pointers and sizes only; no torch / ATen types.
 *  - Every `d_` pointer is a DEVICE pointer on the current CUDA device,
 *    including `d_spatial_shapes` and `d_level_start_index` (the reference
 *    reads both inside its kernels, ms_deform_attn_cuda_kernel.cuh:226-229,
 *    so no host round-trip is introduced here either).
 *  - All tensors are dense, row-major ("contiguous"), with the reference's
 *    layouts:
 *        value               (B, S, M, D)
 *        spatial_shapes      (L, 2)  int64, (H_l, W_l)
 *        level_start_index   (L,)    int64
 *        sampling_locations  (B, Q, M, L, P, 2)   normalised (x, y)
 *        attention_weights   (B, Q, M, L, P)
 *        output / grad_output (B, Q, M*D)
 *  - The caller owns every buffer.  `msda_forward` fully overwrites `d_output`
 *    (no pre-zeroing needed; the reference zero-fills it itself,
 *    ms_deform_attn_cuda.cu:247-248).  `msda_backward` ACCUMULATES into
 *    `d_grad_value` (which the caller must have zeroed, as the reference's
 *    autograd wrapper does, multi_scale_deform_attn.py:72-74) and fully
 *    overwrites `d_grad_sampling_loc` / `d_grad_attn_weight`.
 *  - `stream` is a `cudaStream_t` passed as `void*` (NULL = legacy default
 *    stream).  Calls are asynchronous: they enqueue work and return.
 *  - Every function returns MSDA_OK (0) or a negative MSDA_ERR_* code;
 *    `msda_last_error()` returns a thread-local human-readable message.  The
 *    reference swallows launch errors with printf
 *    (ms_deform_attn_cuda.cu:41-44); this library reports them.
 *  - Re-entrant, no global mutable state except the launch counter.
 *
 * dtype codes: `dtype` is the type of locations, weights, output and their
 * gradients (MSDA_F32 or MSDA_F64, the two types the reference instantiates,
 * ms_deform_attn_cuda.cu:258,329).  `value_dtype` / `grad_value_dtype` may
 * additionally be MSDA_BF16 when `dtype` is MSDA_F32 (bf16 value storage is
 * a capability the reference does not have).
 */
#ifndef PAVENET_MSDA_H_
#define PAVENET_MSDA_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MSDA_ABI_VERSION 2

enum msda_dtype { MSDA_F32 = 0, MSDA_F64 = 1, MSDA_BF16 = 2 };

enum msda_status {
  MSDA_OK = 0,
  MSDA_ERR_INVALID_ARGUMENT = -1, /* NULL pointer, non-positive size, bad dtype combination */
  MSDA_ERR_UNSUPPORTED = -2,      /* valid request this build has no kernel for */
  MSDA_ERR_CUDA = -3,             /* CUDA runtime / launch failure; see msda_last_error() */
  MSDA_ERR_NO_DEVICE = -4         /* no usable CUDA device */
};

/* ABI version of the loaded library (== MSDA_ABI_VERSION it was built with). */
int msda_abi_version(void);

/* Thread-local message for the last non-OK status returned on this thread. */
const char *msda_last_error(void);

/* Number of CUDA kernels this library has launched since load (all threads). */
uint64_t msda_launch_count(void);

/* Kernel-selection knobs for tests and benchmarks (not thread safe; set between calls).
 * Also read once from the environment at load: PAVENET_MSDA_<NAME IN CAPITALS>.
 *   "force_generic" 0/1      any-D scalar kernels instead of the vector ones
 *   "fwd_split", "bwd_split" lane groups per row of the rows kernels (0 = heuristic)
 *   "flat"                   small-Q flat kernels: 0 never, 1 heuristic (default), 2 always
 *   "fwd_variant", "bwd_variant"  large-Q kernel variants (0 = default)
 * Returns MSDA_ERR_INVALID_ARGUMENT for an unknown name. */
int msda_set_option(const char *name, int value);

/* Launches of ONE kernel family since load, so a test can assert that the kernel it means
 * to check is the one that ran.  Families: 0 generic forward, 1 generic backward, 2 rows
 * forward, 3 rows backward, 4 / 5 the same with the fused prologue, 6 flat (small-Q)
 * forward, 7 flat backward, 8 / 9 the same fused, 10 tcgen05 linear, 11 its weight
 * gradient, 12 column sums, 13 LayerNorm, 14 tile-staged (shared-memory) forward.  Unknown codes: 0. */
uint64_t msda_launch_count_family(int family);

/* Name of the kernel family the dispatcher would use for this problem
 * ("rows<D=32,f32>", "generic", ...).  Static storage; never NULL. */
const char *msda_kernel_name(int channels, int dtype, int value_dtype);

/*
 * Forward.  Replaces ms_deform_attn_forward (ms_deform_attn.cpp:38-46).
 * out[b,q,m*D+c] = sum_{l,p} w[b,q,m,l,p] * bilinear(value_l[b,:,m,c]; x*W_l-0.5, y*H_l-0.5)
 * with zero padding outside the map (ms_deform_attn_cuda_kernel.cuh:17-64,200-254).
 * `im2col_step` of the reference is a batching knob of its host loop and has
 * no effect on results; it is validated by the Python wrapper, not here.
 */
int msda_forward(const void *d_value, const int64_t *d_spatial_shapes,
                 const int64_t *d_level_start_index,
                 const void *d_sampling_loc, const void *d_attn_weight,
                 void *d_output, int batch, int spatial_size, int num_heads,
                 int channels, int num_levels, int num_query, int num_point,
                 int dtype, int value_dtype, void *stream);

/*
 * msda_forward that also zero-fills `d_clear` (clear_bytes bytes; both zero = plain
 * msda_forward) on the same stream: the grad_value buffer the coming msda_backward will
 * accumulate into (the reference zero-fills it in its autograd wrapper,
 * multi_scale_deform_attn.py:72).  For the small-Q shapes (pose decoder) the fill is folded
 * into the persistent forward kernel, whose warps leave most of the DRAM write bandwidth
 * idle; otherwise it is a memset ahead of the kernel.  `d_clear` must not alias an input.
 */
int msda_forward_clear(const void *d_value, const int64_t *d_spatial_shapes,
                       const int64_t *d_level_start_index,
                       const void *d_sampling_loc, const void *d_attn_weight,
                       void *d_output, int batch, int spatial_size, int num_heads,
                       int channels, int num_levels, int num_query, int num_point,
                       int dtype, int value_dtype, void *d_clear,
                       size_t clear_bytes, void *stream);

/*
 * Backward.  Replaces ms_deform_attn_backward (ms_deform_attn.cpp:48-60).
 * grad_value += scatter of bilinear weights * grad_output * attention weight,
 * grad_sampling_loc / grad_attn_weight as in ms_deform_attn_cuda_kernel.cuh:66-131.
 * `d_grad_value` is accumulated into and must be zero-initialised by the
 * caller; the other two gradients are fully overwritten.
 */
int msda_backward(const void *d_value, const int64_t *d_spatial_shapes,
                  const int64_t *d_level_start_index,
                  const void *d_sampling_loc, const void *d_attn_weight,
                  const void *d_grad_output, void *d_grad_value,
                  void *d_grad_sampling_loc, void *d_grad_attn_weight,
                  int batch, int spatial_size, int num_heads, int channels,
                  int num_levels, int num_query, int num_point, int dtype,
                  int value_dtype, int grad_value_dtype, void *stream);

/*
 * Fused variants (no counterpart in the reference: they absorb the elementwise
 * chain its modules run around the op, multi_scale_deform_attn.py:373-393,
 * opera/models/utils/transformer.py:390-412).  Inputs are the RAW projections:
 *   d_offsets    (B,Q,M,L,P,2)  sampling_offsets Linear output
 *   d_logits     (B,Q,M,L*P)    attention_weights Linear output (pre-softmax)
 *   d_ref_points (B,Q,L,R,2)    reference points, R = ref_points_per_level = 1 or P
 *   d_scale      (B,Q,L,2)      offset scale, or NULL for 1/(W_l, H_l)
 * and the kernels compute  loc = ref + off * scale  and  w = softmax(logits)
 * over L*P themselves.  fp32 only; channels must be 16, 32 or 64 (MSDA_ERR_UNSUPPORTED
 * otherwise — callers then fall back to msda_forward / msda_backward).
 * d_softmax_stats (B,Q,M,2) is written by the forward (row max, 1/sum) and
 * read by the backward, which also takes the forward's d_output: the softmax
 * backward's row term sum_t w_t dL/dw_t equals <grad_output[row], output[row]>,
 * so no reduction over the samples is needed.  The backward accumulates into the
 * zero-initialised d_grad_value and overwrites d_grad_offsets / d_grad_logits
 * (softmax and scale already differentiated through); d_grad_loc, if not NULL,
 * receives d/d(loc) for callers that need reference-point gradients.
 * d_clear / clear_bytes of the forward: as in msda_forward_clear (NULL, 0 = none).
 */
int msda_fused_forward(const void *d_value, const int64_t *d_spatial_shapes,
                       const int64_t *d_level_start_index, const float *d_offsets,
                       const float *d_logits, const float *d_ref_points,
                       const float *d_scale, float *d_output,
                       float *d_softmax_stats, int batch, int spatial_size,
                       int num_heads, int channels, int num_levels,
                       int num_query, int num_point, int ref_points_per_level,
                       int value_dtype, void *d_clear, size_t clear_bytes,
                       void *stream);

int msda_fused_backward(const void *d_value, const int64_t *d_spatial_shapes,
                        const int64_t *d_level_start_index,
                        const float *d_offsets, const float *d_logits,
                        const float *d_ref_points, const float *d_scale,
                        const float *d_softmax_stats, const float *d_output,
                        const float *d_grad_output,
                        float *d_grad_value, float *d_grad_offsets,
                        float *d_grad_logits, float *d_grad_loc, int batch,
                        int spatial_size, int num_heads, int channels,
                        int num_levels, int num_query, int num_point,
                        int ref_points_per_level, int value_dtype, void *stream);

/*
 * The fp32 Linear layers next to the op (SURVEY.md section 8f ranks 2 and 4:
 * `value_proj` + padding mask + storage dtype, `output_proj` + dropout +
 * residual, `sampling_offsets`, `attention_weights`,
 * multi_scale_deform_attn.py:369-377,406-412, opera/models/utils/transformer.py:1706-1720;
 * and the 256 <-> 1024 feed-forward pair of the transformer layers that carry
 * them, mmcv/cnn/bricks/transformer.py FFN) on the tcgen05 tensor cores:
 *   y[rows,out] = epilogue(x[rows,in] * weight^T)        weight (out, in), fp32
 * in_features and out_features each one of 128, 256, 1024 (128x128 and 1024x1024
 * excluded); anything else returns MSDA_ERR_UNSUPPORTED.  Computed as a 3xTF32
 * split with fp32 accumulation in tensor memory (fp32-level accuracy; plain TF32
 * would not meet the op's parity bound).  The epilogue applies, in this order:
 *   + d_bias (may be NULL);
 *   relu != 0: max(v, 0);
 *   d_gate != NULL ([rows,out] fp32): v = gate > 0 ? v * gate_scale : 0  (the
 *     backward of ReLU followed by dropout, with gate = the saved activation);
 *   dropout_p > 0: v = keep ? v / (1 - p) : 0, keep decided by a counter-based hash
 *     of (element index, seed) -- msda_dropout_backward regenerates it; seed =
 *     dropout_seed, plus *d_dropout_seed when that device pointer is not NULL (a seed
 *     in device memory lets a captured CUDA graph draw a new mask at every replay);
 *   mask_mode 1: rows with d_row_mask[r] != 0 are written as zeros (mask after the
 *     projection); 2: they are written as the bias (input masked before it);
 *   + d_residual ([rows,out] fp32, may be NULL);
 * and stores out_dtype MSDA_F32 or MSDA_BF16.  d_scratch: 2*in*out floats of
 * device scratch (the split weight).
 */
int msda_linear_fused(const float *d_x, const float *d_weight, const float *d_bias,
                      const uint8_t *d_row_mask, int mask_mode, int relu,
                      const float *d_gate, float gate_scale, float dropout_p,
                      uint64_t dropout_seed, const uint64_t *d_dropout_seed,
                      const float *d_residual, void *d_y,
                      int rows, int in_features, int out_features, int out_dtype,
                      float *d_scratch, void *stream);

/* msda_linear_fused with bias and mask only (no ReLU, gate, dropout, residual). */
int msda_linear256(const float *d_x, const float *d_weight, const float *d_bias,
                   const uint8_t *d_row_mask, int mask_mode, void *d_y, int rows,
                   int in_features, int out_features, int out_dtype,
                   float *d_scratch, void *stream);

/* Weight gradient of the same layer: grad_weight[out][in] = grad_y^T x
 * over `rows` rows (overwritten), split-K over the rows on tcgen05 with 3xTF32
 * and a vector-reduction epilogue.  mask_mode as in msda_linear256: 1 drops
 * masked rows of grad_y, 2 drops masked rows of x. */
int msda_linear256_wgrad(const float *d_grad_y, const float *d_x,
                         const uint8_t *d_row_mask, int mask_mode,
                         float *d_grad_weight, int rows, int in_features,
                         int out_features, void *stream);

/* Bias gradient of the same layer: grad_bias[width] = column sums of
 * grad_y[rows][width] (overwritten), width 128, 256 or 1024; rows with
 * d_row_mask[r] != 0 are skipped (pass NULL to sum every row). */
int msda_colsum256(const float *d_grad_y, const uint8_t *d_row_mask,
                   float *d_grad_bias, int rows, int width, void *stream);

/* Backward of msda_linear_fused's dropout: grad_out = grad_y * keep / (1 - p) with
 * the keep decisions regenerated from (dropout_p, dropout_seed, d_dropout_seed) as
 * given to the forward call; when d_grad_bias
 * is not NULL it also receives the column sums of grad_out (the bias gradient),
 * from the same pass. */
int msda_dropout_backward(const float *d_grad_y, float *d_grad_out,
                          float *d_grad_bias, int rows, int width, float dropout_p,
                          uint64_t dropout_seed, const uint64_t *d_dropout_seed,
                          void *stream);

/*
 * LayerNorm over the channel dimension, the `norm` step after every attention
 * module and feed-forward block of the reference's transformer layers
 * (operation_order ('self_attn', 'norm', 'ffn', 'norm'); mmcv builds nn.LayerNorm(256)).
 * width must be 256.  Forward writes y and the per-row mean / 1/sqrt(var + eps) the
 * backward needs; backward makes one pass over (x, grad_y) and overwrites grad_x,
 * grad_gamma[width] and grad_beta[width].  Biased variance, as torch.
 */
int msda_layernorm_forward(const float *d_x, const float *d_gamma, const float *d_beta,
                           float *d_y, float *d_mean, float *d_rstd, int rows,
                           int width, float eps, void *stream);

int msda_layernorm_backward(const float *d_x, const float *d_grad_y,
                            const float *d_gamma, const float *d_mean,
                            const float *d_rstd, float *d_grad_x, float *d_grad_gamma,
                            float *d_grad_beta, int rows, int width, void *stream);

/*
 * Host-buffer convenience entry points (what a cgo / JNI / ctypes caller
 * without its own device memory management binds).  All pointers are HOST
 * pointers (pinned memory gives asynchronous copies; pageable memory works
 * but is slower).  Each call stages inputs host->device, runs the kernels and
 * copies results device->host on an internal stream, and returns after the
 * results are in the host buffers.  Device scratch is cached per workspace.
 */
typedef struct msda_workspace msda_workspace;

int msda_workspace_create(msda_workspace **out_ws);
void msda_workspace_destroy(msda_workspace *ws);
/* Upload bytes per pipeline piece of the staged calls (defaults: 12 MiB for a
 * blocking call, 64 MiB for a queued one).  A size of at least the call's total
 * upload selects the monolithic form: every tensor moves as one copy and the
 * kernels run once over the whole batch -- nothing overlaps inside the call, which
 * suits callers that keep three or more queued calls in flight. */
int msda_workspace_set_piece_bytes(msda_workspace *ws, size_t bytes);

/* Page-locked host memory for callers without their own CUDA binding.  With
 * pinned buffers the staged calls below overlap the upload of the next piece,
 * the kernels of the current one and the download of the previous one (the
 * link is full duplex); with pageable memory they still work, serialised. */
void *msda_host_alloc(size_t bytes);
void msda_host_free(void *ptr);

int msda_forward_host(msda_workspace *ws, const void *h_value,
                      const int64_t *h_spatial_shapes,
                      const int64_t *h_level_start_index,
                      const void *h_sampling_loc, const void *h_attn_weight,
                      void *h_output, int batch, int spatial_size,
                      int num_heads, int channels, int num_levels,
                      int num_query, int num_point, int dtype, int value_dtype);

/* Forward + backward in one staged, pipelined call: inputs and grad_output go
 * up once, output and the three gradients come back; work is cut into
 * (batch entry, query chunk) pieces that flow through three streams.
 * `h_output` may be NULL (backward only).  grad_value is returned in `dtype`. */
int msda_forward_backward_host(
    msda_workspace *ws, const void *h_value, const int64_t *h_spatial_shapes,
    const int64_t *h_level_start_index, const void *h_sampling_loc,
    const void *h_attn_weight, const void *h_grad_output, void *h_output,
    void *h_grad_value, void *h_grad_sampling_loc, void *h_grad_attn_weight,
    int batch, int spatial_size, int num_heads, int channels, int num_levels,
    int num_query, int num_point, int dtype, int value_dtype);

/* The same call, queued only: it returns as soon as every copy and kernel is
 * enqueued on the workspace's streams; the host buffers must stay untouched
 * (inputs) / unread (results) until msda_workspace_wait(ws) has returned.  One
 * call may be in flight per workspace; a caller that alternates between two
 * workspaces (and two sets of result buffers) keeps the link busy across calls:
 * the next call's first upload runs under the previous call's last download.
 * (Additive to ABI version 2.) */
int msda_forward_backward_host_async(
    msda_workspace *ws, const void *h_value, const int64_t *h_spatial_shapes,
    const int64_t *h_level_start_index, const void *h_sampling_loc,
    const void *h_attn_weight, const void *h_grad_output, void *h_output,
    void *h_grad_value, void *h_grad_sampling_loc, void *h_grad_attn_weight,
    int batch, int spatial_size, int num_heads, int channels, int num_levels,
    int num_query, int num_point, int dtype, int value_dtype);
/* Blocks until the call queued by msda_forward_backward_host_async is complete
 * (no-op when nothing is in flight). */
int msda_workspace_wait(msda_workspace *ws);

#ifdef __cplusplus
}
#endif
#endif /* PAVENET_MSDA_H_ */
