// msda_kernels.h — internal interface between the C ABI (msda_capi.cu) and
// the kernel translation units.  Not installed; the public contract is
// include/pavenet_msda.h.
#pragma once

#include "../../include/pavenet_msda.h"
#include "msda_common.cuh"
#include "msda_bwd_io.cuh"

namespace msda {

// Tuning knobs, settable through environment variables read once at load
// (PAVENET_MSDA_FORCE_GENERIC=1, PAVENET_MSDA_BWD_SPLIT=n) — used by tests
// to pin a kernel family and by the bench to sweep.
struct Tuning {
  int force_generic = 0;
  int fwd_split = 0;  // 0 = heuristic
  int bwd_split = 0;  // 0 = heuristic
  int linear_bk = 16; // linear256 K-chunk: 16 (two CTAs per SM) or 32 (one)
  int linear_bm = 0;  // rows per CTA: 128 (default, also 0) or 256
  int copy_streams = 1;  // host-buffer entry points: upload / download streams per direction (1..4)
  int flat = 1;          // flat small-Q kernels (msda_flat.cu): 0 never, 1 heuristic, 2 whenever legal
  int flat_fwd_cfg = 0;    // tuning sweep of the flat kernels (batch, blocks per SM), 0 = default
  int flat_bwd_cfg = 0;
  int clear_policy = 2;    // stores of the folded zero-fill: 0 streaming (evict-first), 1 default, 2 L2 evict-last (measured best:
                           // the backward's first reductions then hit resident zero lines; pose cfg3 step 0.149 -> 0.145 ms)
  int flat_order = 1;      // flat kernels: 1 = every warp walks its piece from chunk 0 upwards (frames in phase), 0 = in storage order
  int clear_mode = 0;      // msda_forward_clear: 0 fold into the flat kernel / memset ahead of the others,
                           // 1 memset on a side stream concurrent with the forward kernel,
                           // 2 flat kernel: TMA bulk stores of a zeroed shared-memory tile
  int l2_prefetch = 0;     // flat kernels stream value into L2 first: bit 0 forward, bit 1 backward
  int l2_prefetch_mb = 120;  // ... when value is at most this many MiB
  int agg_tile_kb = 36;    // bwd_variant 3: shared-memory tile for the privatised coarsest level (KiB per block)
  int agg_min_level = 0;   // bwd_variant 2: aggregate from this level on (0 = every level)
  int bwd_variant = 0;   // large-Q backward: 0 default, 1 plain rows, 2 warp-aggregated, 3 tile
  int fwd_variant = 0;   // large-Q forward: 0 default, 2 head-affine, 3 / 4 256-bit loads, 5 shared-memory tiles
};
const Tuning& tuning();

bool rows_supported(int D, int value_dtype);
int choose_split(const Dims& d, int G, int sm_count);

// Kernel families, for msda_launch_count_family(): tests assert that the kernel they
// mean to check is the one that ran.
enum KernelFamily {
  KF_FWD_GENERIC = 0, KF_BWD_GENERIC, KF_FWD_ROWS, KF_BWD_ROWS, KF_FWD_ROWS_FUSED,
  KF_BWD_ROWS_FUSED, KF_FWD_FLAT, KF_BWD_FLAT, KF_FWD_FLAT_FUSED, KF_BWD_FLAT_FUSED,
  KF_LINEAR, KF_LINEAR_WGRAD, KF_COLSUM, KF_LAYERNORM, KF_FWD_TILE, KF_COUNT
};
void note_kernel(int family);

// `clear` / `clear_bytes`: optional buffer (the coming backward's grad_value) the forward
// zero-fills on the same stream — inside the kernel for the flat family, as a memset otherwise.
cudaError_t launch_forward(const void* value, const int64_t* shapes, const int64_t* lsi,
                           const void* loc, const void* aw, void* out, const Dims& d, int dtype,
                           int value_dtype, int sm_count, int force_generic, void* clear,
                           size_t clear_bytes, cudaStream_t st);

// tile-staged encoder forward (msda_fwd_tile.cu)
bool tile_forward_eligible(const Dims& d, int dtype, int value_dtype);
cudaError_t launch_forward_tile(const void* value, const int64_t* shapes, const int64_t* lsi,
                                const void* loc, const void* aw, void* out, const Dims& d,
                                int sm_count, cudaStream_t st);

// flat small-Q family (msda_flat.cu)
bool flat_preferred(const Dims& d, int G, int sm_count);
cudaError_t launch_forward_flat(const void* value, const int64_t* shapes, const int64_t* lsi,
                                const PlainSource& src, float* out, const Dims& d, int value_dtype,
                                int sm_count, void* clear, size_t clear_bytes, cudaStream_t st);
cudaError_t launch_forward_flat_fused(const void* value, const int64_t* shapes, const int64_t* lsi,
                                      const FusedSource& src, float* out, const Dims& d,
                                      int value_dtype, int sm_count, void* clear,
                                      size_t clear_bytes, cudaStream_t st);
cudaError_t launch_backward_flat(const void* value, const int64_t* shapes, const int64_t* lsi,
                                 const PlainIO& io, const float* grad_out, void* grad_value,
                                 const Dims& d, int value_dtype, int grad_value_dtype, int sm_count,
                                 cudaStream_t st);
// variant 3 of the large-Q backward (msda_bwd_priv.cu): coarsest level privatised in shared memory
cudaError_t launch_backward_priv(const void* value, const int64_t* shapes, const int64_t* lsi,
                                 const PlainIO& io, const float* grad_out, void* grad_value,
                                 const Dims& d, int sm_count, cudaStream_t st);
cudaError_t launch_backward_flat_fused(const void* value, const int64_t* shapes,
                                       const int64_t* lsi, const FusedIO& io,
                                       const float* grad_out, float* grad_value, const Dims& d,
                                       int value_dtype, int sm_count, cudaStream_t st);

cudaError_t launch_backward(const void* value, const int64_t* shapes, const int64_t* lsi,
                            const void* loc, const void* aw, const void* grad_out,
                            void* grad_value, void* grad_loc, void* grad_aw, const Dims& d,
                            int dtype, int value_dtype, int grad_value_dtype, int sm_count,
                            int force_generic, cudaStream_t st);

cudaError_t launch_forward_fused(const void* value, const int64_t* shapes, const int64_t* lsi,
                                 const FusedSource& src, float* out, const Dims& d, int value_dtype,
                                 int sm_count, void* clear, size_t clear_bytes, cudaStream_t st);

cudaError_t launch_backward_fused(const void* value, const int64_t* shapes, const int64_t* lsi,
                                  const FusedSource& src, const float* out, const float* grad_out,
                                  float* grad_value, float* grad_off, float* grad_logit,
                                  float* grad_loc, const Dims& d, int value_dtype, int sm_count,
                                  cudaStream_t st);

// Y = epilogue(X W^T), dW = dY^T X and the bias gradient for the Linear layers of the
// attention modules and transformer layers on tcgen05 (3xTF32), linear256_tc.cu.
// What the GEMM epilogue does to the accumulator, in this order:
struct LinearEpilogue {
  const float* bias = nullptr;        // [out] added first
  const uint8_t* row_mask = nullptr;  // [rows]; mask_mode 1: masked rows = 0, 2: masked rows = bias
  int mask_mode = 0;
  int relu = 0;                       // max(v, 0)
  const float* gate = nullptr;        // [rows, out]: v = gate > 0 ? v * gate_scale : 0
  float gate_scale = 1.f;
  uint32_t dropout_threshold = 0;     // keep iff hash(index, seed) >= threshold; 0 = no dropout
  float dropout_scale = 1.f;          // 1 / (1 - p)
  uint32_t seed_lo = 0, seed_hi = 0;  // the 64-bit seed ...
  const unsigned long long* seed_ptr = nullptr;  // ... plus *seed_ptr when set (device memory: lets a
                                                 // captured CUDA graph draw new masks at every replay)
  const float* residual = nullptr;    // [rows, out] added last
};
bool linear_shape_supported(int in_features, int out_features);
cudaError_t launch_linear256(const float* x, const float* w, const LinearEpilogue& ep, void* y, int rows,
                             int in_features, int out_features, int out_dtype, float* scratch,
                             cudaStream_t st);
cudaError_t launch_linear256_wgrad(const float* dy, const float* x, const uint8_t* row_mask,
                                   int mask_mode, float* dw, int rows, int in_features, int out_features,
                                   int sm_count, cudaStream_t st);
cudaError_t launch_colsum256(const float* dy, const uint8_t* row_mask, float* out, float* dy_out,
                             uint32_t threshold, float scale, uint32_t seed_lo, uint32_t seed_hi,
                             const unsigned long long* seed_ptr, int rows, int width, int sm_count,
                             cudaStream_t st);

// LayerNorm over 256 channels, layernorm.cu
bool layernorm_width_supported(int width);
cudaError_t launch_layernorm_forward(const float* x, const float* gamma, const float* beta, float* y,
                                     float* mean, float* rstd, int rows, float eps, int sm_count,
                                     cudaStream_t st);
cudaError_t launch_layernorm_backward(const float* x, const float* dy, const float* gamma, const float* mean,
                                      const float* rstd, float* dx, float* dgamma, float* dbeta, int rows,
                                      int sm_count, cudaStream_t st);

// adds n to the library-wide launch counter (msda_launch_count)
void note_launches(int n);

}  // namespace msda
