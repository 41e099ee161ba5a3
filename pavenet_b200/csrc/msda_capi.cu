// msda_capi.cu — the extern "C" boundary declared in include/pavenet_msda.h.
// Argument validation, dtype dispatch, error reporting, and the host-buffer
// convenience entry points.  No torch / ATen here: plain pointers and sizes.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "msda_kernels.h"

namespace msda {

static std::atomic<uint64_t> g_launches{0};
void note_launches(int n) { g_launches.fetch_add(static_cast<uint64_t>(n), std::memory_order_relaxed); }
static std::atomic<uint64_t> g_family[KF_COUNT];
void note_kernel(int family) {
  if (family >= 0 && family < KF_COUNT) g_family[family].fetch_add(1, std::memory_order_relaxed);
}

static thread_local char t_err[512] = "";

static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_err, sizeof(t_err), fmt, ap);
  va_end(ap);
  return code;
}

static int env_int(const char* name, int dflt) {
  const char* v = std::getenv(name);
  return (v && *v) ? std::atoi(v) : dflt;
}

static Tuning& tuning_mut() {
  static Tuning t = [] {
    Tuning x;
    x.force_generic = env_int("PAVENET_MSDA_FORCE_GENERIC", 0);
    x.fwd_split = env_int("PAVENET_MSDA_FWD_SPLIT", 0);
    x.bwd_split = env_int("PAVENET_MSDA_BWD_SPLIT", 0);
    x.linear_bk = env_int("PAVENET_MSDA_LINEAR_BK", 16) == 32 ? 32 : 16;
    x.linear_bm = env_int("PAVENET_MSDA_LINEAR_BM", 0);
    if (x.linear_bm != 128 && x.linear_bm != 256) x.linear_bm = 0;
    x.copy_streams = env_int("PAVENET_MSDA_COPY_STREAMS", 1);
    if (x.copy_streams < 1) x.copy_streams = 1;
    if (x.copy_streams > 4) x.copy_streams = 4;
    x.flat = env_int("PAVENET_MSDA_FLAT", 1);
    x.clear_mode = env_int("PAVENET_MSDA_CLEAR_MODE", x.clear_mode);
    x.flat_order = env_int("PAVENET_MSDA_FLAT_ORDER", x.flat_order);
    x.l2_prefetch = env_int("PAVENET_MSDA_L2_PREFETCH", x.l2_prefetch);
    x.l2_prefetch_mb = env_int("PAVENET_MSDA_L2_PREFETCH_MB", x.l2_prefetch_mb);
    x.bwd_variant = env_int("PAVENET_MSDA_BWD_VARIANT", 0);
    x.fwd_variant = env_int("PAVENET_MSDA_FWD_VARIANT", 0);
    return x;
  }();
  return t;
}
const Tuning& tuning() { return tuning_mut(); }

// SM count of the current device, cached per device ordinal.
static int current_sm_count(int* out) {
  static std::mutex mu;
  static int cache[64];
  static bool have[64];
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return fail(MSDA_ERR_NO_DEVICE, "cudaGetDevice: %s", cudaGetErrorString(e));
  std::lock_guard<std::mutex> lk(mu);
  if (dev >= 0 && dev < 64 && have[dev]) {
    *out = cache[dev];
    return MSDA_OK;
  }
  int n = 0;
  e = cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess)
    return fail(MSDA_ERR_NO_DEVICE, "cudaDeviceGetAttribute: %s", cudaGetErrorString(e));
  if (dev >= 0 && dev < 64) {
    cache[dev] = n;
    have[dev] = true;
  }
  *out = n;
  return MSDA_OK;
}

static size_t dtype_size(int dt) {
  switch (dt) {
    case MSDA_F32: return 4;
    case MSDA_F64: return 8;
    case MSDA_BF16: return 2;
    default: return 0;
  }
}

static int check_dims(int batch, int spatial_size, int num_heads, int channels, int num_levels,
                      int num_query, int num_point, Dims* d) {
  if (batch <= 0 || spatial_size <= 0 || num_heads <= 0 || channels <= 0 || num_levels <= 0 ||
      num_query <= 0 || num_point <= 0)
    return fail(MSDA_ERR_INVALID_ARGUMENT,
                "all sizes must be positive (batch=%d spatial_size=%d num_heads=%d channels=%d "
                "num_levels=%d num_query=%d num_point=%d)",
                batch, spatial_size, num_heads, channels, num_levels, num_query, num_point);
  // in-kernel offsets inside one batch entry are int32
  if (static_cast<int64_t>(spatial_size) * num_heads * channels >= (int64_t(1) << 31))
    return fail(MSDA_ERR_UNSUPPORTED, "spatial_size*num_heads*channels must be < 2^31");
  if (static_cast<int64_t>(num_levels) * num_point >= (int64_t(1) << 24))
    return fail(MSDA_ERR_UNSUPPORTED, "num_levels*num_point must be < 2^24");
  d->B = batch; d->S = spatial_size; d->M = num_heads; d->D = channels;
  d->L = num_levels; d->Q = num_query; d->P = num_point;
  return MSDA_OK;
}

static int check_dtypes(int dtype, int value_dtype, int grad_value_dtype) {
  if (dtype != MSDA_F32 && dtype != MSDA_F64)
    return fail(MSDA_ERR_INVALID_ARGUMENT, "dtype must be MSDA_F32 or MSDA_F64, got %d", dtype);
  const bool v_ok = value_dtype == dtype || (dtype == MSDA_F32 && value_dtype == MSDA_BF16);
  if (!v_ok)
    return fail(MSDA_ERR_INVALID_ARGUMENT,
                "value_dtype %d incompatible with dtype %d (same type, or bf16 with f32)",
                value_dtype, dtype);
  if (grad_value_dtype >= 0) {
    const bool g_ok = grad_value_dtype == dtype ||
                      (value_dtype == MSDA_BF16 && grad_value_dtype == MSDA_BF16);
    if (!g_ok)
      return fail(MSDA_ERR_INVALID_ARGUMENT,
                  "grad_value_dtype %d incompatible with dtype %d / value_dtype %d",
                  grad_value_dtype, dtype, value_dtype);
  }
  return MSDA_OK;
}

static bool misaligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) != 0; }

// the forward rows kernels address a batch entry with 31-bit BYTE offsets
static bool entry_bytes_overflow(const Dims& d, int value_dtype) {
  return static_cast<int64_t>(d.S) * d.M * d.D * static_cast<int64_t>(dtype_size(value_dtype)) >=
         (int64_t(1) << 31);
}

// ---- zero-fill on a side stream, concurrent with the forward kernel (knob clear_mode = 1) ----
// fork: the side stream waits for everything enqueued on `st` so far (the buffer may be a block
// the caller's allocator has just recycled from work still pending there), clears, and `st`
// joins after the forward kernel has been enqueued, so the kernel and the memset overlap and
// whatever the caller enqueues next (the backward) sees a cleared buffer.  One side stream and
// one event pair per device, serialised by a mutex for the few host microseconds of the enqueue.
// Capturable: the wait pulls the side stream into an ongoing capture (fork / join pattern).
struct SideClear {
  std::mutex mu;
  cudaStream_t stream[64] = {};
  cudaEvent_t fork[64] = {}, join[64] = {};
  bool ready[64] = {};
};
static SideClear g_side;

class SideClearScope {
 public:
  SideClearScope(void* clear, size_t bytes, cudaStream_t st) : st_(st) {
    if (!clear || !bytes || tuning().clear_mode != 1) return;
    if (cudaGetDevice(&dev_) != cudaSuccess || dev_ < 0 || dev_ >= 64) return;
    g_side.mu.lock();
    locked_ = true;
    if (!g_side.ready[dev_]) {
      if (cudaStreamCreateWithFlags(&g_side.stream[dev_], cudaStreamNonBlocking) != cudaSuccess ||
          cudaEventCreateWithFlags(&g_side.fork[dev_], cudaEventDisableTiming) != cudaSuccess ||
          cudaEventCreateWithFlags(&g_side.join[dev_], cudaEventDisableTiming) != cudaSuccess)
        return;
      g_side.ready[dev_] = true;
    }
    if (cudaEventRecord(g_side.fork[dev_], st) != cudaSuccess) return;
    if (cudaStreamWaitEvent(g_side.stream[dev_], g_side.fork[dev_], 0) != cudaSuccess) return;
    if (cudaMemsetAsync(clear, 0, bytes, g_side.stream[dev_]) != cudaSuccess) return;
    if (cudaEventRecord(g_side.join[dev_], g_side.stream[dev_]) != cudaSuccess) return;
    active_ = true;
  }
  // true: the clear is on its way, the launcher must not clear again
  bool active() const { return active_; }
  ~SideClearScope() {
    if (active_) cudaStreamWaitEvent(st_, g_side.join[dev_], 0);
    if (locked_) g_side.mu.unlock();
  }

 private:
  cudaStream_t st_;
  int dev_ = -1;
  bool locked_ = false, active_ = false;
};

}  // namespace msda

using namespace msda;

extern "C" {

int msda_abi_version(void) { return MSDA_ABI_VERSION; }

const char* msda_last_error(void) { return t_err; }

uint64_t msda_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int msda_set_option(const char* name, int value) {
  if (!name) return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_set_option: NULL name");
  Tuning& t = tuning_mut();
  int* slot = nullptr;
  if (!std::strcmp(name, "force_generic")) slot = &t.force_generic;
  else if (!std::strcmp(name, "fwd_split")) slot = &t.fwd_split;
  else if (!std::strcmp(name, "bwd_split")) slot = &t.bwd_split;
  else if (!std::strcmp(name, "flat")) slot = &t.flat;
  else if (!std::strcmp(name, "agg_min_level")) slot = &t.agg_min_level;
  else if (!std::strcmp(name, "agg_tile_kb")) slot = &t.agg_tile_kb;
  else if (!std::strcmp(name, "flat_fwd_cfg")) slot = &t.flat_fwd_cfg;
  else if (!std::strcmp(name, "flat_bwd_cfg")) slot = &t.flat_bwd_cfg;
  else if (!std::strcmp(name, "flat_order")) slot = &t.flat_order;
  else if (!std::strcmp(name, "clear_policy")) slot = &t.clear_policy;
  else if (!std::strcmp(name, "clear_mode")) slot = &t.clear_mode;
  else if (!std::strcmp(name, "l2_prefetch")) slot = &t.l2_prefetch;
  else if (!std::strcmp(name, "l2_prefetch_mb")) slot = &t.l2_prefetch_mb;
  else if (!std::strcmp(name, "bwd_variant")) slot = &t.bwd_variant;
  else if (!std::strcmp(name, "fwd_variant")) slot = &t.fwd_variant;
  if (!slot) return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_set_option: unknown option '%s'", name);
  *slot = value;
  return MSDA_OK;
}

uint64_t msda_launch_count_family(int family) {
  return (family >= 0 && family < KF_COUNT) ? g_family[family].load(std::memory_order_relaxed) : 0;
}

const char* msda_kernel_name(int channels, int dtype, int value_dtype) {
  if (dtype == MSDA_F32 && !tuning().force_generic && rows_supported(channels, value_dtype)) {
    if (value_dtype == MSDA_BF16)
      return channels == 16 ? "rows<D=16,bf16>" : channels == 32 ? "rows<D=32,bf16>" : "rows<D=64,bf16>";
    return channels == 16 ? "rows<D=16,f32>" : channels == 32 ? "rows<D=32,f32>" : "rows<D=64,f32>";
  }
  return "generic";
}

int msda_forward(const void* d_value, const int64_t* d_spatial_shapes,
                 const int64_t* d_level_start_index, const void* d_sampling_loc,
                 const void* d_attn_weight, void* d_output, int batch, int spatial_size,
                 int num_heads, int channels, int num_levels, int num_query, int num_point,
                 int dtype, int value_dtype, void* stream) {
  return msda_forward_clear(d_value, d_spatial_shapes, d_level_start_index, d_sampling_loc,
                            d_attn_weight, d_output, batch, spatial_size, num_heads, channels,
                            num_levels, num_query, num_point, dtype, value_dtype, nullptr, 0, stream);
}

int msda_forward_clear(const void* d_value, const int64_t* d_spatial_shapes,
                       const int64_t* d_level_start_index, const void* d_sampling_loc,
                       const void* d_attn_weight, void* d_output, int batch, int spatial_size,
                       int num_heads, int channels, int num_levels, int num_query, int num_point,
                       int dtype, int value_dtype, void* d_clear, size_t clear_bytes,
                       void* stream) {
  if ((d_clear == nullptr) != (clear_bytes == 0))
    return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_forward_clear: d_clear and clear_bytes must both be set or both be zero");
  if (!d_value || !d_spatial_shapes || !d_level_start_index || !d_sampling_loc || !d_attn_weight ||
      !d_output)
    return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_forward: NULL pointer argument");
  Dims d;
  int rc = check_dims(batch, spatial_size, num_heads, channels, num_levels, num_query, num_point, &d);
  if (rc) return rc;
  rc = check_dtypes(dtype, value_dtype, -1);
  if (rc) return rc;
  int sms = 0;
  rc = current_sm_count(&sms);
  if (rc) return rc;
  // the vector kernels need 16-byte aligned rows; fall back to scalar access otherwise
  // ... and keep byte offsets inside a batch entry in 31 bits (the generic kernel indexes in int64)
  const int generic = tuning().force_generic || misaligned16(d_value) || misaligned16(d_output) ||
                      misaligned16(d_sampling_loc) || entry_bytes_overflow(d, value_dtype);
  cudaError_t e;
  {
    SideClearScope side(d_clear, clear_bytes, static_cast<cudaStream_t>(stream));
    e = launch_forward(d_value, d_spatial_shapes, d_level_start_index, d_sampling_loc, d_attn_weight,
                       d_output, d, dtype, value_dtype, sms, generic, side.active() ? nullptr : d_clear,
                       side.active() ? 0 : clear_bytes, static_cast<cudaStream_t>(stream));
  }
  if (e != cudaSuccess)
    return fail(MSDA_ERR_CUDA, "msda_forward launch failed: %s", cudaGetErrorString(e));
  return MSDA_OK;
}

int msda_backward(const void* d_value, const int64_t* d_spatial_shapes,
                  const int64_t* d_level_start_index, const void* d_sampling_loc,
                  const void* d_attn_weight, const void* d_grad_output, void* d_grad_value,
                  void* d_grad_sampling_loc, void* d_grad_attn_weight, int batch,
                  int spatial_size, int num_heads, int channels, int num_levels, int num_query,
                  int num_point, int dtype, int value_dtype, int grad_value_dtype, void* stream) {
  if (!d_value || !d_spatial_shapes || !d_level_start_index || !d_sampling_loc || !d_attn_weight ||
      !d_grad_output || !d_grad_value || !d_grad_sampling_loc || !d_grad_attn_weight)
    return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_backward: NULL pointer argument");
  Dims d;
  int rc = check_dims(batch, spatial_size, num_heads, channels, num_levels, num_query, num_point, &d);
  if (rc) return rc;
  rc = check_dtypes(dtype, value_dtype, grad_value_dtype);
  if (rc) return rc;
  int sms = 0;
  rc = current_sm_count(&sms);
  if (rc) return rc;
  const int generic = tuning().force_generic || misaligned16(d_value) ||
                      misaligned16(d_grad_output) || misaligned16(d_grad_value) ||
                      misaligned16(d_sampling_loc) || misaligned16(d_grad_sampling_loc);
  const cudaError_t e = launch_backward(
      d_value, d_spatial_shapes, d_level_start_index, d_sampling_loc, d_attn_weight, d_grad_output,
      d_grad_value, d_grad_sampling_loc, d_grad_attn_weight, d, dtype, value_dtype,
      grad_value_dtype, sms, generic, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess)
    return fail(MSDA_ERR_CUDA, "msda_backward launch failed: %s", cudaGetErrorString(e));
  return MSDA_OK;
}

// ---------------------------------------------------------------------------
// fused entry points: softmax and the reference-point transform inside the kernels
// ---------------------------------------------------------------------------
static int fused_source(FusedSource* src, const float* off, const float* logit, const float* ref,
                        const float* scale, float* stats, int ref_per_level, int num_point) {
  if (ref_per_level != 1 && ref_per_level != num_point)
    return fail(MSDA_ERR_INVALID_ARGUMENT,
                "ref_points_per_level must be 1 or num_point (%d), got %d", num_point, ref_per_level);
  src->off = off; src->logit = logit; src->ref = ref; src->scale = scale; src->stats = stats;
  src->R = ref_per_level; src->P = num_point; src->stats_ready = 0; src->mx = 0.f; src->inv = 0.f;
  return MSDA_OK;
}

int msda_fused_forward(const void* d_value, const int64_t* d_spatial_shapes,
                       const int64_t* d_level_start_index, const float* d_offsets,
                       const float* d_logits, const float* d_ref_points, const float* d_scale,
                       float* d_output, float* d_softmax_stats, int batch, int spatial_size,
                       int num_heads, int channels, int num_levels, int num_query, int num_point,
                       int ref_points_per_level, int value_dtype, void* d_clear, size_t clear_bytes,
                       void* stream) {
  if (!d_value || !d_spatial_shapes || !d_level_start_index || !d_offsets || !d_logits ||
      !d_ref_points || !d_output || !d_softmax_stats)
    return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_fused_forward: NULL pointer argument");
  if ((d_clear == nullptr) != (clear_bytes == 0))
    return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_fused_forward: d_clear and clear_bytes must both be set or both be zero");
  Dims d;
  int rc = check_dims(batch, spatial_size, num_heads, channels, num_levels, num_query, num_point, &d);
  if (rc) return rc;
  rc = check_dtypes(MSDA_F32, value_dtype, -1);
  if (rc) return rc;
  if (misaligned16(d_value) || misaligned16(d_output) || misaligned16(d_offsets))
    return fail(MSDA_ERR_UNSUPPORTED, "msda_fused_forward needs 16-byte aligned buffers");
  if (entry_bytes_overflow(d, value_dtype))
    return fail(MSDA_ERR_UNSUPPORTED, "msda_fused_forward: one batch entry of value must be < 2 GiB");
  FusedSource src;
  rc = fused_source(&src, d_offsets, d_logits, d_ref_points, d_scale, d_softmax_stats,
                    ref_points_per_level, num_point);
  if (rc) return rc;
  int sms = 0;
  rc = current_sm_count(&sms);
  if (rc) return rc;
  cudaError_t e;
  {
    SideClearScope side(d_clear, clear_bytes, static_cast<cudaStream_t>(stream));
    e = launch_forward_fused(d_value, d_spatial_shapes, d_level_start_index, src, d_output, d,
                             value_dtype, sms, side.active() ? nullptr : d_clear,
                             side.active() ? 0 : clear_bytes, static_cast<cudaStream_t>(stream));
  }
  if (e == cudaErrorNotSupported)
    return fail(MSDA_ERR_UNSUPPORTED, "msda_fused_forward: only channels in {16, 32, 64} and <= %d levels",
                kMaxSmemLevels);
  if (e != cudaSuccess)
    return fail(MSDA_ERR_CUDA, "msda_fused_forward launch failed: %s", cudaGetErrorString(e));
  return MSDA_OK;
}

int msda_fused_backward(const void* d_value, const int64_t* d_spatial_shapes,
                        const int64_t* d_level_start_index, const float* d_offsets,
                        const float* d_logits, const float* d_ref_points, const float* d_scale,
                        const float* d_softmax_stats, const float* d_output,
                        const float* d_grad_output,
                        float* d_grad_value, float* d_grad_offsets, float* d_grad_logits,
                        float* d_grad_loc, int batch, int spatial_size, int num_heads,
                        int channels, int num_levels, int num_query, int num_point,
                        int ref_points_per_level, int value_dtype, void* stream) {
  if (!d_value || !d_spatial_shapes || !d_level_start_index || !d_offsets || !d_logits ||
      !d_ref_points || !d_softmax_stats || !d_output || !d_grad_output || !d_grad_value ||
      !d_grad_offsets || !d_grad_logits)
    return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_fused_backward: NULL pointer argument");
  Dims d;
  int rc = check_dims(batch, spatial_size, num_heads, channels, num_levels, num_query, num_point, &d);
  if (rc) return rc;
  rc = check_dtypes(MSDA_F32, value_dtype, MSDA_F32);
  if (rc) return rc;
  if (misaligned16(d_value) || misaligned16(d_grad_output) || misaligned16(d_grad_value) ||
      misaligned16(d_offsets) || misaligned16(d_grad_offsets) || misaligned16(d_output))
    return fail(MSDA_ERR_UNSUPPORTED, "msda_fused_backward needs 16-byte aligned buffers");
  FusedSource src;
  rc = fused_source(&src, d_offsets, d_logits, d_ref_points, d_scale,
                    const_cast<float*>(d_softmax_stats), ref_points_per_level, num_point);
  if (rc) return rc;
  int sms = 0;
  rc = current_sm_count(&sms);
  if (rc) return rc;
  const cudaError_t e = launch_backward_fused(
      d_value, d_spatial_shapes, d_level_start_index, src, d_output, d_grad_output, d_grad_value,
      d_grad_offsets, d_grad_logits, d_grad_loc, d, value_dtype, sms,
      static_cast<cudaStream_t>(stream));
  if (e == cudaErrorNotSupported)
    return fail(MSDA_ERR_UNSUPPORTED, "msda_fused_backward: only channels in {16, 32, 64} and <= %d levels",
                kMaxSmemLevels);
  if (e != cudaSuccess)
    return fail(MSDA_ERR_CUDA, "msda_fused_backward launch failed: %s", cudaGetErrorString(e));
  return MSDA_OK;
}

namespace {
// p in [0, 1) -> keep threshold on a 32-bit hash and the survivor scale
int dropout_params(const char* who, float p, uint32_t* threshold, float* scale) {
  if (!(p >= 0.f) || p >= 1.f)
    return fail(MSDA_ERR_INVALID_ARGUMENT, "%s: dropout_p must be in [0, 1), got %g", who, p);
  *threshold = static_cast<uint32_t>(static_cast<double>(p) * 4294967296.0);
  *scale = *threshold ? 1.f / (1.f - p) : 1.f;
  return MSDA_OK;
}
}  // namespace

int msda_linear_fused(const float* d_x, const float* d_weight, const float* d_bias,
                      const uint8_t* d_row_mask, int mask_mode, int relu, const float* d_gate,
                      float gate_scale, float dropout_p, uint64_t dropout_seed,
                      const uint64_t* d_dropout_seed, const float* d_residual, void* d_y, int rows, int in_features, int out_features,
                      int out_dtype, float* d_scratch, void* stream) {
  if (!d_x || !d_weight || !d_y || !d_scratch)
    return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_linear_fused: NULL pointer argument");
  if (rows <= 0) return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_linear_fused: rows must be positive");
  if (!linear_shape_supported(in_features, out_features))
    return fail(MSDA_ERR_UNSUPPORTED,
                "msda_linear_fused: (in, out) = (%d, %d): widths must be 128, 256 or 1024 (and not 128x128 / "
                "1024x1024)", in_features, out_features);
  if (out_dtype != MSDA_F32 && out_dtype != MSDA_BF16)
    return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_linear_fused: out_dtype must be MSDA_F32 or MSDA_BF16");
  if (mask_mode < 0 || mask_mode > 2 || (mask_mode != 0 && !d_row_mask))
    return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_linear_fused: bad mask_mode / row_mask");
  if (misaligned16(d_x) || misaligned16(d_weight) || misaligned16(d_y) || misaligned16(d_scratch) ||
      misaligned16(d_gate) || misaligned16(d_residual))
    return fail(MSDA_ERR_UNSUPPORTED, "msda_linear_fused needs 16-byte aligned buffers");
  LinearEpilogue ep;
  ep.bias = d_bias;
  ep.row_mask = mask_mode ? d_row_mask : nullptr;
  ep.mask_mode = mask_mode;
  ep.relu = relu != 0;
  ep.gate = d_gate;
  ep.gate_scale = gate_scale;
  const int rc = dropout_params("msda_linear_fused", dropout_p, &ep.dropout_threshold, &ep.dropout_scale);
  if (rc) return rc;
  ep.seed_lo = static_cast<uint32_t>(dropout_seed);
  ep.seed_hi = static_cast<uint32_t>(dropout_seed >> 32);
  ep.seed_ptr = reinterpret_cast<const unsigned long long*>(d_dropout_seed);
  ep.residual = d_residual;
  const cudaError_t e = launch_linear256(d_x, d_weight, ep, d_y, rows, in_features, out_features, out_dtype,
                                         d_scratch, static_cast<cudaStream_t>(stream));
  if (e == cudaErrorNotSupported)
    return fail(MSDA_ERR_UNSUPPORTED, "msda_linear_fused: cuTensorMapEncodeTiled unavailable or failed");
  if (e != cudaSuccess)
    return fail(MSDA_ERR_CUDA, "msda_linear_fused launch failed: %s", cudaGetErrorString(e));
  return MSDA_OK;
}

int msda_linear256(const float* d_x, const float* d_weight, const float* d_bias,
                   const uint8_t* d_row_mask, int mask_mode, void* d_y, int rows, int in_features,
                   int out_features, int out_dtype, float* d_scratch, void* stream) {
  return msda_linear_fused(d_x, d_weight, d_bias, d_row_mask, mask_mode, 0, nullptr, 1.f, 0.f, 0, nullptr,
                           nullptr, d_y, rows, in_features, out_features, out_dtype, d_scratch, stream);
}

int msda_linear256_wgrad(const float* d_grad_y, const float* d_x, const uint8_t* d_row_mask,
                         int mask_mode, float* d_grad_weight, int rows, int in_features,
                         int out_features, void* stream) {
  if (!d_grad_y || !d_x || !d_grad_weight)
    return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_linear256_wgrad: NULL pointer argument");
  if (rows <= 0) return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_linear256_wgrad: rows must be positive");
  if (!linear_shape_supported(in_features, out_features))
    return fail(MSDA_ERR_UNSUPPORTED,
                "msda_linear256_wgrad: (in, out) = (%d, %d): widths must be 128, 256 or 1024 (and not 128x128 / "
                "1024x1024)", in_features, out_features);
  if (mask_mode < 0 || mask_mode > 2 || (mask_mode != 0 && !d_row_mask))
    return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_linear256_wgrad: bad mask_mode / row_mask");
  if (misaligned16(d_grad_y) || misaligned16(d_x) || misaligned16(d_grad_weight))
    return fail(MSDA_ERR_UNSUPPORTED, "msda_linear256_wgrad needs 16-byte aligned buffers");
  int sms = 0;
  const int rc = current_sm_count(&sms);
  if (rc) return rc;
  const cudaError_t e = launch_linear256_wgrad(d_grad_y, d_x, d_row_mask, mask_mode, d_grad_weight, rows,
                                               in_features, out_features, sms,
                                               static_cast<cudaStream_t>(stream));
  if (e == cudaErrorNotSupported)
    return fail(MSDA_ERR_UNSUPPORTED, "msda_linear256_wgrad: cuTensorMapEncodeTiled unavailable or failed");
  if (e != cudaSuccess)
    return fail(MSDA_ERR_CUDA, "msda_linear256_wgrad launch failed: %s", cudaGetErrorString(e));
  return MSDA_OK;
}

int msda_colsum256(const float* d_grad_y, const uint8_t* d_row_mask, float* d_grad_bias, int rows,
                   int width, void* stream) {
  if (!d_grad_y || !d_grad_bias)
    return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_colsum256: NULL pointer argument");
  if (rows <= 0) return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_colsum256: rows must be positive");
  if (width != 128 && width != 256 && width != 1024)
    return fail(MSDA_ERR_UNSUPPORTED, "msda_colsum256: width must be 128, 256 or 1024, got %d", width);
  if (misaligned16(d_grad_y) || misaligned16(d_grad_bias))
    return fail(MSDA_ERR_UNSUPPORTED, "msda_colsum256 needs 16-byte aligned buffers");
  int sms = 0;
  const int rc = current_sm_count(&sms);
  if (rc) return rc;
  const cudaError_t e = launch_colsum256(d_grad_y, d_row_mask, d_grad_bias, nullptr, 0u, 1.f, 0u, 0u, nullptr,
                                         rows, width, sms, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess)
    return fail(MSDA_ERR_CUDA, "msda_colsum256 launch failed: %s", cudaGetErrorString(e));
  return MSDA_OK;
}

int msda_dropout_backward(const float* d_grad_y, float* d_grad_out, float* d_grad_bias, int rows,
                          int width, float dropout_p, uint64_t dropout_seed,
                          const uint64_t* d_dropout_seed, void* stream) {
  if (!d_grad_y || !d_grad_out)
    return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_dropout_backward: NULL pointer argument");
  if (rows <= 0) return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_dropout_backward: rows must be positive");
  if (width != 128 && width != 256 && width != 1024)
    return fail(MSDA_ERR_UNSUPPORTED, "msda_dropout_backward: width must be 128, 256 or 1024, got %d", width);
  if (misaligned16(d_grad_y) || misaligned16(d_grad_out) || misaligned16(d_grad_bias))
    return fail(MSDA_ERR_UNSUPPORTED, "msda_dropout_backward needs 16-byte aligned buffers");
  uint32_t threshold = 0;
  float scale = 1.f;
  int rc = dropout_params("msda_dropout_backward", dropout_p, &threshold, &scale);
  if (rc) return rc;
  if (threshold == 0)
    return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_dropout_backward: dropout_p rounds to zero; nothing to undo");
  int sms = 0;
  rc = current_sm_count(&sms);
  if (rc) return rc;
  const cudaError_t e = launch_colsum256(d_grad_y, nullptr, d_grad_bias, d_grad_out, threshold, scale,
                                         static_cast<uint32_t>(dropout_seed),
                                         static_cast<uint32_t>(dropout_seed >> 32),
                                         reinterpret_cast<const unsigned long long*>(d_dropout_seed), rows,
                                         width, sms,
                                         static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess)
    return fail(MSDA_ERR_CUDA, "msda_dropout_backward launch failed: %s", cudaGetErrorString(e));
  return MSDA_OK;
}

int msda_layernorm_forward(const float* d_x, const float* d_gamma, const float* d_beta, float* d_y,
                           float* d_mean, float* d_rstd, int rows, int width, float eps, void* stream) {
  if (!d_x || !d_gamma || !d_beta || !d_y || !d_mean || !d_rstd)
    return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_layernorm_forward: NULL pointer argument");
  if (rows <= 0) return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_layernorm_forward: rows must be positive");
  if (!layernorm_width_supported(width))
    return fail(MSDA_ERR_UNSUPPORTED, "msda_layernorm_forward: width must be 256, got %d", width);
  if (misaligned16(d_x) || misaligned16(d_gamma) || misaligned16(d_beta) || misaligned16(d_y))
    return fail(MSDA_ERR_UNSUPPORTED, "msda_layernorm_forward needs 16-byte aligned buffers");
  int sms = 0;
  const int rc = current_sm_count(&sms);
  if (rc) return rc;
  const cudaError_t e = launch_layernorm_forward(d_x, d_gamma, d_beta, d_y, d_mean, d_rstd, rows, eps, sms,
                                                 static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess)
    return fail(MSDA_ERR_CUDA, "msda_layernorm_forward launch failed: %s", cudaGetErrorString(e));
  return MSDA_OK;
}

int msda_layernorm_backward(const float* d_x, const float* d_grad_y, const float* d_gamma,
                            const float* d_mean, const float* d_rstd, float* d_grad_x,
                            float* d_grad_gamma, float* d_grad_beta, int rows, int width, void* stream) {
  if (!d_x || !d_grad_y || !d_gamma || !d_mean || !d_rstd || !d_grad_x || !d_grad_gamma || !d_grad_beta)
    return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_layernorm_backward: NULL pointer argument");
  if (rows <= 0) return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_layernorm_backward: rows must be positive");
  if (!layernorm_width_supported(width))
    return fail(MSDA_ERR_UNSUPPORTED, "msda_layernorm_backward: width must be 256, got %d", width);
  if (misaligned16(d_x) || misaligned16(d_grad_y) || misaligned16(d_gamma) || misaligned16(d_grad_x) ||
      misaligned16(d_grad_gamma) || misaligned16(d_grad_beta))
    return fail(MSDA_ERR_UNSUPPORTED, "msda_layernorm_backward needs 16-byte aligned buffers");
  int sms = 0;
  const int rc = current_sm_count(&sms);
  if (rc) return rc;
  const cudaError_t e = launch_layernorm_backward(d_x, d_grad_y, d_gamma, d_mean, d_rstd, d_grad_x,
                                                  d_grad_gamma, d_grad_beta, rows, sms,
                                                  static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess)
    return fail(MSDA_ERR_CUDA, "msda_layernorm_backward launch failed: %s", cudaGetErrorString(e));
  return MSDA_OK;
}

// ---------------------------------------------------------------------------
// host-buffer entry points
//
// The staged call is pipelined: work is cut into (batch entry, query chunk)
// pieces; piece k+1 is copied host->device on an upload stream while piece k
// runs forward + backward on the compute stream and piece k-1's results return
// on a download stream.  PCIe is full duplex, so with pinned host memory the
// call costs about max(bytes up, bytes down) / link bandwidth instead of their
// sum plus the kernels.  One stream per direction by default: alternating the
// pieces between two or more streams per direction (PAVENET_MSDA_COPY_STREAMS)
// was measured on B200 and is slower (7.8 vs 6.4 ms on config 2) -- concurrent
// copies in the same direction take turns on the link instead of hiding each
// other's start-up cost.
// ---------------------------------------------------------------------------
constexpr int kMaxCopyStreams = 4;
struct msda_workspace {
  cudaStream_t s_in[kMaxCopyStreams] = {}, s_cmp = nullptr, s_out[kMaxCopyStreams] = {};
  int n_copy = 1;
  void* buf = nullptr;   // one grow-only device arena
  size_t cap = 0;
  std::vector<cudaEvent_t> events;  // grow-only pool, reused across calls
  size_t next_event = 0;
  size_t piece_bytes = size_t(12) << 20;  // upload bytes per pipeline piece of a blocking call
  // ... and of a queued call: its upload-only head and download-only tail overlap with the neighbouring
  // calls, so larger copies (better duplex rate) win: 5.84 ms per step at 12 MiB, 5.60 at 32 MiB, 5.37-5.5 with
  // one piece per batch entry (config 2: 57 MB)
  size_t piece_bytes_async = size_t(64) << 20;
  // PAVENET_MSDA_TRACE_E2E=<file>: timing events at every pipeline stage, written as CSV after each call
  bool batch_copies = true;    // a piece's copies go out as one cudaMemcpyBatchAsync (e2e 5.78 -> 5.60 ms); PAVENET_MSDA_BATCH_COPIES=0 disables
  bool in_flight = false;   // an asynchronous call has been queued and not waited for yet
  const char* trace_path = nullptr;
  struct Mark { cudaEvent_t ev; char kind; int b, piece; };
  std::vector<Mark> marks;
};

int msda_workspace_create(msda_workspace** out_ws) {
  if (!out_ws) return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_workspace_create: NULL out pointer");
  msda_workspace* ws = new msda_workspace();
  ws->n_copy = msda::tuning().copy_streams;
  ws->trace_path = std::getenv("PAVENET_MSDA_TRACE_E2E");
  {
    const char* bc = std::getenv("PAVENET_MSDA_BATCH_COPIES");   // default on; 0 = one cudaMemcpyAsync per array
    ws->batch_copies = !(bc && bc[0] == '0');
  }
  std::vector<cudaStream_t*> st = {&ws->s_cmp};
  for (int i = 0; i < ws->n_copy; ++i) {
    st.push_back(&ws->s_in[i]);
    st.push_back(&ws->s_out[i]);
  }
  for (auto* p : st) {
    const cudaError_t e = cudaStreamCreateWithFlags(p, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
      msda_workspace_destroy(ws);
      return fail(MSDA_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e));
    }
  }
  *out_ws = ws;
  return MSDA_OK;
}

void msda_workspace_destroy(msda_workspace* ws) {
  if (!ws) return;
  if (ws->in_flight) {   // a queued call still reads / writes the caller's host buffers
    for (int i = 0; i < kMaxCopyStreams; ++i) {
      if (ws->s_out[i]) cudaStreamSynchronize(ws->s_out[i]);
      if (ws->s_in[i]) cudaStreamSynchronize(ws->s_in[i]);
    }
    if (ws->s_cmp) cudaStreamSynchronize(ws->s_cmp);
  }
  for (cudaEvent_t e : ws->events) cudaEventDestroy(e);
  if (ws->buf) cudaFree(ws->buf);
  for (int i = 0; i < kMaxCopyStreams; ++i) {
    if (ws->s_in[i]) cudaStreamDestroy(ws->s_in[i]);
    if (ws->s_out[i]) cudaStreamDestroy(ws->s_out[i]);
  }
  if (ws->s_cmp) cudaStreamDestroy(ws->s_cmp);
  delete ws;
}

int msda_workspace_set_piece_bytes(msda_workspace* ws, size_t bytes) {
  if (!ws || bytes == 0)
    return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_workspace_set_piece_bytes: NULL workspace or 0 bytes");
  ws->piece_bytes = bytes;
  ws->piece_bytes_async = bytes;
  return MSDA_OK;
}

void* msda_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) {
    fail(MSDA_ERR_CUDA, "cudaHostAlloc(%zu) failed", bytes);
    return nullptr;
  }
  return p;
}

void msda_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

namespace {
struct Arena {
  char* base;
  size_t off = 0;
  void* take(size_t bytes) {
    void* p = base + off;
    off += (bytes + 255) & ~size_t(255);
    return p;
  }
};
size_t pad256(size_t b) { return (b + 255) & ~size_t(255); }

int ws_reserve(msda_workspace* ws, size_t bytes) {
  if (bytes <= ws->cap) return MSDA_OK;
  if (ws->buf) cudaFree(ws->buf);
  ws->buf = nullptr;
  ws->cap = 0;
  const cudaError_t e = cudaMalloc(&ws->buf, bytes);
  if (e != cudaSuccess) return fail(MSDA_ERR_CUDA, "cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e));
  ws->cap = bytes;
  return MSDA_OK;
}

int ws_event(msda_workspace* ws, cudaEvent_t* out) {
  if (ws->next_event == ws->events.size()) {
    cudaEvent_t e;
    const cudaError_t rc = cudaEventCreateWithFlags(&e, ws->trace_path ? cudaEventDefault : cudaEventDisableTiming);
    if (rc != cudaSuccess) return fail(MSDA_ERR_CUDA, "cudaEventCreate: %s", cudaGetErrorString(rc));
    ws->events.push_back(e);
  }
  *out = ws->events[ws->next_event++];
  return MSDA_OK;
}

// trace mode: remember an already recorded event, or record a fresh one on `st`
int ws_mark(msda_workspace* ws, char kind, int b, int piece, cudaEvent_t ev, cudaStream_t st) {
  if (!ws->trace_path) return MSDA_OK;
  if (!ev) {
    const int rc = ws_event(ws, &ev);
    if (rc) return rc;
    if (cudaEventRecord(ev, st) != cudaSuccess) return fail(MSDA_ERR_CUDA, "trace: cudaEventRecord");
  }
  ws->marks.push_back({ev, kind, b, piece});
  return MSDA_OK;
}
void ws_write_trace(msda_workspace* ws) {
  if (!ws->trace_path || ws->marks.empty()) return;
  if (FILE* f = std::fopen(ws->trace_path, "w")) {
    std::fprintf(f, "kind,batch,piece,ms\n");   // S start, V value up, I piece inputs up, C compute done, O piece outputs down, G grad_value down
    for (const auto& m : ws->marks) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, ws->marks[0].ev, m.ev);
      std::fprintf(f, "%c,%d,%d,%.4f\n", m.kind, m.b, m.piece, ms);
    }
    std::fclose(f);
  }
  ws->marks.clear();
}

#define MSDA_CU(expr)                                                                     \
  do {                                                                                    \
    const cudaError_t e_ = (expr);                                                        \
    if (e_ != cudaSuccess) return fail(MSDA_ERR_CUDA, #expr ": %s", cudaGetErrorString(e_)); \
  } while (0)
#define MSDA_RC(expr)          \
  do {                         \
    const int rc_ = (expr);    \
    if (rc_) return rc_;       \
  } while (0)

// up to 4 copies of one direction queued on `st`: as one cudaMemcpyBatchAsync (fewer gaps between the
// copies of a piece: 3 % on the whole call) or one by one; sizes of 0 are skipped
cudaError_t copy_group(msda_workspace* ws, void* const* dsts, const void* const* srcs,
                       const size_t* sizes, int n, cudaMemcpyKind kind, cudaStream_t st) {
  void* d[4];
  void* s_[4];
  size_t z[4];
  int m = 0;
  for (int i = 0; i < n; ++i)
    if (sizes[i]) {
      d[m] = dsts[i];
      s_[m] = const_cast<void*>(srcs[i]);
      z[m] = sizes[i];
      ++m;
    }
  if (m == 0) return cudaSuccess;
  if (ws->batch_copies && m > 1) {
    cudaMemcpyAttributes attr = {};
    attr.srcAccessOrder = cudaMemcpySrcAccessOrderStream;
    size_t attr_idx = 0, fail_idx = 0;
    const cudaError_t e = cudaMemcpyBatchAsync(d, s_, z, static_cast<size_t>(m), &attr, &attr_idx, 1, &fail_idx, st);
    if (e == cudaSuccess) return e;
    (void)cudaGetLastError();      // not supported here: fall back for good
    ws->batch_copies = false;
  }
  for (int i = 0; i < m; ++i) {
    const cudaError_t e = cudaMemcpyAsync(d[i], s_[i], z[i], kind, st);
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

// queries per pipeline piece: pieces of >= ~8 MiB keep the link efficient
int pick_chunk(int num_query, size_t bytes_per_query, size_t target) {
  int64_t q = static_cast<int64_t>(target / (bytes_per_query ? bytes_per_query : 1));
  if (q < 1) q = 1;
  if (q > num_query) q = num_query;
  const int pieces = static_cast<int>((num_query + q - 1) / q);
  return (num_query + pieces - 1) / pieces;  // even pieces
}
}  // namespace

// the pipeline proper; on an error it returns at once, possibly with copies still in flight
static int host_pipeline(msda_workspace* ws, const void* h_value,
                               const int64_t* h_spatial_shapes, const int64_t* h_level_start_index,
                               const void* h_sampling_loc, const void* h_attn_weight,
                               const void* h_grad_output, void* h_output, void* h_grad_value,
                               void* h_grad_sampling_loc, void* h_grad_attn_weight, int batch,
                               int spatial_size, int num_heads, int channels, int num_levels,
                               int num_query, int num_point, int dtype, int value_dtype,
                               size_t piece_bytes) {
  const bool do_bwd = h_grad_output != nullptr;
  if (!ws || !h_value || !h_spatial_shapes || !h_level_start_index || !h_sampling_loc ||
      !h_attn_weight || (!do_bwd && !h_output) ||
      (do_bwd && (!h_grad_value || !h_grad_sampling_loc || !h_grad_attn_weight)))
    return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_*_host: NULL pointer argument");
  Dims d;
  MSDA_RC(check_dims(batch, spatial_size, num_heads, channels, num_levels, num_query, num_point, &d));
  MSDA_RC(check_dtypes(dtype, value_dtype, do_bwd ? dtype : -1));
  const size_t es = dtype_size(dtype), vs = dtype_size(value_dtype);
  // per batch entry
  const size_t n_val = static_cast<size_t>(spatial_size) * num_heads * channels;
  const size_t smp_q = static_cast<size_t>(num_heads) * num_levels * num_point;  // samples per query
  const size_t out_q = static_cast<size_t>(num_heads) * channels;                // outputs per query
  const size_t b_val = n_val * vs, b_gval = n_val * es;  // grad_value is accumulated and returned in `dtype`
  const size_t b_loc = smp_q * 2 * es * num_query, b_aw = smp_q * es * num_query;
  const size_t b_out = out_q * es * num_query;
  const size_t b_shp = static_cast<size_t>(num_levels) * 2 * 8, b_lsi = static_cast<size_t>(num_levels) * 8;
  size_t need = pad256(b_shp) + pad256(b_lsi) +
                batch * (pad256(b_val) + pad256(b_loc) + pad256(b_aw) + pad256(b_out));
  if (do_bwd) need += batch * (pad256(b_out) + pad256(b_gval) + pad256(b_loc) + pad256(b_aw));
  if (ws->in_flight)   // (checked before the arena may be re-allocated underneath the queued call)
    return fail(MSDA_ERR_INVALID_ARGUMENT,
                "msda_*_host: the previous asynchronous call on this workspace has not been waited for "
                "(msda_workspace_wait)");
  MSDA_RC(ws_reserve(ws, need));
  ws->next_event = 0;
  ws->marks.clear();
  MSDA_RC(ws_mark(ws, 'S', 0, 0, nullptr, ws->s_in[0]));
  Arena ar{static_cast<char*>(ws->buf)};
  int64_t* d_shp = static_cast<int64_t*>(ar.take(b_shp));
  int64_t* d_lsi = static_cast<int64_t*>(ar.take(b_lsi));
  MSDA_CU(cudaMemcpyAsync(d_shp, h_spatial_shapes, b_shp, cudaMemcpyHostToDevice, ws->s_in[0]));
  MSDA_CU(cudaMemcpyAsync(d_lsi, h_level_start_index, b_lsi, cudaMemcpyHostToDevice, ws->s_in[0]));

  const size_t up_q = (smp_q * 3 + (do_bwd ? out_q : 0)) * es;
  const int chunk = pick_chunk(num_query, up_q, piece_bytes);
  auto hoff = [](const void* p, size_t bytes) { return static_cast<const char*>(p) + bytes; };
  auto hoffw = [](void* p, size_t bytes) { return static_cast<char*>(p) + bytes; };

  // Monolithic form (piece size >= the whole call's uploads): every input tensor goes up as ONE copy (they are
  // contiguous over the batch entries), all four in one batch, the kernels run once over the whole batch, and the
  // results come back the same way.  Nothing overlaps inside the call; it is meant for callers that keep three or
  // more calls in flight, for whom the link then sees only 34-68 MB copies.
  const size_t up_total = (b_val + b_loc + b_aw + (do_bwd ? b_out : 0)) * static_cast<size_t>(batch);
  if (piece_bytes >= up_total && batch > 1) {
    const size_t nb = static_cast<size_t>(batch);
    char* d_val = static_cast<char*>(ar.take(nb * b_val));
    char* d_loc = static_cast<char*>(ar.take(nb * b_loc));
    char* d_aw = static_cast<char*>(ar.take(nb * b_aw));
    char* d_out = static_cast<char*>(ar.take(nb * b_out));
    char *d_go = nullptr, *d_gval = nullptr, *d_gloc = nullptr, *d_gaw = nullptr;
    if (do_bwd) {
      d_go = static_cast<char*>(ar.take(nb * b_out));
      d_gval = static_cast<char*>(ar.take(nb * b_gval));
      d_gloc = static_cast<char*>(ar.take(nb * b_loc));
      d_gaw = static_cast<char*>(ar.take(nb * b_aw));
      MSDA_CU(cudaMemsetAsync(d_gval, 0, nb * b_gval, ws->s_cmp));
    }
    {
      void* const dsts[4] = {d_val, d_loc, d_aw, d_go};
      const void* const srcs[4] = {h_value, h_sampling_loc, h_attn_weight, h_grad_output};
      const size_t sizes[4] = {nb * b_val, nb * b_loc, nb * b_aw, do_bwd ? nb * b_out : 0};
      MSDA_CU(copy_group(ws, dsts, srcs, sizes, 4, cudaMemcpyHostToDevice, ws->s_in[0]));
    }
    cudaEvent_t ev_in, ev_done;
    MSDA_RC(ws_event(ws, &ev_in));
    MSDA_CU(cudaEventRecord(ev_in, ws->s_in[0]));
    MSDA_CU(cudaStreamWaitEvent(ws->s_cmp, ev_in, 0));
    MSDA_RC(ws_mark(ws, 'I', 0, 0, ev_in, nullptr));
    if (h_output)
      MSDA_RC(msda_forward(d_val, d_shp, d_lsi, d_loc, d_aw, d_out, batch, spatial_size, num_heads, channels,
                           num_levels, num_query, num_point, dtype, value_dtype, ws->s_cmp));
    if (do_bwd)
      MSDA_RC(msda_backward(d_val, d_shp, d_lsi, d_loc, d_aw, d_go, d_gval, d_gloc, d_gaw, batch, spatial_size,
                            num_heads, channels, num_levels, num_query, num_point, dtype, value_dtype, dtype,
                            ws->s_cmp));
    MSDA_RC(ws_event(ws, &ev_done));
    MSDA_CU(cudaEventRecord(ev_done, ws->s_cmp));
    MSDA_CU(cudaStreamWaitEvent(ws->s_out[0], ev_done, 0));
    MSDA_RC(ws_mark(ws, 'C', 0, 0, ev_done, nullptr));
    {
      void* const dsts[4] = {h_output, h_grad_sampling_loc, h_grad_attn_weight, h_grad_value};
      const void* const srcs[4] = {d_out, d_gloc, d_gaw, d_gval};
      const size_t sizes[4] = {h_output ? nb * b_out : 0, do_bwd ? nb * b_loc : 0, do_bwd ? nb * b_aw : 0,
                               do_bwd ? nb * b_gval : 0};
      MSDA_CU(copy_group(ws, dsts, srcs, sizes, 4, cudaMemcpyDeviceToHost, ws->s_out[0]));
    }
    MSDA_RC(ws_mark(ws, 'G', 0, 0, nullptr, ws->s_out[0]));
    ws->in_flight = true;
    return MSDA_OK;
  }

  int piece = 0;  // global piece counter: piece k uploads on s_in[k % n], downloads on s_out[k % n]
  for (int b = 0; b < batch; ++b) {
    char* d_val = static_cast<char*>(ar.take(b_val));
    char* d_loc = static_cast<char*>(ar.take(b_loc));
    char* d_aw = static_cast<char*>(ar.take(b_aw));
    char* d_out = static_cast<char*>(ar.take(b_out));
    char *d_go = nullptr, *d_gval = nullptr, *d_gloc = nullptr, *d_gaw = nullptr;
    if (do_bwd) {
      d_go = static_cast<char*>(ar.take(b_out));
      d_gval = static_cast<char*>(ar.take(b_gval));
      d_gloc = static_cast<char*>(ar.take(b_loc));
      d_gaw = static_cast<char*>(ar.take(b_aw));
      MSDA_CU(cudaMemsetAsync(d_gval, 0, b_gval, ws->s_cmp));
    }
    // the entry's value (and, for the first entry, the level tables queued on s_in[0] above) must be
    // resident before any of its pieces runs
    MSDA_CU(cudaMemcpyAsync(d_val, hoff(h_value, b * b_val), b_val, cudaMemcpyHostToDevice, ws->s_in[0]));
    cudaEvent_t ev_val;
    MSDA_RC(ws_event(ws, &ev_val));
    MSDA_CU(cudaEventRecord(ev_val, ws->s_in[0]));
    MSDA_CU(cudaStreamWaitEvent(ws->s_cmp, ev_val, 0));
    MSDA_RC(ws_mark(ws, 'V', b, piece, ev_val, nullptr));
    cudaEvent_t ev_done = nullptr;
    for (int q0 = 0; q0 < num_query; q0 += chunk, ++piece) {
      cudaStream_t s_in = ws->s_in[piece % ws->n_copy], s_out = ws->s_out[piece % ws->n_copy];
      const int nq = (num_query - q0 < chunk) ? num_query - q0 : chunk;
      const size_t o_loc = smp_q * 2 * es * q0, o_aw = smp_q * es * q0, o_out = out_q * es * q0;
      const size_t n_loc = smp_q * 2 * es * nq, n_aw = smp_q * es * nq, n_out = out_q * es * nq;
      {
        void* const dsts[3] = {d_loc + o_loc, d_aw + o_aw, do_bwd ? d_go + o_out : nullptr};
        const void* const srcs[3] = {hoff(h_sampling_loc, b * b_loc + o_loc), hoff(h_attn_weight, b * b_aw + o_aw),
                                     do_bwd ? hoff(h_grad_output, b * b_out + o_out) : nullptr};
        const size_t sizes[3] = {n_loc, n_aw, do_bwd ? n_out : 0};
        MSDA_CU(copy_group(ws, dsts, srcs, sizes, 3, cudaMemcpyHostToDevice, s_in));
      }
      cudaEvent_t ev_in;
      MSDA_RC(ws_event(ws, &ev_in));
      MSDA_CU(cudaEventRecord(ev_in, s_in));
      MSDA_CU(cudaStreamWaitEvent(ws->s_cmp, ev_in, 0));
      MSDA_RC(ws_mark(ws, 'I', b, piece, ev_in, nullptr));
      if (h_output)
        MSDA_RC(msda_forward(d_val, d_shp, d_lsi, d_loc + o_loc, d_aw + o_aw, d_out + o_out, 1,
                             spatial_size, num_heads, channels, num_levels, nq, num_point, dtype,
                             value_dtype, ws->s_cmp));
      if (do_bwd)
        MSDA_RC(msda_backward(d_val, d_shp, d_lsi, d_loc + o_loc, d_aw + o_aw, d_go + o_out, d_gval,
                              d_gloc + o_loc, d_gaw + o_aw, 1, spatial_size, num_heads, channels,
                              num_levels, nq, num_point, dtype, value_dtype, dtype, ws->s_cmp));
      MSDA_RC(ws_event(ws, &ev_done));
      MSDA_CU(cudaEventRecord(ev_done, ws->s_cmp));
      MSDA_CU(cudaStreamWaitEvent(s_out, ev_done, 0));
      MSDA_RC(ws_mark(ws, 'C', b, piece, ev_done, nullptr));
      {
        void* const dsts[3] = {h_output ? hoffw(h_output, b * b_out + o_out) : nullptr,
                               do_bwd ? hoffw(h_grad_sampling_loc, b * b_loc + o_loc) : nullptr,
                               do_bwd ? hoffw(h_grad_attn_weight, b * b_aw + o_aw) : nullptr};
        const void* const srcs[3] = {d_out + o_out, do_bwd ? d_gloc + o_loc : nullptr,
                                     do_bwd ? d_gaw + o_aw : nullptr};
        const size_t sizes[3] = {h_output ? n_out : 0, do_bwd ? n_loc : 0, do_bwd ? n_aw : 0};
        MSDA_CU(copy_group(ws, dsts, srcs, sizes, 3, cudaMemcpyDeviceToHost, s_out));
      }
      MSDA_RC(ws_mark(ws, 'O', b, piece, nullptr, s_out));
    }
    if (do_bwd) {
      // every piece of this batch entry has been scattered once the last one's backward is done
      cudaStream_t s_out = ws->s_out[piece % ws->n_copy];
      MSDA_CU(cudaStreamWaitEvent(s_out, ev_done, 0));
      MSDA_CU(cudaMemcpyAsync(hoffw(h_grad_value, b * b_gval), d_gval, b_gval, cudaMemcpyDeviceToHost,
                              s_out));
      MSDA_RC(ws_mark(ws, 'G', b, piece, nullptr, s_out));
    }
  }
  ws->in_flight = true;     // everything is queued; msda_workspace_wait completes the call
  return MSDA_OK;
}

// block until everything queued on the workspace's streams is done (results are then in host memory)
static int ws_drain(msda_workspace* ws) {
  cudaError_t first = cudaSuccess;
  auto sync = [&](cudaStream_t st) {
    const cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess && first == cudaSuccess) first = e;
  };
  for (int i = 0; i < ws->n_copy; ++i) sync(ws->s_out[i]);
  sync(ws->s_cmp);
  for (int i = 0; i < ws->n_copy; ++i) sync(ws->s_in[i]);
  ws->in_flight = false;
  if (first != cudaSuccess)
    return fail(MSDA_ERR_CUDA, "msda_workspace_wait: %s", cudaGetErrorString(first));
  return MSDA_OK;
}

int msda_forward_backward_host(msda_workspace* ws, const void* h_value,
                               const int64_t* h_spatial_shapes, const int64_t* h_level_start_index,
                               const void* h_sampling_loc, const void* h_attn_weight,
                               const void* h_grad_output, void* h_output, void* h_grad_value,
                               void* h_grad_sampling_loc, void* h_grad_attn_weight, int batch,
                               int spatial_size, int num_heads, int channels, int num_levels,
                               int num_query, int num_point, int dtype, int value_dtype) {
  const int rc = host_pipeline(ws, h_value, h_spatial_shapes, h_level_start_index, h_sampling_loc,
                               h_attn_weight, h_grad_output, h_output, h_grad_value,
                               h_grad_sampling_loc, h_grad_attn_weight, batch, spatial_size,
                               num_heads, channels, num_levels, num_query, num_point, dtype,
                               value_dtype, ws ? ws->piece_bytes : 0);
  if (rc != MSDA_OK) {
    // never hand the caller's host buffers back while a copy into or out of them is still queued
    // (the message of the original failure stays in msda_last_error)
    if (ws && !ws->in_flight) {
      for (int i = 0; i < ws->n_copy; ++i) cudaStreamSynchronize(ws->s_out[i]);
      cudaStreamSynchronize(ws->s_cmp);
      for (int i = 0; i < ws->n_copy; ++i) cudaStreamSynchronize(ws->s_in[i]);
    }
    return rc;
  }
  const int wrc = ws_drain(ws);
  ws_write_trace(ws);
  return wrc;
}

int msda_forward_backward_host_async(msda_workspace* ws, const void* h_value,
                                     const int64_t* h_spatial_shapes,
                                     const int64_t* h_level_start_index, const void* h_sampling_loc,
                                     const void* h_attn_weight, const void* h_grad_output,
                                     void* h_output, void* h_grad_value, void* h_grad_sampling_loc,
                                     void* h_grad_attn_weight, int batch, int spatial_size,
                                     int num_heads, int channels, int num_levels, int num_query,
                                     int num_point, int dtype, int value_dtype) {
  const int rc = host_pipeline(ws, h_value, h_spatial_shapes, h_level_start_index, h_sampling_loc,
                               h_attn_weight, h_grad_output, h_output, h_grad_value,
                               h_grad_sampling_loc, h_grad_attn_weight, batch, spatial_size,
                               num_heads, channels, num_levels, num_query, num_point, dtype,
                               value_dtype, ws ? ws->piece_bytes_async : 0);
  if (rc != MSDA_OK && ws && !ws->in_flight) {   // failed half way: nothing may stay queued
    for (int i = 0; i < ws->n_copy; ++i) cudaStreamSynchronize(ws->s_out[i]);
    cudaStreamSynchronize(ws->s_cmp);
    for (int i = 0; i < ws->n_copy; ++i) cudaStreamSynchronize(ws->s_in[i]);
  }
  return rc;
}

int msda_workspace_wait(msda_workspace* ws) {
  if (!ws) return fail(MSDA_ERR_INVALID_ARGUMENT, "msda_workspace_wait: NULL workspace");
  if (!ws->in_flight) return MSDA_OK;
  const int rc = ws_drain(ws);
  ws_write_trace(ws);
  return rc;
}

int msda_forward_host(msda_workspace* ws, const void* h_value, const int64_t* h_spatial_shapes,
                      const int64_t* h_level_start_index, const void* h_sampling_loc,
                      const void* h_attn_weight, void* h_output, int batch, int spatial_size,
                      int num_heads, int channels, int num_levels, int num_query, int num_point,
                      int dtype, int value_dtype) {
  return msda_forward_backward_host(ws, h_value, h_spatial_shapes, h_level_start_index,
                                    h_sampling_loc, h_attn_weight, nullptr, h_output, nullptr,
                                    nullptr, nullptr, batch, spatial_size, num_heads, channels,
                                    num_levels, num_query, num_point, dtype, value_dtype);
}

}  // extern "C"
