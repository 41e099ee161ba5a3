// refcuda_driver.cu — C entry points around the REFERENCE's own CUDA kernels
// (third_party/mmcv/mmcv/ops/csrc/common/cuda/ms_deform_attn_cuda_kernel.cuh, included from the
// read-only reference checkout at build time; nothing of it is copied here), so that the
// reference's GPU implementation can be timed and compared on the same B200 as the new kernels
// ("the bar is the generic SIMT kernel recompiled for sm_100").
//
// TEST / BENCH INFRASTRUCTURE ONLY (oracle/): never linked into or called by pavenet_b200.
//
// Launch configuration restated from the reference's host code
// (third_party/mmcv/mmcv/ops/csrc/pytorch/cuda/ms_deform_attn_cuda.cu:22-207, 209-351):
//   forward : one thread per output scalar, 512 threads per block, at most 4096 blocks
//             (:35-43), output zero-filled first (:247-248);
//   backward: blockDim = channels; the kernel is picked by the channel count (:63-206) —
//             powers of two up to 32 use ..._shm_blocksize_aware_reduce_v1<channels>,
//             64..512 ..._reduce_v2<channels>, other counts the run-time-sized shm_reduce_v1/v2,
//             more than 512 channels the multi-block / global-memory variants;
//   both loop over the batch in chunks of im2col_step = min(batch, 64) (:233-236, 259-276).
#include <cstdint>
#include <cstdio>

#include "ms_deform_attn_cuda_kernel.cuh"

namespace {

template <typename T>
cudaError_t forward_t(const T* value, const int64_t* shapes, const int64_t* lsi, const T* loc,
                      const T* aw, T* out, int B, int S, int M, int D, int L, int Q, int P,
                      cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(out, 0, sizeof(T) * B * Q * M * D, st);
  if (e != cudaSuccess) return e;
  const int step = B < 64 ? B : 64;
  for (int n = 0; n < B / step; ++n) {
    const int num_kernels = step * Q * M * D;
    ms_deformable_im2col_gpu_kernel<T><<<GET_BLOCKS(num_kernels, THREADS_PER_BLOCK),
                                         THREADS_PER_BLOCK, 0, st>>>(
        num_kernels, value + (size_t)n * step * S * M * D, shapes, lsi,
        loc + (size_t)n * step * Q * M * L * P * 2, aw + (size_t)n * step * Q * M * L * P, step, S,
        M, D, L, Q, P, out + (size_t)n * step * Q * M * D);
  }
  return cudaGetLastError();
}

template <typename T>
cudaError_t backward_t(const T* value, const int64_t* shapes, const int64_t* lsi, const T* loc,
                       const T* aw, const T* go, T* gv, T* gl, T* ga, int B, int S, int M, int D,
                       int L, int Q, int P, cudaStream_t st) {
  const int step = B < 64 ? B : 64;
  for (int n = 0; n < B / step; ++n) {
    const int num_kernels = step * Q * M * D;
    const int threads = D > THREADS_PER_BLOCK ? THREADS_PER_BLOCK : D;
    const int blocks = GET_BLOCKS(num_kernels, threads);
    const T* v = value + (size_t)n * step * S * M * D;
    const T* lo = loc + (size_t)n * step * Q * M * L * P * 2;
    const T* a = aw + (size_t)n * step * Q * M * L * P;
    const T* g = go + (size_t)n * step * Q * M * D;
    T* gvn = gv + (size_t)n * step * S * M * D;
    T* gln = gl + (size_t)n * step * Q * M * L * P * 2;
    T* gan = ga + (size_t)n * step * Q * M * L * P;
#define REF_ARGS num_kernels, g, v, shapes, lsi, lo, a, step, S, M, D, L, Q, P, gvn, gln, gan
#define REF_V1(C) case C: ms_deformable_col2im_gpu_kernel_shm_blocksize_aware_reduce_v1<T, C> \
                      <<<blocks, threads, 0, st>>>(REF_ARGS); break;
#define REF_V2(C) case C: ms_deformable_col2im_gpu_kernel_shm_blocksize_aware_reduce_v2<T, C> \
                      <<<blocks, threads, 0, st>>>(REF_ARGS); break;
    if (D > THREADS_PER_BLOCK) {
      if ((D & (THREADS_PER_BLOCK - 1)) == 0)
        ms_deformable_col2im_gpu_kernel_shm_reduce_v2_multi_blocks<T>
            <<<blocks, threads, threads * 3 * sizeof(T), st>>>(REF_ARGS);
      else
        ms_deformable_col2im_gpu_kernel_gm<T><<<blocks, threads, 0, st>>>(REF_ARGS);
    } else {
      switch (D) {
        REF_V1(1) REF_V1(2) REF_V1(4) REF_V1(8) REF_V1(16) REF_V1(32)
        REF_V2(64) REF_V2(128) REF_V2(256) REF_V2(512)
        default:
          if (D < 64)
            ms_deformable_col2im_gpu_kernel_shm_reduce_v1<T>
                <<<blocks, threads, threads * 3 * sizeof(T), st>>>(REF_ARGS);
          else
            ms_deformable_col2im_gpu_kernel_shm_reduce_v2<T>
                <<<blocks, threads, threads * 3 * sizeof(T), st>>>(REF_ARGS);
      }
    }
#undef REF_V1
#undef REF_V2
#undef REF_ARGS
  }
  return cudaGetLastError();
}

}  // namespace

extern "C" {

// dtype: 0 = float, 1 = double.  All pointers are device pointers; the caller zero-fills the
// three gradient buffers (as the reference's autograd wrapper does, multi_scale_deform_attn.py:72-74).
int refcuda_forward(const void* value, const int64_t* shapes, const int64_t* lsi, const void* loc,
                    const void* aw, void* out, int B, int S, int M, int D, int L, int Q, int P,
                    int dtype, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const cudaError_t e =
      dtype == 0 ? forward_t<float>((const float*)value, shapes, lsi, (const float*)loc,
                                    (const float*)aw, (float*)out, B, S, M, D, L, Q, P, st)
                 : forward_t<double>((const double*)value, shapes, lsi, (const double*)loc,
                                     (const double*)aw, (double*)out, B, S, M, D, L, Q, P, st);
  return e == cudaSuccess ? 0 : -(int)e;
}

int refcuda_backward(const void* value, const int64_t* shapes, const int64_t* lsi, const void* loc,
                     const void* aw, const void* grad_out, void* grad_value, void* grad_loc,
                     void* grad_aw, int B, int S, int M, int D, int L, int Q, int P, int dtype,
                     void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const cudaError_t e =
      dtype == 0
          ? backward_t<float>((const float*)value, shapes, lsi, (const float*)loc,
                              (const float*)aw, (const float*)grad_out, (float*)grad_value,
                              (float*)grad_loc, (float*)grad_aw, B, S, M, D, L, Q, P, st)
          : backward_t<double>((const double*)value, shapes, lsi, (const double*)loc,
                               (const double*)aw, (const double*)grad_out, (double*)grad_value,
                               (double*)grad_loc, (double*)grad_aw, B, S, M, D, L, Q, P, st);
  return e == cudaSuccess ? 0 : -(int)e;
}

}  // extern "C"
