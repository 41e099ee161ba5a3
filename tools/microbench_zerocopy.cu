// microbench_zerocopy.cu — can SM-issued loads / stores on mapped pinned host memory move data across the
// host link as fast as the copy engines, in both directions at once?  (msda_forward_backward_host is bound
// by the duplex rate of queued cudaMemcpyAsync copies: 44 GB/s each way, tools/exp_copy_sizes.py.)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench_zerocopy microbench_zerocopy.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int UNROLL>
__global__ void copy_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, size_t n16) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + (UNROLL - 1) * stride < n16; i += UNROLL * stride) {
    uint4 v[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) v[u] = __ldcs(src + i + u * stride);
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) __stcs(dst + i + u * stride, v[u]);
  }
  for (; i < n16; i += stride) __stcs(dst + i, __ldcs(src + i));
}

int main() {
  const size_t bytes = 240u << 20, n16 = bytes / 16;
  uint4 *h_up, *h_dn, *d_up, *d_dn;
  cudaHostAlloc(&h_up, bytes, cudaHostAllocDefault);
  cudaHostAlloc(&h_dn, bytes, cudaHostAllocDefault);
  cudaMalloc(&d_up, bytes); cudaMalloc(&d_dn, bytes);
  cudaMemset(d_dn, 1, bytes);
  for (size_t i = 0; i < n16; ++i) h_up[i] = make_uint4(1, 2, 3, 4);
  cudaStream_t s1, s2; cudaStreamCreate(&s1); cudaStreamCreate(&s2);
  cudaEvent_t e0, e1, e2; cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2);
  printf("blocks x threads   H2D alone   D2H alone   duplex up / down   [GB/s, SM loads/stores on mapped pinned memory]\n");
  for (int blocks : {8, 16, 32, 64, 148, 296}) {
    for (int threads : {256, 1024}) {
      float best[3] = {1e9f, 1e9f, 1e9f}, bestd = 1e9f, upd = 0, dnd = 0;
      for (int rep = 0; rep < 3; ++rep) {
        float ms;
        cudaDeviceSynchronize();
        cudaEventRecord(e0, s1);
        copy_kernel<8><<<blocks, threads, 0, s1>>>(h_up, d_up, n16);
        cudaEventRecord(e1, s1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1); if (ms < best[0]) best[0] = ms;
        cudaEventRecord(e0, s2);
        copy_kernel<8><<<blocks, threads, 0, s2>>>(d_dn, h_dn, n16);
        cudaEventRecord(e2, s2); cudaEventSynchronize(e2);
        cudaEventElapsedTime(&ms, e0, e2); if (ms < best[1]) best[1] = ms;
        cudaDeviceSynchronize();
        cudaEventRecord(e0, s1);
        cudaStreamWaitEvent(s2, e0, 0);
        copy_kernel<8><<<blocks, threads, 0, s1>>>(h_up, d_up, n16);
        cudaEventRecord(e1, s1);
        copy_kernel<8><<<blocks, threads, 0, s2>>>(d_dn, h_dn, n16);
        cudaEventRecord(e2, s2);
        cudaDeviceSynchronize();
        float a, b; cudaEventElapsedTime(&a, e0, e1); cudaEventElapsedTime(&b, e0, e2);
        if ((a > b ? a : b) < bestd) { bestd = a > b ? a : b; upd = a; dnd = b; }
      }
      printf("%4d x %4d        %7.1f     %7.1f     %7.1f / %7.1f\n", blocks, threads, bytes / best[0] * 1e-6,
             bytes / best[1] * 1e-6, bytes / upd * 1e-6, bytes / dnd * 1e-6);
    }
  }
  // the copy engines on the same buffers, one big copy each way
  {
    float a, b;
    cudaDeviceSynchronize();
    cudaEventRecord(e0, s1); cudaStreamWaitEvent(s2, e0, 0);
    cudaMemcpyAsync(d_up, h_up, bytes, cudaMemcpyHostToDevice, s1); cudaEventRecord(e1, s1);
    cudaMemcpyAsync(h_dn, d_dn, bytes, cudaMemcpyDeviceToHost, s2); cudaEventRecord(e2, s2);
    cudaDeviceSynchronize();
    cudaEventElapsedTime(&a, e0, e1); cudaEventElapsedTime(&b, e0, e2);
    printf("copy engines, 240 MiB each way at once: %7.1f / %7.1f GB/s\n", bytes / a * 1e-6, bytes / b * 1e-6);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) printf("CUDA error: %s\n", cudaGetErrorString(e));
  return 0;
}
