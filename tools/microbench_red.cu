// microbench_red.cu — how fast can sm_100a scatter 128-byte rows into a
// grad_value-sized buffer?  Compares, on the access pattern of the backward
// kernel (a lane group owns one 32-channel row at a pseudo-random pixel):
//   ld.v4.f32 gather / st.v4.f32 / red.add.f32 (scalar, 32 lanes per row) /
//   red.add.v2.f32 / red.add.v4.f32 (8 lanes per row) / red.add.v4.bf16x2 (4 lanes per row)
// for (a) uniformly random rows and (b) rows drawn from a small hot set
// (coarse pyramid levels).  Build: nvcc -arch=sm_100a -O3 -o microbench_red microbench_red.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t mix(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x;
}

enum Mode { LD4 = 0, ST4, RED1, RED2, RED4, REDBF, RED4W };  // RED4W: 256-byte rows (two adjacent fp32 rows), 16 lanes

template <int MODE>
__global__ void k(float* buf, uint32_t n_rows, uint32_t hot_rows, int iters, float* sink) {
  // lanes per 128-byte fp32 row: 32 (scalar), 16 (v2), 8 (v4); bf16 row is 64 bytes: 4 lanes
  constexpr int G = MODE == RED1 ? 32 : (MODE == RED2 || MODE == RED4W) ? 16 : MODE == REDBF ? 4 : 8;
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t grp = tid / G, gl = tid % G;
  float acc = 0.f;
  for (int i = 0; i < iters; ++i) {
    uint32_t r = mix(grp * 9781u + i * 7919u);
    r = hot_rows ? (r % hot_rows) * (n_rows / hot_rows) : r % n_rows;
    float* row = buf + (size_t)r * 32;
    if (MODE == LD4) {
      float4 v = *reinterpret_cast<const float4*>(row + gl * 4);
      acc += v.x + v.y + v.z + v.w;
    } else if (MODE == ST4) {
      *reinterpret_cast<float4*>(row + gl * 4) = make_float4(1.f, 2.f, 3.f, 4.f);
    } else if (MODE == RED1) {
      asm volatile("red.global.add.f32 [%0], %1;" ::"l"(row + gl), "f"(1.0f) : "memory");
    } else if (MODE == RED2) {
      asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(row + gl * 2), "f"(1.0f), "f"(2.0f) : "memory");
    } else if (MODE == RED4W) {
      float* row2 = buf + (size_t)(r & ~1u) * 32;   // an aligned pair of rows = 256 contiguous bytes
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(row2 + gl * 4), "f"(1.0f), "f"(2.0f), "f"(3.0f), "f"(4.0f) : "memory");
    } else if (MODE == RED4) {
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(row + gl * 4), "f"(1.0f), "f"(2.0f), "f"(3.0f), "f"(4.0f) : "memory");
    } else {
      // 32 bf16 = 64 bytes per row; rows packed at 64-byte pitch in the same buffer
      float* row16 = buf + (size_t)r * 16;
      asm volatile("red.global.add.noftz.v4.bf16x2 [%0], {%1, %2, %3, %4};" ::"l"(row16 + gl * 4), "r"(0x3f803f80u), "r"(0x3f803f80u), "r"(0x3f803f80u), "r"(0x3f803f80u) : "memory");
    }
  }
  if (MODE == LD4 && acc == 123.456f) *sink = acc;
}

template <int MODE>
void run(const char* name, float* buf, uint32_t n_rows, uint32_t hot, float* sink) {
  constexpr int G = MODE == RED1 ? 32 : (MODE == RED2 || MODE == RED4W) ? 16 : MODE == REDBF ? 4 : 8;
  const int iters = 64;
  const long rows_total = 1L << 25;                 // 32 Mi rows per launch
  const long groups = rows_total / iters;
  const long threads = groups * G;
  const int block = 256;
  const unsigned grid = (unsigned)((threads + block - 1) / block);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  k<MODE><<<grid, block>>>(buf, n_rows, hot, iters, sink);
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(a);
    k<MODE><<<grid, block>>>(buf, n_rows, hot, iters, sink);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
  }
  const double row_bytes = MODE == REDBF ? 64.0 : MODE == RED4W ? 256.0 : 128.0;
  printf("%-22s hot=%-6u  %8.3f ms  %7.2f Grows/s  %8.1f GB/s payload\n", name, hot, best,
         rows_total / best * 1e-6, rows_total * row_bytes / best * 1e-6);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) printf("  CUDA error: %s\n", cudaGetErrorString(e));
}

int main() {
  const uint32_t n_rows = 3u * 22223u * 8u;        // grad_value of config 2: rows of 32 floats
  float* buf; float* sink;
  cudaMalloc(&buf, (size_t)n_rows * 128); cudaMalloc(&sink, 4);
  cudaMemset(buf, 0, (size_t)n_rows * 128);
  for (uint32_t hot : {0u, 2184u}) {               // 2184 = 13*21*8 rows of the coarsest level
    run<LD4>("ld.v4.f32", buf, n_rows, hot, sink);
    run<ST4>("st.v4.f32", buf, n_rows, hot, sink);
    run<RED1>("red.add.f32 (x32)", buf, n_rows, hot, sink);
    run<RED2>("red.add.v2.f32 (x16)", buf, n_rows, hot, sink);
    run<RED4>("red.add.v4.f32 (x8)", buf, n_rows, hot, sink);
    run<REDBF>("red.add.v4.bf16x2 (x4)", buf, n_rows, hot, sink);
    run<RED4W>("red.add.v4.f32 (x16, 256 B)", buf, n_rows, hot, sink);
  }
  return 0;
}
