// microbench_gather.cu — how fast can one sm_100a SM gather 128-byte value rows, as a function of
// how many rows ONE warp instruction touches?  The forward kernel reads a row with 8 lanes x 16 bytes
// (4 different rows per LDG.128); the microarchitecture notes quote 2.07 clocks per wavefront *within*
// a multi-line LDG against ~1.0 across LDGs, so narrower loads (16 lanes x 8 B = 2 rows, 32 lanes x 4 B =
// 1 row per instruction) might move more rows per clock.  Also: the same rows read from shared memory
// (LDS.128, 4 rows per instruction) as the ceiling of a TMA-staged tile.
// Rows: uniformly random over config 2's value (L2 resident), or a hot set that fits in L1.
// Build: nvcc -arch=sm_100a -O3 -o microbench_gather microbench_gather.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t mix(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x;
}

// G lanes per 128-byte row; each lane loads 128/G bytes
template <int G, bool SMEM>
__global__ void __launch_bounds__(256) k(const float* __restrict__ buf, uint32_t n_rows, int iters, float* sink) {
  extern __shared__ float tile[];
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t grp = tid / G, gl = tid % G;
  if (SMEM) {
    for (uint32_t i = threadIdx.x; i < n_rows * 32; i += blockDim.x) tile[i] = buf[i];
    __syncthreads();
  }
  const float* base = SMEM ? tile : buf;
  float acc = 0.f;
  // cheap per-row index (one multiply-add + one mask) so that the 1-row-per-instruction variant is not
  // bound by issue slots; n_rows is a power of two
  uint32_t state = mix(grp * 9781u + 17u);
  const uint32_t mask = n_rows - 1;
#pragma unroll 1
  for (int i = 0; i < iters; i += 8) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      state = state * 1664525u + 1013904223u;
      const uint32_t r = (state >> 9) & mask;
      const float* row = base + (size_t)r * 32;
      if (G == 8) {
        const float4 v = *reinterpret_cast<const float4*>(row + gl * 4);
        acc += v.x + v.y + v.z + v.w;
      } else if (G == 16) {
        const float2 v = *reinterpret_cast<const float2*>(row + gl * 2);
        acc += v.x + v.y;
      } else {
        acc += row[gl];
      }
    }
  }
  if (acc == 123.456f) *sink = acc;
}

template <int G, bool SMEM>
void run(const char* name, const float* buf, uint32_t n_rows, float* sink) {
  const int iters = SMEM ? 2048 : 64;              // shared memory: amortise staging the tile
  const long rows_total = 1L << 25;                 // 32 Mi rows per launch
  const long threads = rows_total / iters * G;
  const int block = 256;
  const unsigned grid = (unsigned)((threads + block - 1) / block);
  const size_t smem = SMEM ? (size_t)n_rows * 128 : 0;
  if (SMEM) cudaFuncSetAttribute(k<G, SMEM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  k<G, SMEM><<<grid, block, smem>>>(buf, n_rows, iters, sink);
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(a);
    k<G, SMEM><<<grid, block, smem>>>(buf, n_rows, iters, sink);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
  }
  int dev = 0, sms = 0, khz = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  const double rows_per_s = rows_total / (best * 1e-3);
  printf("%-34s rows=%-8u %8.3f ms  %7.2f Grows/s  %6.2f clk/row/SM (at %d MHz)\n", name, n_rows, best,
         rows_per_s * 1e-9, sms * (khz * 1e3) / rows_per_s, khz / 1000);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) printf("  CUDA error: %s\n", cudaGetErrorString(e));
}

int main() {
  const uint32_t n_rows = 1u << 19;                // ~ value of config 2: 524 288 rows of 32 floats (64 MB, L2 resident)
  float* buf; float* sink;
  cudaMalloc(&buf, (size_t)n_rows * 128); cudaMalloc(&sink, 4);
  cudaMemset(buf, 0, (size_t)n_rows * 128);
  for (uint32_t rows : {n_rows, 512u}) {           // 512 rows = 64 KB: L1 resident
    run<8, false>("ld.v4.f32  (4 rows / instruction)", buf, rows, sink);
    run<16, false>("ld.v2.f32  (2 rows / instruction)", buf, rows, sink);
    run<32, false>("ld.f32     (1 row  / instruction)", buf, rows, sink);
  }
  // the same gathers from a 64 KB shared-memory tile (what a TMA-staged window would cost to read)
  run<8, true>("lds.128    (4 rows / instruction)", buf, 512u, sink);
  run<16, true>("lds.64     (2 rows / instruction)", buf, 512u, sink);
  run<32, true>("lds.32     (1 row  / instruction)", buf, 512u, sink);
  return 0;
}
