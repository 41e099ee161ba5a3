// microbench_tma_red.cu — can the TMA engine reduce 128-byte rows into global memory faster
// than the LSU's red.global.add.v4.f32 (54 G rows/s = 5.4 clocks per row and SM, the floor of
// the backward kernels)?  Each warp stages weighted rows in shared memory (st.shared.v4, 4 rows
// per instruction), makes them visible to the async proxy, and one lane per row issues
//   cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [row], [smem], BYTES
// (SASS UBLKRED) at a pseudo-random row; a ring of staging buffers is recycled with
// cp.async.bulk.wait_group.read.  Compared with red.global.add.v4.f32 on the same rows.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench_tma_red microbench_tma_red.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t mix(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x;
}

constexpr int kRowsPerStage = 16;    // rows a warp stages per round (one lane issues one row)
constexpr int kStages = 4;

// MODE 0: red.global.add.v4.f32 (LSU)   1: TMA reduce, BYTES per op   2: TMA plain store
// MODE 3: both at once -- every round a warp reduces 16 rows through the LSU AND 16 through the TMA
template <int MODE, int BYTES>
__global__ void __launch_bounds__(256) k(float* buf, uint32_t n_rows, uint32_t hot_rows, int rounds) {
  extern __shared__ __align__(128) float smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t gw = blockIdx.x * (blockDim.x >> 5) + warp;
  constexpr int kRowFloats = BYTES / 4;
  float* stage0 = smem + warp * (kStages * kRowsPerStage * kRowFloats);
  for (int it = 0; it < rounds; ++it) {
    if (MODE == 0 || MODE == 3) {
      // 16 rows per round like the TMA modes: 4 instructions x 4 rows
#pragma unroll
      for (int j = 0; j < kRowsPerStage / 4; ++j) {
        uint32_t r = mix(gw * 9781u + (it * 4 + j) * 7919u + (lane >> 3) * 104729u);
        r = hot_rows ? (r % hot_rows) * (n_rows / hot_rows) : r % n_rows;
        float* row = buf + (size_t)r * 32 + (lane & 7) * 4;
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(row), "f"(1.f), "f"(2.f),
                     "f"(3.f), "f"(4.f) : "memory");
      }
    }
    if (MODE != 0) {
      float* stage = stage0 + (it % kStages) * (kRowsPerStage * kRowFloats);
      if (it >= kStages) {   // the bulk op that last read this stage must have finished reading
        if (lane < kRowsPerStage)
          asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kStages - 1) : "memory");
        __syncwarp();
      }
      // fill the stage: every lane writes 16 bytes per instruction
#pragma unroll
      for (int j = 0; j < kRowsPerStage * kRowFloats / 128; ++j)
        *reinterpret_cast<float4*>(stage + j * 128 + lane * 4) = make_float4(1.f, 2.f, 3.f, 4.f);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane < kRowsPerStage) {
        uint32_t r = mix(gw * 9781u + it * 7919u + lane * 104729u);
        r = hot_rows ? (r % hot_rows) * (n_rows / hot_rows) : r % n_rows;
        if (BYTES > 128) r = r / (BYTES / 128) * (BYTES / 128);
        float* row = buf + (size_t)r * 32;
        const uint32_t s = (uint32_t)__cvta_generic_to_shared(stage + lane * kRowFloats);
        if (MODE == 1 || MODE == 3)
          asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;"
                       ::"l"(row), "r"(s), "n"(BYTES) : "memory");
        else
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                       ::"l"(row), "r"(s), "n"(BYTES) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    }
  }
  if (MODE != 0 && lane < kRowsPerStage) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <int MODE, int BYTES>
void run(const char* name, float* buf, uint32_t n_rows, uint32_t hot, int blocks_per_sm, int sm_div = 1) {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  sms /= sm_div;     // sm_div > 1: only a fraction of the SMs issue (is the limit per SM or in L2?)
  const int block = 256, warps = block / 32;
  const size_t smem = (size_t)warps * kStages * kRowsPerStage * BYTES;
  cudaFuncSetAttribute(k<MODE, BYTES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const unsigned grid = sms * blocks_per_sm;
  const long rows_total = 1L << 25;
  const int rounds = (int)(rows_total / ((long)grid * warps * kRowsPerStage));
  const double rows_done = (double)rounds * grid * warps * kRowsPerStage * (BYTES / 128.0) * (MODE == 3 ? 2 : 1);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  k<MODE, BYTES><<<grid, block, smem>>>(buf, n_rows, hot, rounds);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%-40s CUDA error: %s\n", name, cudaGetErrorString(e)); return; }
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(a);
    k<MODE, BYTES><<<grid, block, smem>>>(buf, n_rows, hot, rounds);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
  }
  printf("%-40s hot=%-6u SMs=%3d blocks/SM=%d  %8.3f ms  %7.2f G rows(128B)/s  %8.1f GB/s payload  %5.2f clk/row/SM\n",
         name, hot, sms, blocks_per_sm, best, rows_done / best * 1e-6, rows_done * 128.0 / best * 1e-6,
         best * 1e-3 * 1.965e9 * sms / rows_done);
}

int main() {
  const uint32_t n_rows = 3u * 22223u * 8u;        // grad_value of config 2: rows of 32 floats
  float* buf;
  cudaMalloc(&buf, (size_t)n_rows * 128 + 4096);
  cudaMemset(buf, 0, (size_t)n_rows * 128 + 4096);
  for (uint32_t hot : {0u, 2184u}) {
    for (int bps : {1, 2, 4}) {
      run<0, 128>("red.global.add.v4.f32 (LSU)", buf, n_rows, hot, bps);
      run<1, 128>("cp.reduce.async.bulk add.f32 128 B", buf, n_rows, hot, bps);
      run<3, 128>("LSU red + TMA reduce together", buf, n_rows, hot, bps);
      run<2, 128>("cp.async.bulk store 128 B", buf, n_rows, hot, bps);
    }
  }
  // the same with half / a quarter of the SMs issuing: a per-SM limit halves the total, an L2-side
  // limit leaves it where it was
  for (int div : {2, 4}) {
    run<0, 128>("red.global.add.v4.f32 (LSU)", buf, n_rows, 0u, 2, div);
    run<1, 128>("cp.reduce.async.bulk add.f32 128 B", buf, n_rows, 0u, 2, div);
    run<3, 128>("LSU red + TMA reduce together", buf, n_rows, 0u, 2, div);
  }
  return 0;
}
