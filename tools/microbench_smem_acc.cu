// microbench_smem_acc.cu — what does it cost an sm_100a SM to ACCUMULATE 128-byte fp32 rows in its
// own shared memory instead of sending them to the L2 as red.global.add.v4.f32?
//
// Access pattern of the large-Q backward kernel: a group of 8 lanes owns one 32-channel row, the 4
// groups of a warp hit 4 pseudo-random rows of a tile (the coarse pyramid levels of one head:
// 1 323 rows = 169 KB for 800x1333).  sm_100a has no floating-point shared-memory atomic (atomicAdd
// and red.shared.add.f32 compile to an ATOMS.CAST.SPIN loop), but it does have ATOMS.CAS.128, so a
// lane can retire its 16-byte slice of a row with one LDS.128 + 4 FADD + one 128-bit CAS.
//
// modes: 0 red.global.add.v4.f32 (the kernel's current path, rows spread over a 68 MB buffer)
//        1 shared ATOMS.CAS.128 loop          2 shared ATOMS.CAS.64 loop (two per lane)
//        3 atomicAdd(float) x4 (CAST.SPIN), bank-swizzled across the 4 groups
//        4 plain LDS.128 / FADD / STS.128 (racy: the upper bound of any ownership scheme)
//        5 integer ATOMS.ADD x4, bank-swizzled (what a fixed-point accumulator would cost)
//        6 half of the rows by mode 0, half by mode 1 (the mix a privatised backward would issue)
//        7 mode 0 plus one 128-byte value-row gather (ld.global.nc.v4.f32, 68 MB buffer) per reduction: the
//          backward kernel's real mix        8 mode 6 plus the same gathers
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench_smem_acc microbench_smem_acc.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t mix(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x;
}

__device__ __forceinline__ void smem_add_cas128(uint32_t saddr, float4 v) {
  float4 old;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(old.x), "=f"(old.y), "=f"(old.z), "=f"(old.w) : "r"(saddr));
  while (true) {
    const float4 nv = make_float4(old.x + v.x, old.y + v.y, old.z + v.z, old.w + v.w);
    unsigned long long c0 = (unsigned long long)__float_as_uint(old.x) | ((unsigned long long)__float_as_uint(old.y) << 32);
    unsigned long long c1 = (unsigned long long)__float_as_uint(old.z) | ((unsigned long long)__float_as_uint(old.w) << 32);
    unsigned long long n0 = (unsigned long long)__float_as_uint(nv.x) | ((unsigned long long)__float_as_uint(nv.y) << 32);
    unsigned long long n1 = (unsigned long long)__float_as_uint(nv.z) | ((unsigned long long)__float_as_uint(nv.w) << 32);
    unsigned long long o0, o1;
    asm volatile("{\n .reg .b128 c, v, o;\n mov.b128 c, {%3, %4};\n mov.b128 v, {%5, %6};\n"
                 " atom.shared.cas.b128 o, [%2], c, v;\n mov.b128 {%0, %1}, o;\n}\n"
                 : "=l"(o0), "=l"(o1) : "r"(saddr), "l"(c0), "l"(c1), "l"(n0), "l"(n1) : "memory");
    if (o0 == c0 && o1 == c1) break;
    old.x = __uint_as_float((uint32_t)o0); old.y = __uint_as_float((uint32_t)(o0 >> 32));
    old.z = __uint_as_float((uint32_t)o1); old.w = __uint_as_float((uint32_t)(o1 >> 32));
  }
}

__device__ __forceinline__ void smem_add_cas64(float* p, float a, float b) {
  unsigned long long* q = reinterpret_cast<unsigned long long*>(p);
  unsigned long long old = *q;
  while (true) {
    const float x = __uint_as_float((uint32_t)old) + a, y = __uint_as_float((uint32_t)(old >> 32)) + b;
    const unsigned long long nv = (unsigned long long)__float_as_uint(x) | ((unsigned long long)__float_as_uint(y) << 32);
    const unsigned long long seen = atomicCAS(q, old, nv);
    if (seen == old) break;
    old = seen;
  }
}

extern __shared__ float4 s_tile[];

template <int MODE>
__global__ void __launch_bounds__(1024, 1)
k(float* gbuf, const float* gbuf2, uint32_t g_rows, uint32_t tile_rows, int iters, float* check) {
  float acc = 0.f;
  float* tile = reinterpret_cast<float*>(s_tile);
  for (uint32_t i = threadIdx.x; i < tile_rows * 8; i += blockDim.x) s_tile[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();
  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(tile);
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t grp = tid >> 3, gl = tid & 7, g4 = (threadIdx.x >> 3) & 3;
  const float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
  for (int i = 0; i < iters; ++i) {
    const uint32_t h = mix(grp * 9781u + i * 7919u);
    const uint32_t r = __umulhi(h, tile_rows);
    bool to_global = MODE == 0 || MODE == 7 || ((MODE == 6 || MODE == 8) && (i & 1));
    if (MODE == 7 || MODE == 8) {
      const float4 lv = __ldg(reinterpret_cast<const float4*>(gbuf2 + (size_t)__umulhi(h * 0x9E3779B1u, g_rows) * 32 + gl * 4));
      acc += lv.x + lv.y + lv.z + lv.w;
    }
    if (to_global) {
      float* row = gbuf + (size_t)__umulhi(h * 2654435761u, g_rows) * 32 + gl * 4;
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(row), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
    } else if (MODE == 1 || MODE == 6 || MODE == 8) {
      smem_add_cas128(sbase + r * 128 + gl * 16, v);
    } else if (MODE == 2) {
      smem_add_cas64(tile + r * 32 + gl * 4, v.x, v.y);
      smem_add_cas64(tile + r * 32 + gl * 4 + 2, v.z, v.w);
    } else if (MODE == 3) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) atomicAdd(tile + r * 32 + gl * 4 + ((kk + g4) & 3), 1.f);
    } else if (MODE == 4) {
      float4* p = reinterpret_cast<float4*>(tile + r * 32 + gl * 4);
      float4 o = *p;
      o.x += v.x; o.y += v.y; o.z += v.z; o.w += v.w;
      *p = o;
    } else if (MODE == 5) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) atomicAdd(reinterpret_cast<int*>(tile) + r * 32 + gl * 4 + ((kk + g4) & 3), 1);
    }
  }
  __syncthreads();
  // checksum of the tile (modes 1, 2: must equal the number of updates x 10)
  float s = 0.f;
  for (uint32_t i = threadIdx.x; i < tile_rows * 32; i += blockDim.x) s += tile[i];
  atomicAdd(check + blockIdx.x, s);
  if (acc == 123.456f) check[1000] = acc;
}

template <int MODE>
void run(const char* name, float* gbuf, const float* gbuf2, uint32_t g_rows, uint32_t tile_rows, int threads, float* check, int sms) {
  const int iters = 1024;
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(tile_rows * 128));
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaMemset(check, 0, sms * 4);
  k<MODE><<<sms, threads, tile_rows * 128>>>(gbuf, gbuf2, g_rows, tile_rows, iters, check);
  cudaDeviceSynchronize();
  float h0 = 0; cudaMemcpy(&h0, check, 4, cudaMemcpyDeviceToHost);
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(a);
    k<MODE><<<sms, threads, tile_rows * 128>>>(gbuf, gbuf2, g_rows, tile_rows, iters, check);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
  }
  const double rows_sm = (double)(threads / 8) * iters;
  int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double expect = rows_sm * 80.0 * ((MODE == 6 || MODE == 8) ? 0.5 : 1.0);
  printf("%-44s tile=%4u rows thr=%4d  %7.3f ms  %6.2f Grows/s  %5.2f clk/row/SM  checksum %s\n", name, tile_rows,
         threads, best, rows_sm * sms / best * 1e-6, best * 1e-3 * khz * 1e3 / rows_sm,
         (MODE == 1 || MODE == 2 || MODE == 6 || MODE == 8) ? (h0 == (float)expect ? "ok" : "MISMATCH") : "-");
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) printf("  CUDA error: %s\n", cudaGetErrorString(e));
}

int main() {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const uint32_t g_rows = 3u * 22223u * 8u;
  float* gbuf; float* check;
  cudaMalloc(&gbuf, (size_t)g_rows * 128); cudaMalloc(&check, 8192);
  cudaMemset(gbuf, 0, (size_t)g_rows * 128);
  float* gbuf2; cudaMalloc(&gbuf2, (size_t)g_rows * 128); cudaMemset(gbuf2, 0, (size_t)g_rows * 128);
  for (int threads : {512, 1024}) {
    for (uint32_t tile_rows : {273u, 1323u}) {
      run<0>("red.global.add.v4.f32", gbuf, gbuf2, g_rows, tile_rows, threads, check, sms);
      run<1>("shared CAS.128 loop", gbuf, gbuf2, g_rows, tile_rows, threads, check, sms);
      run<2>("shared CAS.64 loop x2", gbuf, gbuf2, g_rows, tile_rows, threads, check, sms);
      run<3>("shared atomicAdd(float) x4 swizzled", gbuf, gbuf2, g_rows, tile_rows, threads, check, sms);
      run<4>("shared LDS.128+STS.128 (racy bound)", gbuf, gbuf2, g_rows, tile_rows, threads, check, sms);
      run<5>("shared ATOMS.ADD int x4 swizzled", gbuf, gbuf2, g_rows, tile_rows, threads, check, sms);
      run<6>("half red.global, half shared CAS.128", gbuf, gbuf2, g_rows, tile_rows, threads, check, sms);
      run<7>("gather + red.global (the kernel's mix)", gbuf, gbuf2, g_rows, tile_rows, threads, check, sms);
      run<8>("gather + half red.global, half CAS.128", gbuf, gbuf2, g_rows, tile_rows, threads, check, sms);
    }
  }
  return 0;
}
