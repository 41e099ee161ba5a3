// layernorm.cu — LayerNorm over 256 channels, forward and backward, for the
// `norm` step that follows every attention module and feed-forward block of the
// reference's transformer layers (operation_order ('self_attn', 'norm', 'ffn', 'norm'),
// configs/videopose/2025-2-13/2025_2_13_res50_num_frames_3_posetrack17.py:57; mmcv builds
// it as nn.LayerNorm(256)).
//
// Pure streaming work: HBM-bound.  One warp owns a row (each lane 2 x float4 = 8 of the
// 256 channels, loads fully coalesced), statistics by warp shuffles, two-pass variance in
// registers.  The backward makes ONE pass over (x, dy): it writes dx and keeps the
// gamma / beta gradient partial sums of its 8 channels in registers across all the rows
// the warp visits, folds the block's warps through shared memory and reduces into the
// outputs with one red.global.add.v4.f32 per 4 channels and block (torch runs a second
// kernel over the same tensors for these two vectors, at ~0.3 TB/s).
#include <cuda_runtime.h>
#include <stdint.h>

#include "msda_kernels.h"

namespace msda {
namespace {

constexpr int kLnWidth = 256;
constexpr int kLnWarps = 8;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float sum8(const float4& a, const float4& b) {
  return (a.x + a.y) + (a.z + a.w) + (b.x + b.y) + (b.z + b.w);
}

__global__ void __launch_bounds__(kLnWarps * 32)
layernorm256_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                        const float* __restrict__ beta, float* __restrict__ y, float* __restrict__ mean_out,
                        float* __restrict__ rstd_out, int rows, float eps) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int stride = gridDim.x * kLnWarps;
  const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma) + lane);
  const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma) + 32 + lane);
  const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta) + lane);
  const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta) + 32 + lane);
  for (int r = blockIdx.x * kLnWarps + warp; r < rows; r += stride) {
    const float4* xr = reinterpret_cast<const float4*>(x + static_cast<int64_t>(r) * kLnWidth);
    const float4 a = __ldcs(xr + lane), b = __ldcs(xr + 32 + lane);
    const float mean = warp_sum(sum8(a, b)) * (1.f / kLnWidth);
    const float4 da = make_float4(a.x - mean, a.y - mean, a.z - mean, a.w - mean);
    const float4 db = make_float4(b.x - mean, b.y - mean, b.z - mean, b.w - mean);
    const float4 qa = make_float4(da.x * da.x, da.y * da.y, da.z * da.z, da.w * da.w);
    const float4 qb = make_float4(db.x * db.x, db.y * db.y, db.z * db.z, db.w * db.w);
    const float rstd = rsqrtf(warp_sum(sum8(qa, qb)) * (1.f / kLnWidth) + eps);
    float4* yr = reinterpret_cast<float4*>(y + static_cast<int64_t>(r) * kLnWidth);
    yr[lane] = make_float4(da.x * rstd * g0.x + b0.x, da.y * rstd * g0.y + b0.y, da.z * rstd * g0.z + b0.z,
                           da.w * rstd * g0.w + b0.w);
    yr[32 + lane] = make_float4(db.x * rstd * g1.x + b1.x, db.y * rstd * g1.y + b1.y,
                                db.z * rstd * g1.z + b1.z, db.w * rstd * g1.w + b1.w);
    if (lane == 0) {
      mean_out[r] = mean;
      rstd_out[r] = rstd;
    }
  }
}

__device__ __forceinline__ void fma4(float4& acc, const float4& a, const float4& b) {
  acc.x = fmaf(a.x, b.x, acc.x); acc.y = fmaf(a.y, b.y, acc.y);
  acc.z = fmaf(a.z, b.z, acc.z); acc.w = fmaf(a.w, b.w, acc.w);
}
__device__ __forceinline__ void add4(float4& acc, const float4& a) {
  acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w;
}
__device__ __forceinline__ void red4(float* dst, const float4& v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}

__global__ void __launch_bounds__(kLnWarps * 32)
layernorm256_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                        const float* __restrict__ gamma, const float* __restrict__ mean_in,
                        const float* __restrict__ rstd_in, float* __restrict__ dx,
                        float* __restrict__ dgamma, float* __restrict__ dbeta, int rows) {
  __shared__ float4 s_g[kLnWarps][64], s_b[kLnWarps][64];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int stride = gridDim.x * kLnWarps;
  const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma) + lane);
  const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma) + 32 + lane);
  float4 ag0 = make_float4(0.f, 0.f, 0.f, 0.f), ag1 = ag0, ab0 = ag0, ab1 = ag0;
  for (int r = blockIdx.x * kLnWarps + warp; r < rows; r += stride) {
    const int64_t off = static_cast<int64_t>(r) * kLnWidth;
    const float4* xr = reinterpret_cast<const float4*>(x + off);
    const float4* dr = reinterpret_cast<const float4*>(dy + off);
    const float4 a = __ldcs(xr + lane), b = __ldcs(xr + 32 + lane);
    const float4 da = __ldcs(dr + lane), db = __ldcs(dr + 32 + lane);
    const float mean = __ldg(mean_in + r), rstd = __ldg(rstd_in + r);
    const float4 ha = make_float4((a.x - mean) * rstd, (a.y - mean) * rstd, (a.z - mean) * rstd,
                                  (a.w - mean) * rstd);
    const float4 hb = make_float4((b.x - mean) * rstd, (b.y - mean) * rstd, (b.z - mean) * rstd,
                                  (b.w - mean) * rstd);
    const float4 ga = make_float4(da.x * g0.x, da.y * g0.y, da.z * g0.z, da.w * g0.w);
    const float4 gb = make_float4(db.x * g1.x, db.y * g1.y, db.z * g1.z, db.w * g1.w);
    const float c1 = warp_sum(sum8(ga, gb)) * (1.f / kLnWidth);
    const float4 pa = make_float4(ga.x * ha.x, ga.y * ha.y, ga.z * ha.z, ga.w * ha.w);
    const float4 pb = make_float4(gb.x * hb.x, gb.y * hb.y, gb.z * hb.z, gb.w * hb.w);
    const float c2 = warp_sum(sum8(pa, pb)) * (1.f / kLnWidth);
    float4* out = reinterpret_cast<float4*>(dx + off);
    out[lane] = make_float4(rstd * (ga.x - c1 - ha.x * c2), rstd * (ga.y - c1 - ha.y * c2),
                            rstd * (ga.z - c1 - ha.z * c2), rstd * (ga.w - c1 - ha.w * c2));
    out[32 + lane] = make_float4(rstd * (gb.x - c1 - hb.x * c2), rstd * (gb.y - c1 - hb.y * c2),
                                 rstd * (gb.z - c1 - hb.z * c2), rstd * (gb.w - c1 - hb.w * c2));
    fma4(ag0, da, ha);
    fma4(ag1, db, hb);
    add4(ab0, da);
    add4(ab1, db);
  }
  s_g[warp][lane] = ag0; s_g[warp][32 + lane] = ag1;
  s_b[warp][lane] = ab0; s_b[warp][32 + lane] = ab1;
  __syncthreads();
  if (threadIdx.x < 128) {                           // 64 float4 of dgamma, 64 of dbeta
    const int c = threadIdx.x & 63;
    const bool is_beta = threadIdx.x >= 64;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int w = 0; w < kLnWarps; ++w) add4(acc, is_beta ? s_b[w][c] : s_g[w][c]);
    red4((is_beta ? dbeta : dgamma) + 4 * c, acc);
  }
}

}  // namespace

bool layernorm_width_supported(int width) { return width == kLnWidth; }

static unsigned ln_grid(int rows, int sm_count) {
  const int blocks_needed = (rows + kLnWarps - 1) / kLnWarps;
  const int cap = sm_count * 8;                      // 8 blocks of 256 threads per SM: full occupancy
  return static_cast<unsigned>(blocks_needed < cap ? blocks_needed : cap);
}

cudaError_t launch_layernorm_forward(const float* x, const float* gamma, const float* beta, float* y,
                                     float* mean, float* rstd, int rows, float eps, int sm_count,
                                     cudaStream_t st) {
  layernorm256_fwd_kernel<<<ln_grid(rows, sm_count), kLnWarps * 32, 0, st>>>(x, gamma, beta, y, mean, rstd,
                                                                             rows, eps);
  note_launches(1);
  note_kernel(KF_LAYERNORM);
  return cudaGetLastError();
}

cudaError_t launch_layernorm_backward(const float* x, const float* dy, const float* gamma, const float* mean,
                                      const float* rstd, float* dx, float* dgamma, float* dbeta, int rows,
                                      int sm_count, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(dgamma, 0, sizeof(float) * kLnWidth, st);
  if (e != cudaSuccess) return e;
  e = cudaMemsetAsync(dbeta, 0, sizeof(float) * kLnWidth, st);
  if (e != cudaSuccess) return e;
  layernorm256_bwd_kernel<<<ln_grid(rows, sm_count), kLnWarps * 32, 0, st>>>(x, dy, gamma, mean, rstd, dx,
                                                                             dgamma, dbeta, rows);
  note_launches(1);
  note_kernel(KF_LAYERNORM);
  return cudaGetLastError();
}

}  // namespace msda
