// msda_common.cuh — shared device helpers for the sm_100a multi-scale
// deformable attention kernels.
//
// Semantics follow the reference op (SURVEY.md Appendix A;
// third_party/mmcv/mmcv/ops/csrc/common/cuda/ms_deform_attn_cuda_kernel.cuh:17-64,
// 66-131, 200-254): pixel coordinates h = y*H - 0.5, w = x*W - 0.5, a sample
// contributes only if -1 < h < H and -1 < w < W, and each of the four
// bilinear corners contributes only if it lies inside the map.
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace msda {

struct Dims {
  int B;  // batch (frames, or clips for the fused multi-frame view)
  int S;  // keys per batch entry = sum_l H_l*W_l
  int M;  // heads
  int D;  // channels per head
  int L;  // levels
  int Q;  // queries
  int P;  // points per level
};

constexpr int kMaxSmemLevels = 64;  // level table cached in shared memory up to this many levels

// One level of the pyramid as the kernels want it: all int32, pre-multiplied.
struct LevelInfo {
  int H;
  int W;
  int start;       // level_start_index[l]
  int row_stride;  // W * M * D  (elements between vertically adjacent pixels)
};

__device__ __forceinline__ LevelInfo load_level(const int64_t* __restrict__ shapes,
                                                const int64_t* __restrict__ lsi, int l,
                                                int MD) {
  LevelInfo li;
  li.H = static_cast<int>(shapes[2 * l]);
  li.W = static_cast<int>(shapes[2 * l + 1]);
  li.start = static_cast<int>(lsi[l]);
  li.row_stride = li.W * MD;
  return li;
}

// ---- 16-byte row-segment loads -------------------------------------------
// A "row" is the D channels of one head at one pixel.  Each lane owns VEC
// consecutive channels = one 16-byte load.

template <typename VT>
struct Vec16;

template <>
struct Vec16<float> {
  static constexpr int VEC = 4;
  __device__ __forceinline__ static void load(const float* p, float (&v)[4]) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
};

template <>
struct Vec16<__nv_bfloat16> {
  static constexpr int VEC = 8;
  __device__ __forceinline__ static void load(const __nv_bfloat16* p, float (&v)[8]) {
    const uint4 t = __ldg(reinterpret_cast<const uint4*>(p));
    // bf16 -> f32 is a 16-bit left shift
    v[0] = __uint_as_float(t.x << 16); v[1] = __uint_as_float(t.x & 0xffff0000u);
    v[2] = __uint_as_float(t.y << 16); v[3] = __uint_as_float(t.y & 0xffff0000u);
    v[4] = __uint_as_float(t.z << 16); v[5] = __uint_as_float(t.z & 0xffff0000u);
    v[6] = __uint_as_float(t.w << 16); v[7] = __uint_as_float(t.w & 0xffff0000u);
  }
};

// Streaming (read-once) loads for locations / weights / grad_output: keep them
// from displacing value rows in L1.
__device__ __forceinline__ float2 ld_stream_f2(const float* p) {
  return __ldcs(reinterpret_cast<const float2*>(p));
}
__device__ __forceinline__ float ld_stream_f(const float* p) { return __ldcs(p); }

// ---- sample record -------------------------------------------------------
// What the lane that "owns" a sample computes once and the G lanes of the
// row group consume.  32 bytes in shared memory.
struct __align__(16) SampleRec {
  int off00;   // element offset (within the batch entry) of corner (h0, w0), head 0, channel 0
  int meta;    // bits 0..3 corner validity (v1,v2,v3,v4), bits 4.. level index
  float lh;    // h - h0
  float lw;    // w - w0
  float a;     // attention weight (0 when the sample is outside the map)
  int rs;      // row stride of the sample's level, W*M*D elements
  int pad0, pad1;
};

// Corner order matches the reference: v1=(h0,w0) v2=(h0,w1) v3=(h1,w0) v4=(h1,w1).
__device__ __forceinline__ void make_sample(float x, float y, float& a, const LevelInfo& lv,
                                            int level, int MD, int& off00, int& meta,
                                            float& lh, float& lw) {
  const float h_im = y * static_cast<float>(lv.H) - 0.5f;
  const float w_im = x * static_cast<float>(lv.W) - 0.5f;
  off00 = 0; meta = 0; lh = 0.f; lw = 0.f;
  if (h_im > -1.f && w_im > -1.f && h_im < static_cast<float>(lv.H) &&
      w_im < static_cast<float>(lv.W)) {
    const float hf = floorf(h_im), wf = floorf(w_im);
    const int h0 = static_cast<int>(hf), w0 = static_cast<int>(wf);
    lh = h_im - hf;
    lw = w_im - wf;
    const int r0 = h0 >= 0, r1 = h0 + 1 <= lv.H - 1;
    const int c0 = w0 >= 0, c1 = w0 + 1 <= lv.W - 1;
    meta = (r0 & c0) | ((r0 & c1) << 1) | ((r1 & c0) << 2) | ((r1 & c1) << 3) | (level << 4);
    off00 = (lv.start + h0 * lv.W + w0) * MD;
  } else {
    a = 0.f;  // out-of-range samples contribute nothing, whatever their weight (even NaN)
  }
}

}  // namespace msda
