// msda_bwd_io.cuh — pieces shared by the backward kernels (msda_bwd.cu rows / generic,
// msda_flat.cu flat): vector reductions into global memory, the lane layout of a row in
// the backward, the transposing shuffle reduction, and where per-sample gradients go.
#pragma once

#include "msda_common.cuh"

namespace msda {

// ---- vector reductions into global memory --------------------------------
__device__ __forceinline__ void red_add_row(float* p, const float (&v)[4]) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v[0]), "f"(v[1]),
               "f"(v[2]), "f"(v[3])
               : "memory");
}
// How the lanes of a group cover a row in the backward: the width of the GRADIENT element
// decides.  Under bf16 value storage with fp32 gradients a lane takes 4 channels - an 8-byte
// value load and ONE 16-byte reduction - so a row leaves the SM as one 128-byte request, like
// the fp32 kernel.  (Measured on B200: 64-byte reduction requests reach only ~80 % of the byte
// rate of 128-byte ones; covering the row with 4 lanes x two 64-byte halves ran config 2 in
// 0.786 ms against 0.642 ms for fp32 values.)
template <typename VT, typename GT>
struct BwdVec : Vec16<VT> {};
template <>
struct BwdVec<__nv_bfloat16, float> {
  static constexpr int VEC = 4;
  __device__ __forceinline__ static void load(const __nv_bfloat16* p, float (&v)[4]) {
    const uint2 t = __ldg(reinterpret_cast<const uint2*>(p));
    v[0] = __uint_as_float(t.x << 16); v[1] = __uint_as_float(t.x & 0xffff0000u);
    v[2] = __uint_as_float(t.y << 16); v[3] = __uint_as_float(t.y & 0xffff0000u);
  }
};

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  const __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&t);
}
__device__ __forceinline__ void red_add_row(__nv_bfloat16* p, const float (&v)[8]) {
  asm volatile("red.global.add.noftz.v4.bf16x2 [%0], {%1, %2, %3, %4};" ::"l"(p),
               "r"(pack_bf16x2(v[0], v[1])), "r"(pack_bf16x2(v[2], v[3])),
               "r"(pack_bf16x2(v[4], v[5])), "r"(pack_bf16x2(v[6], v[7]))
               : "memory");
}
__device__ __forceinline__ void red_add_row(__nv_bfloat16* p, const float (&v)[4]) {
  asm volatile("red.global.add.noftz.v2.bf16x2 [%0], {%1, %2};" ::"l"(p),
               "r"(pack_bf16x2(v[0], v[1])), "r"(pack_bf16x2(v[2], v[3]))
               : "memory");
}

// Sum p[j] over the G lanes of a group; lane gl ends up with the total of
// sample j == gl.  log2(G) rounds, G-1 shuffles in all.
template <int G>
__device__ __forceinline__ float group_transpose_reduce(float (&p)[G], int gl) {
#pragma unroll
  for (int half = G / 2; half >= 1; half >>= 1) {
    const bool upper = (gl & half) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const float send = upper ? p[i] : p[i + half];
      const float keep = upper ? p[i + half] : p[i];
      p[i] = keep + __shfl_xor_sync(0xffffffffu, send, half);
    }
  }
  return p[0];
}

// ---- where the per-sample gradients go -------------------------------------
// PlainIO: the reference op's outputs, grad_sampling_loc and grad_attn_weight.
// FusedIO: gradients of the raw projections — grad_offsets = grad_loc * scale
// and the softmax backward
//   grad_logit_s = w_s * (gw_s - sum_t w_t gw_t)
// (what autograd would compute through softmax and the location transform,
// multi_scale_deform_attn.py:375-393), optionally grad_loc for callers that
// need reference-point gradients.  The row-wide term needs no reduction over
// the samples: gw_t = <grad_out[row], sampled_t>, so
//   sum_t w_t gw_t = <grad_out[row], sum_t w_t sampled_t> = <grad_out[row], out[row]>
// with out the forward's own output, 128 bytes per row.
struct PlainIO {
  PlainSource src;
  float* grad_loc;
  float* grad_aw;
  static constexpr bool kFused = false;
  __device__ __forceinline__ void bind(int64_t unit, int LP, int M, int64_t bq) {
    src.bind(unit, LP, M, bq);     // grad_loc / grad_aw stay untouched (constant bank): src.so indexes them
  }
  template <int G, int VEC>
  __device__ __forceinline__ void row_dot(const float (&)[VEC], int64_t, int, int) {}
  // gw: d/d(attention weight); (tx, ty): d/d(pixel coordinate), so d/d(location) = (W tx, H ty)
  __device__ __forceinline__ void store(int s, int, float gw, float tx, float ty, float, float Wf,
                                        float Hf) {
    const int64_t e = src.so + s;
    __stcs(grad_aw + e, gw);
    __stcs(reinterpret_cast<float2*>(grad_loc + 2 * e), make_float2(Wf * tx, Hf * ty));
  }
};

struct FusedIO {
  FusedSource src;
  float* grad_off;    // (B,Q,M,L,P,2)
  float* grad_logit;  // (B,Q,M,L*P)
  float* grad_loc;    // optional (B,Q,M,L,P,2), NULL if reference points need no gradient
  const float* out;   // (B,Q,M*D) the forward's output
  float dot;          // <grad_out[row], out[row]>
  static constexpr bool kFused = true;
  __device__ __forceinline__ void bind(int64_t unit, int LP, int M, int64_t bq) {
    src.bind(unit, LP, M, bq);     // the gradient pointers stay untouched (constant bank): src.so indexes them
    dot = 0.f;
  }
  // g: this lane's VEC channels (gl*VEC ...) of grad_out[row]; G lanes cover the row
  template <int G, int VEC>
  __device__ __forceinline__ void row_dot(const float (&g)[VEC], int64_t row, int D, int gl) {
    const float* o = out + row * D + gl * VEC;
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < VEC; i += 4) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(o + i));
      t += g[i] * v.x + g[i + 1] * v.y + g[i + 2] * v.z + g[i + 3] * v.w;
    }
#pragma unroll
    for (int o2 = G / 2; o2 >= 1; o2 >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o2);
    dot = t;
  }
  __device__ __forceinline__ void store(int s, int l, float gw, float tx, float ty, float w,
                                        float Wf, float Hf) {
    float2 go;
    if (src.scale) {
      const float2 sc = __ldg(reinterpret_cast<const float2*>(src.scale) + (src.ro + l));
      go = make_float2(Wf * tx * sc.x, Hf * ty * sc.y);
    } else {
      go = make_float2(tx, ty);   // d loc / d off = 1 / (W, H) cancels the pixel scale
    }
    const int64_t e = src.so + s;
    __stcs(reinterpret_cast<float2*>(grad_off + 2 * e), go);
    if (grad_loc) __stcs(reinterpret_cast<float2*>(grad_loc + 2 * e), make_float2(Wf * tx, Hf * ty));
    __stcs(grad_logit + e, w * (gw - dot));
  }
};

}  // namespace msda
