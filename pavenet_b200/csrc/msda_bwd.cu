// msda_bwd.cu — backward kernels of multi-scale deformable attention sampling
// for sm_100a.
//
// Maths (reference: ms_deform_attn_cuda_kernel.cuh:66-131 inside the loop of
// :256-345), per sample that passes the range test, g = grad_out[b,q,m,c]:
//   grad_value[corner_i]      += w_i * g * a                 (valid corners only)
//   grad_attn_weight[b,q,m,l,p] = sum_c g * (w1 v1 + w2 v2 + w3 v3 + w4 v4)
//   grad_loc[...,0]             = W * sum_c (hh (v2-v1) + lh (v4-v3)) * g * a
//   grad_loc[...,1]             = H * sum_c (hw (v3-v1) + lw (v4-v2)) * g * a
// Samples outside the range get zero location / weight gradients.
//
// rows<D,VT,GT>: same lane-group-per-row layout as the forward.  The value
// gradient leaves each lane as ONE 16-byte vector reduction per corner
// (red.global.add.v4.f32, or v4.bf16x2 when the gradient is stored in bf16)
// instead of the reference's 4 scalar atomics per channel, and the
// per-sample channel sums are finished with a transposing shuffle reduction
// over the G lanes of the group (7 shuffles per 8 samples per quantity)
// instead of a shared-memory pass with thread 0 summing serially
// (ms_deform_attn_cuda_kernel.cuh:319-336).
#include <type_traits>

#include "msda_kernels.h"
#include "msda_bwd_io.cuh"

namespace msda {

constexpr int kRowsThreads = 256;
constexpr int kRowsWarps = kRowsThreads / 32;

// "Warp-aggregated atomics" (variant 2 of the large-Q backward, BASELINE.json north_star): the
// NG lane groups of a warp process the same sample index of NG consecutive queries; when two
// of them are about to reduce into the SAME value row (same pixel, same head) their weighted
// rows are added with shuffles and only the lowest group issues the red.  match.any finds the
// coincidences; the shuffle pass runs only in steps where the warp has one.
template <int G, int VEC, typename GT>
__device__ __forceinline__ void scatter_corner_aggregated(bool valid, int row_off, GT* ptr,
                                                          const float (&t)[VEC], int grp, int gl) {
  constexpr int NG = 32 / G;
  const int key = valid ? row_off : (-1 - grp);          // invalid corners match nobody
  const unsigned same = __match_any_sync(0xffffffffu, key);
  if (__any_sync(0xffffffffu, __popc(same) > G)) {
    float sum[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) sum[i] = t[i];
#pragma unroll
    for (int dgrp = 1; dgrp < NG; ++dgrp) {
      const int partner = ((grp + dgrp) % NG) * G + gl;
      const int pkey = __shfl_sync(0xffffffffu, key, partner);
#pragma unroll
      for (int i = 0; i < VEC; ++i) {
        const float pv = __shfl_sync(0xffffffffu, t[i], partner);
        sum[i] += (pkey == key) ? pv : 0.f;
      }
    }
    if (valid && (__ffs(same) - 1) / G == grp) red_add_row(ptr, sum);
  } else if (valid) {
    red_add_row(ptr, t);
  }
}

// resident blocks per SM the fused-prologue instantiation is compiled for (its IO object carries ~16 more
// live registers than the plain one, which fits three blocks at 80 registers by itself)
#ifndef MSDA_FUSED_BWD_MINB
#define MSDA_FUSED_BWD_MINB 2
#endif
#ifndef MSDA_BWD_D64_MINB
#define MSDA_BWD_D64_MINB 2
#endif

// (plain, bf16 value with fp32 gradients, D <= 32: three resident blocks = 80 registers stated explicitly -- left to
// itself the compiler gave this instantiation 93 registers after an unrelated change and config 2's bf16 backward went
// from 0.643 to 0.713 ms.  The fp32 instantiation lands on 80 registers, no spills, by itself and is left alone.)
template <typename VT, typename GT, class IO, int D, int AGG>
constexpr int bwd_rows_min_blocks() {
  if (IO::kFused) return MSDA_FUSED_BWD_MINB;
  if (D == 64 && AGG == 0) return MSDA_BWD_D64_MINB;   // 130-144 registers unbounded = one block per SM; 128 fits two
  return (!std::is_same<VT, float>::value && std::is_same<GT, float>::value && D <= 32 && AGG == 0) ? 3 : 0;
}

template <int D, typename VT, typename GT, class IO, int AGG = 0>
__global__ void __launch_bounds__(kRowsThreads, bwd_rows_min_blocks<VT, GT, IO, D, AGG>())
msda_bwd_rows_kernel(const VT* __restrict__ value, const int64_t* __restrict__ shapes,
                     const int64_t* __restrict__ lsi, IO io, const float* __restrict__ grad_out,
                     GT* __restrict__ grad_value, Dims d, int nsplit, int agg_min_level) {
  using VL = BwdVec<VT, GT>;
  constexpr int VEC = VL::VEC;
  constexpr int G = D / VEC;
  static_assert(D % VEC == 0 && G >= 1 && G <= 32 && (G & (G - 1)) == 0, "bad D");

  __shared__ LevelInfo s_lvl[kMaxSmemLevels];
  __shared__ int4 s_board[kRowsWarps][G * (2 * (32 / G) + 1)];

  const int MD = d.M * D;
  for (int l = threadIdx.x; l < d.L; l += blockDim.x) s_lvl[l] = load_level(shapes, lsi, l, MD);
  __syncthreads();

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int gl = lane & (G - 1);
  const int grp = lane / G;

  // same block -> (batch entry, query chunk, head) mapping as the forward
  constexpr int GPB = kRowsThreads / G;   // row groups per block
  const int qpb = GPB / nsplit;           // queries per block (nsplit divides GPB)
  const int n_chunks = (d.Q + qpb - 1) / qpb;
  int blk = blockIdx.x;
  const int m = blk % d.M;
  blk /= d.M;
  const int chunk = blk % n_chunks;
  const int64_t b = blk / n_chunks;
  const int gib = threadIdx.x / G;
  const int split = gib % nsplit;
  int q_idx = chunk * qpb + gib / nsplit;
  const bool live = q_idx < d.Q;
  if (!live) q_idx = d.Q - 1;
  const int64_t unit = (b * d.Q + q_idx) * d.M + m;

  const int64_t boff = b * d.S * MD + m * D + gl * VEC;
  const VT* vbase = value + boff;
  GT* gvbase = grad_value + boff;
  const int LP = d.L * d.P;
  io.bind(unit, LP, d.M, b * d.Q + q_idx);
  io.src.template prepass<G>(LP, gl, false);   // fused: row max and 1/sum saved by the forward

  float g[VEC];
  {
    const float* gp = grad_out + unit * D + gl * VEC;
#pragma unroll
    for (int c = 0; c < VEC; c += 4) {
      const float4 t = __ldcs(reinterpret_cast<const float4*>(gp + c));
      g[c] = t.x; g[c + 1] = t.y; g[c + 2] = t.z; g[c + 3] = t.w;
    }
  }
  io.template row_dot<G, VEC>(g, unit, D, gl);   // fused: <grad_out[row], out[row]>

  const int per = ((LP + nsplit - 1) / nsplit + G - 1) / G * G;
  const int s_begin = split * per;
  const int s_end = live ? min(LP, s_begin + per) : 0;  // dead groups touch nothing

  // per-warp record board, same conflict-free layout as the forward kernel
  constexpr int NG = 32 / G;
  int4* board = s_board[warp];
  auto unit_of = [](int j, int grp_, int half) { return j * (2 * NG + 1) + 2 * grp_ + half; };
  const FastDivP level_of(d.P);

  RawSample nxt;
  nxt.x = nxt.y = nxt.w = 0.f;
  {
    const int s = s_begin + gl;
    if (s < s_end) nxt = io.src.load(s);
  }
  for (int s0 = s_begin; s0 < s_begin + per; s0 += G) {
    const int s = s0 + gl;
    float Wf = 0.f, Hf = 0.f, w_true = 0.f;
    int lvl = 0;
    {
      SampleRec r;
      r.off00 = 0; r.meta = 0; r.lh = 0.f; r.lw = 0.f; r.a = 0.f; r.rs = 0;
      if (s < s_end) {
        lvl = level_of(s);
        const LevelInfo lv = s_lvl[lvl];
        RawSample cur = nxt;
        io.src.finish(cur, s, lvl, lv);
        w_true = cur.w;
        r.a = cur.w;
        Wf = static_cast<float>(lv.W);
        Hf = static_cast<float>(lv.H);
        r.rs = lv.row_stride;
        make_sample(cur.x, cur.y, r.a, lv, lvl, MD, r.off00, r.meta, r.lh, r.lw);
      }
      board[unit_of(gl, grp, 0)] =
          make_int4(r.off00, r.meta, __float_as_int(r.lh), __float_as_int(r.lw));
      *reinterpret_cast<int2*>(&board[unit_of(gl, grp, 1)]) = make_int2(__float_as_int(r.a), r.rs);
      const int sn = s + G;
      if (sn < s_end) nxt = io.src.load(sn);
    }
    __syncwarp();

    float pw[G], px[G], py[G];
#pragma unroll
    for (int j = 0; j < G; ++j) {
      const int4 q = board[unit_of(j, grp, 0)];
      const int meta = q.y;
      {
        const int2 ar = *reinterpret_cast<const int2*>(&board[unit_of(j, grp, 1)]);
      const float a = __int_as_float(ar.x);  // 0 for samples outside the map
        const float lh = __int_as_float(q.z), lw = __int_as_float(q.w);
        const float hh = 1.f - lh, hw = 1.f - lw;
        const int rs = ar.y;
        const VT* p = vbase + q.x;
        GT* gp = gvbase + q.x;
        float v1[VEC], v2[VEC], v3[VEC], v4[VEC];
#pragma unroll
        for (int c = 0; c < VEC; ++c) v1[c] = v2[c] = v3[c] = v4[c] = 0.f;
        if (meta & 1) VL::load(p, v1);
        if (meta & 2) VL::load(p + MD, v2);
        if (meta & 4) VL::load(p + rs, v3);
        if (meta & 8) VL::load(p + rs + MD, v4);
        const float w1 = hh * hw, w2 = hh * lw, w3 = lh * hw, w4 = lh * lw;
        float tg[VEC];
        float sw = 0.f, sx = 0.f, sy = 0.f;
#pragma unroll
        for (int c = 0; c < VEC; ++c) {
          tg[c] = g[c] * a;
          const float val = w1 * v1[c] + w2 * v2[c] + w3 * v3[c] + w4 * v4[c];
          const float gw = hh * (v2[c] - v1[c]) + lh * (v4[c] - v3[c]);
          const float gh = hw * (v3[c] - v1[c]) + lw * (v4[c] - v2[c]);
          sw += g[c] * val;
          sx += gw * tg[c];
          sy += gh * tg[c];
        }
        pw[j] = sw; px[j] = sx; py[j] = sy;
        float t[VEC];
#define MSDA_SCATTER(BIT, WK, PTR)                                   \
  if (meta & BIT) {                                                  \
    _Pragma("unroll") for (int c = 0; c < VEC; ++c) t[c] = (WK) * a * g[c]; \
    red_add_row(PTR, t);                                             \
  }
        // the level of sample s0 + j is the same in every group of the warp (nsplit == 1), dead
        // groups included, so the warp-wide match / shuffles below are executed convergently
        if (AGG && level_of(s0 + j) >= agg_min_level) {
#define MSDA_SCATTER_AGG(BIT, WK, OFF)                                                   \
  {                                                                                      \
    _Pragma("unroll") for (int c = 0; c < VEC; ++c) t[c] = (WK) * a * g[c];              \
    scatter_corner_aggregated<G, VEC, GT>((meta & BIT) != 0, q.x + (OFF), gp + (OFF), t, grp, gl); \
  }
          MSDA_SCATTER_AGG(1, w1, 0)
          MSDA_SCATTER_AGG(2, w2, MD)
          MSDA_SCATTER_AGG(4, w3, rs)
          MSDA_SCATTER_AGG(8, w4, rs + MD)
#undef MSDA_SCATTER_AGG
        } else {
          MSDA_SCATTER(1, w1, gp)
          MSDA_SCATTER(2, w2, gp + MD)
          MSDA_SCATTER(4, w3, gp + rs)
          MSDA_SCATTER(8, w4, gp + rs + MD)
        }
#undef MSDA_SCATTER
      }
    }
    // lane gl receives the channel-summed gradients of sample s0 + gl
    const float tw = group_transpose_reduce<G>(pw, gl);
    const float tx = group_transpose_reduce<G>(px, gl);
    const float ty = group_transpose_reduce<G>(py, gl);
    if (s < s_end) io.store(s, lvl, tw, tx, ty, w_true, Wf, Hf);
    __syncwarp();
  }
}

// --------------------------------------------------------------------------
// generic kernel: one block per (b,q,m) row, threads stride the channels
// --------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T to_T(float v) { return static_cast<T>(v); }
template <typename T>
__device__ __forceinline__ T to_T(double v) { return static_cast<T>(v); }
template <typename T>
__device__ __forceinline__ T to_T(__nv_bfloat16 v) { return static_cast<T>(__bfloat162float(v)); }

__device__ __forceinline__ void atomic_add_any(float* p, float v) { atomicAdd(p, v); }
__device__ __forceinline__ void atomic_add_any(double* p, double v) { atomicAdd(p, v); }
__device__ __forceinline__ void atomic_add_any(__nv_bfloat16* p, float v) {
  atomicAdd(p, __float2bfloat16_rn(v));
}

constexpr int kGenericBwdThreads = 128;

template <typename T>
__device__ __forceinline__ T block_sum(T v, T* s_part) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) s_part[warp] = v;
  __syncthreads();
  T tot = 0;
#pragma unroll
  for (int w = 0; w < kGenericBwdThreads / 32; ++w) tot += s_part[w];
  __syncthreads();
  return tot;
}

template <typename T, typename VT, typename GT>
__global__ void __launch_bounds__(kGenericBwdThreads)
msda_bwd_generic_kernel(const VT* __restrict__ value, const int64_t* __restrict__ shapes,
                        const int64_t* __restrict__ lsi, const T* __restrict__ loc,
                        const T* __restrict__ aw, const T* __restrict__ grad_out,
                        GT* __restrict__ grad_value, T* __restrict__ grad_loc,
                        T* __restrict__ grad_aw, Dims d) {
  __shared__ T s_part[kGenericBwdThreads / 32];
  const int64_t n_units = static_cast<int64_t>(d.B) * d.Q * d.M;
  const int64_t MD = static_cast<int64_t>(d.M) * d.D;
  for (int64_t unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
    const int m = static_cast<int>(unit % d.M);
    const int64_t b = unit / d.M / d.Q;
    const int64_t boff = b * d.S * MD + static_cast<int64_t>(m) * d.D;
    const T* go = grad_out + unit * d.D;
    const int LP = d.L * d.P;
    for (int l = 0; l < d.L; ++l) {
      const int H = static_cast<int>(shapes[2 * l]);
      const int W = static_cast<int>(shapes[2 * l + 1]);
      const int64_t loff = boff + lsi[l] * MD;
      for (int p = 0; p < d.P; ++p) {
        const int64_t si = unit * LP + static_cast<int64_t>(l) * d.P + p;
        const T x = loc[2 * si], y = loc[2 * si + 1], a = aw[si];
        const T h_im = y * H - T(0.5), w_im = x * W - T(0.5);
        T sw = 0, sx = 0, sy = 0;
        const bool inside = h_im > T(-1) && w_im > T(-1) && h_im < T(H) && w_im < T(W);
        if (inside) {
          const T hf = floor(h_im), wf = floor(w_im);
          const int h0 = static_cast<int>(hf), w0 = static_cast<int>(wf);
          const T lh = h_im - hf, lw = w_im - wf, hh = T(1) - lh, hw = T(1) - lw;
          const bool r0 = h0 >= 0, r1 = h0 + 1 <= H - 1, c0 = w0 >= 0, c1 = w0 + 1 <= W - 1;
          const int64_t o00 = loff + (static_cast<int64_t>(h0) * W + w0) * MD;
          const int64_t o01 = o00 + MD, o10 = o00 + W * MD, o11 = o10 + MD;
          const T w1 = hh * hw, w2 = hh * lw, w3 = lh * hw, w4 = lh * lw;
          for (int c = threadIdx.x; c < d.D; c += kGenericBwdThreads) {
            const T gc = go[c];
            const T tg = gc * a;
            const T v1 = (r0 && c0) ? to_T<T>(value[o00 + c]) : T(0);
            const T v2 = (r0 && c1) ? to_T<T>(value[o01 + c]) : T(0);
            const T v3 = (r1 && c0) ? to_T<T>(value[o10 + c]) : T(0);
            const T v4 = (r1 && c1) ? to_T<T>(value[o11 + c]) : T(0);
            if (r0 && c0) atomic_add_any(grad_value + o00 + c, w1 * tg);
            if (r0 && c1) atomic_add_any(grad_value + o01 + c, w2 * tg);
            if (r1 && c0) atomic_add_any(grad_value + o10 + c, w3 * tg);
            if (r1 && c1) atomic_add_any(grad_value + o11 + c, w4 * tg);
            sw += gc * (w1 * v1 + w2 * v2 + w3 * v3 + w4 * v4);
            sx += (hh * (v2 - v1) + lh * (v4 - v3)) * tg;
            sy += (hw * (v3 - v1) + lw * (v4 - v2)) * tg;
          }
        }
        // `inside` is block-uniform, so the barriers inside block_sum are safe
        sw = block_sum(sw, s_part);
        sx = block_sum(sx, s_part);
        sy = block_sum(sy, s_part);
        if (threadIdx.x == 0) {
          grad_aw[si] = sw;
          grad_loc[2 * si] = T(W) * sx;
          grad_loc[2 * si + 1] = T(H) * sy;
        }
      }
    }
  }
}

// --------------------------------------------------------------------------
// launchers
// --------------------------------------------------------------------------
template <int D, typename VT, typename GT, class IO>
static cudaError_t launch_bwd_rows(const void* value, const int64_t* shapes, const int64_t* lsi,
                                   const IO& io, const float* go, void* gv, const Dims& d,
                                   int nsplit, cudaStream_t st) {
  constexpr int G = D / BwdVec<VT, GT>::VEC;
  constexpr int GPB = kRowsThreads / G;
  if (nsplit > GPB) nsplit = GPB;   // (a power of two, so it divides GPB)
  const int qpb = GPB / nsplit;
  const int64_t blocks = static_cast<int64_t>(d.B) * ((d.Q + qpb - 1) / qpb) * d.M;
  if (blocks >= (int64_t(1) << 31)) return cudaErrorInvalidConfiguration;
  // variant 2: warp-aggregated atomics (the benchmark's kernel only; rows not split)
  if constexpr (D == 32 && std::is_same<VT, float>::value && std::is_same<GT, float>::value &&
                !IO::kFused) {
    if (tuning().bwd_variant == 2 && nsplit == 1) {
      msda_bwd_rows_kernel<D, VT, GT, IO, 1><<<static_cast<unsigned>(blocks), kRowsThreads, 0, st>>>(
          static_cast<const VT*>(value), shapes, lsi, io, go, static_cast<GT*>(gv), d, nsplit,
          tuning().agg_min_level);
      note_launches(1);
      note_kernel(KF_BWD_ROWS);
      return cudaGetLastError();
    }
  }
  msda_bwd_rows_kernel<D, VT, GT, IO><<<static_cast<unsigned>(blocks), kRowsThreads, 0, st>>>(
      static_cast<const VT*>(value), shapes, lsi, io, go, static_cast<GT*>(gv), d, nsplit, 0);
  note_launches(1);
  note_kernel(IO::kFused ? KF_BWD_ROWS_FUSED : KF_BWD_ROWS);
  return cudaGetLastError();
}

static int choose_bwd_split(const Dims& d, int G, int sm_count) {
  if (tuning().bwd_split > 0) {
    int s = 1;
    while (s * 2 <= tuning().bwd_split && s < 32) s *= 2;  // power of two
    return s;
  }
  const int64_t units = static_cast<int64_t>(d.B) * d.Q * d.M;
  const int64_t want_groups = static_cast<int64_t>(sm_count) * 64 * (32 / G);
  const int LP = d.L * d.P;
  int split = 1;
  while (split < 32 && units * split < want_groups && LP / (split * 2) >= G) split *= 2;
  return split;
}

cudaError_t launch_backward(const void* value, const int64_t* shapes, const int64_t* lsi,
                            const void* loc, const void* aw, const void* grad_out,
                            void* grad_value, void* grad_loc, void* grad_aw, const Dims& d,
                            int dtype, int value_dtype, int grad_value_dtype, int sm_count,
                            int force_generic, cudaStream_t st) {
  if (dtype == MSDA_F32 && !force_generic && d.L <= kMaxSmemLevels &&
      rows_supported(d.D, value_dtype) &&
      (grad_value_dtype == MSDA_F32 || (grad_value_dtype == MSDA_BF16 && value_dtype == MSDA_BF16))) {
    PlainIO io;
    io.src.loc = static_cast<const float*>(loc);
    io.src.aw = static_cast<const float*>(aw);
    io.grad_loc = static_cast<float*>(grad_loc);
    io.grad_aw = static_cast<float*>(grad_aw);
    const float* gof = static_cast<const float*>(grad_out);
    {
      const int vec = (value_dtype == MSDA_BF16 && grad_value_dtype == MSDA_BF16) ? 8 : 4;
      if (flat_preferred(d, d.D / vec, sm_count))
        return launch_backward_flat(value, shapes, lsi, io, gof, grad_value, d, value_dtype,
                                    grad_value_dtype, sm_count, st);
    }
    // variant 3: coarsest level privatised in shared memory (the benchmark's kernel only)
    if (tuning().bwd_variant == 3 && d.D == 32 && value_dtype == MSDA_F32 && grad_value_dtype == MSDA_F32)
      return launch_backward_priv(value, shapes, lsi, io, gof, grad_value, d, sm_count, st);
#define MSDA_BWD_CASE(DD)                                                                         \
  case DD:                                                                                        \
    if (value_dtype == MSDA_F32) {                                                                \
      return launch_bwd_rows<DD, float, float, PlainIO>(                                          \
          value, shapes, lsi, io, gof, grad_value, d, choose_bwd_split(d, DD / 4, sm_count), st); \
    } else if (grad_value_dtype == MSDA_F32) {                                                    \
      return launch_bwd_rows<DD, __nv_bfloat16, float, PlainIO>(                                  \
          value, shapes, lsi, io, gof, grad_value, d, choose_bwd_split(d, DD / 4, sm_count), st); \
    } else {                                                                                      \
      return launch_bwd_rows<DD, __nv_bfloat16, __nv_bfloat16, PlainIO>(                          \
          value, shapes, lsi, io, gof, grad_value, d, choose_bwd_split(d, DD / 8, sm_count), st); \
    }
    switch (d.D) {
      MSDA_BWD_CASE(16)
      MSDA_BWD_CASE(32)
      MSDA_BWD_CASE(64)
      default: break;
    }
#undef MSDA_BWD_CASE
  }
  const int64_t n_units = static_cast<int64_t>(d.B) * d.Q * d.M;
  const unsigned blocks = static_cast<unsigned>(n_units < (1 << 20) ? n_units : (1 << 20));
#define MSDA_GEN(T, VT, GT)                                                                    \
  msda_bwd_generic_kernel<T, VT, GT><<<blocks, kGenericBwdThreads, 0, st>>>(                   \
      static_cast<const VT*>(value), shapes, lsi, static_cast<const T*>(loc),                  \
      static_cast<const T*>(aw), static_cast<const T*>(grad_out), static_cast<GT*>(grad_value), \
      static_cast<T*>(grad_loc), static_cast<T*>(grad_aw), d)
  if (dtype == MSDA_F32 && value_dtype == MSDA_F32 && grad_value_dtype == MSDA_F32) {
    MSDA_GEN(float, float, float);
  } else if (dtype == MSDA_F32 && value_dtype == MSDA_BF16 && grad_value_dtype == MSDA_F32) {
    MSDA_GEN(float, __nv_bfloat16, float);
  } else if (dtype == MSDA_F32 && value_dtype == MSDA_BF16 && grad_value_dtype == MSDA_BF16) {
    MSDA_GEN(float, __nv_bfloat16, __nv_bfloat16);
  } else if (dtype == MSDA_F64 && value_dtype == MSDA_F64 && grad_value_dtype == MSDA_F64) {
    MSDA_GEN(double, double, double);
  } else {
    return cudaErrorInvalidValue;
  }
#undef MSDA_GEN
  note_launches(1);
  note_kernel(KF_BWD_GENERIC);
  return cudaGetLastError();
}

// fused epilogue: D in {16, 32, 64}, fp32 gradients, fp32 or bf16 value
cudaError_t launch_backward_fused(const void* value, const int64_t* shapes, const int64_t* lsi,
                                  const FusedSource& src, const float* out, const float* grad_out,
                                  float* grad_value, float* grad_off, float* grad_logit,
                                  float* grad_loc, const Dims& d, int value_dtype, int sm_count,
                                  cudaStream_t st) {
  if (!rows_supported(d.D, value_dtype) || d.L > kMaxSmemLevels || !out) return cudaErrorNotSupported;
  FusedIO io;
  io.out = out;
  io.src = src;
  io.src.stats_ready = 1;
  io.grad_off = grad_off;
  io.grad_logit = grad_logit;
  io.grad_loc = grad_loc;
  io.dot = 0.f;
  const int G = d.D / 4;     // fp32 gradients: 4 channels per lane whatever the value type
  if (flat_preferred(d, G, sm_count))
    return launch_backward_flat_fused(value, shapes, lsi, io, grad_out, grad_value, d, value_dtype,
                                      sm_count, st);
  const int split = choose_bwd_split(d, G, sm_count);
#define MSDA_FUSED_CASE(DD)                                                                          \
  case DD:                                                                                           \
    return value_dtype == MSDA_F32                                                                   \
               ? launch_bwd_rows<DD, float, float, FusedIO>(value, shapes, lsi, io, grad_out,        \
                                                            grad_value, d, split, st)                \
               : launch_bwd_rows<DD, __nv_bfloat16, float, FusedIO>(value, shapes, lsi, io, grad_out, \
                                                                    grad_value, d, split, st);
  switch (d.D) {
    MSDA_FUSED_CASE(16)
    MSDA_FUSED_CASE(32)
    MSDA_FUSED_CASE(64)
    default: break;
  }
#undef MSDA_FUSED_CASE
  return cudaErrorNotSupported;
}

}  // namespace msda
