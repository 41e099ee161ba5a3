/*
 * msda_ref.c — CPU ORACLE (test infrastructure, not product code).
 *
 * Plain-C restatement of the reference's multi-scale deformable attention
 * sampling, forward and backward, in float and double.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this; the product path (pavenet_b200/) never does.
 *
 * Parity: PINNED.  Checked by tests/test_oracle.py against
 *   (1) golden vectors produced in the build container by running the
 *       reference's own `multi_scale_deformable_attn_pytorch`
 *       (third_party/mmcv/mmcv/ops/multi_scale_deform_attn.py:92-149,
 *       extracted from the reference checkout by tests/golden/gen_golden.py)
 *       and its autograd gradients, on the seeded inputs of the reference's
 *       test-suite (third_party/mmcv/tests/test_ops/test_ms_deformable_attn.py:54-182)
 *       and on PAVE-Net-shaped inputs; and
 *   (2) the independent torch grid_sample port in oracle/msda_oracle.py.
 *
 * What each function follows:
 *   msda_ref_forward_*   the per-output loop of ms_deformable_im2col_gpu_kernel
 *                        (csrc/common/cuda/ms_deform_attn_cuda_kernel.cuh:200-254)
 *                        with the bilinear gather of :17-64.
 *   msda_ref_backward_*  the per-sample body of the col2im kernels (:256-345)
 *                        with the bilinear scatter / coordinate gradients of
 *                        :66-131; the block reduction over channels becomes a
 *                        plain sum.
 * Layouts: value (B,S,M,D); shapes (L,2) int64 (H,W); lsi (L,) int64;
 * loc (B,Q,M,L,P,2) normalised (x,y); aw (B,Q,M,L,P); out/grad_out (B,Q,M*D).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define DEFINE_MSDA_REF(SUFFIX, T, FLOOR)                                              \
  void msda_ref_forward_##SUFFIX(const T *value, const int64_t *shapes,                \
                                 const int64_t *lsi, const T *loc, const T *aw,        \
                                 T *out, int B, int S, int M, int D, int L, int Q,     \
                                 int P) {                                              \
    const int64_t MD = (int64_t)M * D;                                                 \
    _Pragma("omp parallel for schedule(static)")                                       \
    for (int64_t bq = 0; bq < (int64_t)B * Q; ++bq) {                                  \
      const int64_t b = bq / Q;                                                        \
      for (int m = 0; m < M; ++m) {                                                    \
        const int64_t unit = bq * M + m;                                               \
        T *o = out + unit * D;                                                         \
        for (int c = 0; c < D; ++c) o[c] = 0;                                          \
        for (int l = 0; l < L; ++l) {                                                  \
          const int H = (int)shapes[2 * l], W = (int)shapes[2 * l + 1];                \
          const T *vl = value + (b * S + lsi[l]) * MD + (int64_t)m * D;                \
          for (int p = 0; p < P; ++p) {                                                \
            const int64_t si = (unit * L + l) * P + p;                                 \
            const T x = loc[2 * si], y = loc[2 * si + 1], a = aw[si];                  \
            const T h_im = y * H - (T)0.5, w_im = x * W - (T)0.5;                      \
            if (!(h_im > -1 && w_im > -1 && h_im < H && w_im < W)) continue;           \
            const int h0 = (int)FLOOR(h_im), w0 = (int)FLOOR(w_im);                    \
            const int h1 = h0 + 1, w1 = w0 + 1;                                        \
            const T lh = h_im - h0, lw = w_im - w0, hh = 1 - lh, hw = 1 - lw;          \
            const T w1_ = hh * hw, w2_ = hh * lw, w3_ = lh * hw, w4_ = lh * lw;        \
            const T *p1 = (h0 >= 0 && w0 >= 0) ? vl + ((int64_t)h0 * W + w0) * MD : 0; \
            const T *p2 = (h0 >= 0 && w1 <= W - 1) ? vl + ((int64_t)h0 * W + w1) * MD : 0; \
            const T *p3 = (h1 <= H - 1 && w0 >= 0) ? vl + ((int64_t)h1 * W + w0) * MD : 0; \
            const T *p4 = (h1 <= H - 1 && w1 <= W - 1) ? vl + ((int64_t)h1 * W + w1) * MD : 0; \
            for (int c = 0; c < D; ++c) {                                              \
              const T v1 = p1 ? p1[c] : 0, v2 = p2 ? p2[c] : 0;                        \
              const T v3 = p3 ? p3[c] : 0, v4 = p4 ? p4[c] : 0;                        \
              o[c] += (w1_ * v1 + w2_ * v2 + w3_ * v3 + w4_ * v4) * a;                 \
            }                                                                          \
          }                                                                            \
        }                                                                              \
      }                                                                                \
    }                                                                                  \
  }                                                                                    \
                                                                                       \
  /* grad_value must be zero-filled by the caller (accumulated into);               */ \
  /* grad_loc / grad_aw are fully written.                                          */ \
  void msda_ref_backward_##SUFFIX(const T *value, const int64_t *shapes,               \
                                  const int64_t *lsi, const T *loc, const T *aw,       \
                                  const T *grad_out, T *grad_value, T *grad_loc,       \
                                  T *grad_aw, int B, int S, int M, int D, int L,       \
                                  int Q, int P) {                                      \
    const int64_t MD = (int64_t)M * D;                                                 \
    _Pragma("omp parallel for schedule(static)")                                       \
    for (int64_t b = 0; b < B; ++b) {                                                  \
      for (int64_t q = 0; q < Q; ++q) {                                                \
        for (int m = 0; m < M; ++m) {                                                  \
          const int64_t unit = (b * Q + q) * M + m;                                    \
          const T *g = grad_out + unit * D;                                            \
          for (int l = 0; l < L; ++l) {                                                \
            const int H = (int)shapes[2 * l], W = (int)shapes[2 * l + 1];              \
            const int64_t base = (b * S + lsi[l]) * MD + (int64_t)m * D;               \
            for (int p = 0; p < P; ++p) {                                              \
              const int64_t si = (unit * L + l) * P + p;                               \
              grad_loc[2 * si] = 0; grad_loc[2 * si + 1] = 0; grad_aw[si] = 0;         \
              const T x = loc[2 * si], y = loc[2 * si + 1], a = aw[si];                \
              const T h_im = y * H - (T)0.5, w_im = x * W - (T)0.5;                    \
              if (!(h_im > -1 && w_im > -1 && h_im < H && w_im < W)) continue;         \
              const int h0 = (int)FLOOR(h_im), w0 = (int)FLOOR(w_im);                  \
              const int h1 = h0 + 1, w1 = w0 + 1;                                      \
              const T lh = h_im - h0, lw = w_im - w0, hh = 1 - lh, hw = 1 - lw;        \
              const T w1_ = hh * hw, w2_ = hh * lw, w3_ = lh * hw, w4_ = lh * lw;      \
              const int ok1 = h0 >= 0 && w0 >= 0, ok2 = h0 >= 0 && w1 <= W - 1;        \
              const int ok3 = h1 <= H - 1 && w0 >= 0, ok4 = h1 <= H - 1 && w1 <= W - 1; \
              const int64_t o1 = base + ((int64_t)h0 * W + w0) * MD;                   \
              const int64_t o2 = o1 + MD, o3 = o1 + (int64_t)W * MD, o4 = o3 + MD;     \
              T s_w = 0, s_x = 0, s_y = 0;                                             \
              for (int c = 0; c < D; ++c) {                                            \
                const T top = g[c], tga = top * a;                                     \
                T v1 = 0, v2 = 0, v3 = 0, v4 = 0, gh = 0, gw = 0;                      \
                if (ok1) { v1 = value[o1 + c]; gh -= hw * v1; gw -= hh * v1; grad_value[o1 + c] += w1_ * tga; } \
                if (ok2) { v2 = value[o2 + c]; gh -= lw * v2; gw += hh * v2; grad_value[o2 + c] += w2_ * tga; } \
                if (ok3) { v3 = value[o3 + c]; gh += hw * v3; gw -= lh * v3; grad_value[o3 + c] += w3_ * tga; } \
                if (ok4) { v4 = value[o4 + c]; gh += lw * v4; gw += lh * v4; grad_value[o4 + c] += w4_ * tga; } \
                s_w += top * (w1_ * v1 + w2_ * v2 + w3_ * v3 + w4_ * v4);              \
                s_x += W * gw * tga;                                                   \
                s_y += H * gh * tga;                                                   \
              }                                                                        \
              grad_aw[si] = s_w; grad_loc[2 * si] = s_x; grad_loc[2 * si + 1] = s_y;   \
            }                                                                          \
          }                                                                            \
        }                                                                              \
      }                                                                                \
    }                                                                                  \
  }

DEFINE_MSDA_REF(f32, float, floorf)
DEFINE_MSDA_REF(f64, double, floor)

int msda_ref_abi(void) { return 1; }
