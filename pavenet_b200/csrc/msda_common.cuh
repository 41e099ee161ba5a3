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

// s / P for 0 <= s < L*P < 2^24 without an integer division per sample:
// magic = floor(2^32 / P) + 1 is exact for s * P < 2^32.
struct FastDivP {
  uint32_t magic;
  int P;
  __device__ __forceinline__ explicit FastDivP(int p) : magic(0xFFFFFFFFu / p + 1u), P(p) {}
  __device__ __forceinline__ int operator()(int s) const {
    return P == 1 ? s : static_cast<int>(__umulhi(static_cast<uint32_t>(s), magic));
  }
};

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

// 32-byte row segments: sm_100 has 256-bit global loads (ld.global.nc.v8.f32, SASS
// LDG.E.ENL2.256), so 4 lanes cover an fp32 row of 32 channels and one warp instruction moves
// 8 rows instead of 4 -- half the load, address and record instructions per row.
// RowLoad<VT, 16> is Vec16<VT>; RowLoad<float, 32> the wide form (needs 32-byte aligned rows).
template <typename VT, int BYTES>
struct RowLoad : Vec16<VT> {
  static constexpr int kBytes = 16;
};

template <>
struct RowLoad<float, 32> {
  static constexpr int VEC = 8;
  static constexpr int kBytes = 32;
  __device__ __forceinline__ static void load(const float* p, float (&v)[8]) {
    asm("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
        : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]),
          "=f"(v[7])
        : "l"(p));
  }
};

// Streaming (read-once) loads for locations / weights / grad_output: keep them
// from displacing value rows in L1.
__device__ __forceinline__ float2 ld_stream_f2(const float* p) {
  return __ldcs(reinterpret_cast<const float2*>(p));
}
__device__ __forceinline__ float ld_stream_f(const float* p) { return __ldcs(p); }

// ---- where a sample's (x, y, weight) comes from ----------------------------
// PlainSource: the reference op's inputs — materialised sampling_locations and
// attention_weights.  FusedSource: the raw projections the modules compute
// (offsets, attention logits) plus reference points; the location transform
//   loc = ref[b,q,l,(p)] + off * scale[b,q,l]      (scale == NULL: off / (W_l, H_l))
// and the softmax over the L*P logits of a row happen in the kernel, so the
// two largest intermediates of the modules are never written or re-read
// (multi_scale_deform_attn.py:373-393; transformer.py:390-412).
struct RawSample {
  float x, y, w;   // location (or offset) and weight (or logit), as loaded
};

struct PlainSource {
  const float* loc;      // never modified (constant bank); `so` indexes the bound row
  const float* aw;
  int64_t so;            // first sample of the bound row: unit * L*P
  __device__ __forceinline__ void bind(int64_t unit, int LP, int, int64_t) { so = unit * LP; }
  template <int G>
  __device__ __forceinline__ void prepass(int, int, bool) {}
  __device__ __forceinline__ RawSample load(int s) const {
    const float2 xy = ld_stream_f2(loc + 2 * (so + s));
    RawSample r;
    r.x = xy.x; r.y = xy.y; r.w = ld_stream_f(aw + so + s);
    return r;
  }
  __device__ __forceinline__ void finish(RawSample& r, int, int, const LevelInfo&) const {}
};

struct FusedSource {
  // The base pointers are never modified: as members of a kernel parameter they stay in the constant bank and
  // cost no registers.  Only the two element offsets of the bound row live in registers (the rows backward
  // needs three resident blocks per SM, i.e. <= 85 registers, and every 64-bit pointer it keeps is two of them).
  const float* off;      // (B,Q,M,L,P,2)
  const float* logit;    // (B,Q,M,L*P)
  const float* ref;      // (B,Q,L,R,2)
  const float* scale;    // (B,Q,L,2) or NULL
  float* stats;          // (B,Q,M,2): row max and 1/sum(exp) — written by the forward, read by the backward
  int R;                 // reference points per level: 1 or P
  int P;
  int stats_ready;       // backward: read stats instead of recomputing them
  float mx, inv;
  int64_t so;            // first sample of the bound row: unit * L*P
  int64_t ro;            // first (level) entry of the bound query in ref / scale: bq * L
  int64_t su;            // the bound row (only prepass reads it, so it is dead right after bind + prepass)
  __device__ __forceinline__ void bind(int64_t unit, int LP, int M, int64_t bq) {
    su = unit;
    so = unit * LP;
    ro = bq * (LP / P);
    (void)M;
  }
  // max and sum over the row's logits, by the G lanes of the group
  template <int G>
  __device__ __forceinline__ void prepass(int LP, int gl, bool writer) {
    float* st = stats + su * 2;
    if (stats_ready) {
      mx = st[0];
      inv = st[1];
      return;
    }
    const float* lg = logit + so;
    float m_ = -INFINITY;
    for (int s = gl; s < LP; s += G) m_ = fmaxf(m_, __ldg(lg + s));
#pragma unroll
    for (int o = G / 2; o >= 1; o >>= 1) m_ = fmaxf(m_, __shfl_xor_sync(0xffffffffu, m_, o));
    float sum = 0.f;
    for (int s = gl; s < LP; s += G) sum += __expf(__ldg(lg + s) - m_);
#pragma unroll
    for (int o = G / 2; o >= 1; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    mx = m_;
    inv = 1.f / sum;
    if (writer && gl == 0) {
      st[0] = mx;
      st[1] = inv;
    }
  }
  __device__ __forceinline__ RawSample load(int s) const {
    const float2 xy = ld_stream_f2(off + 2 * (so + s));
    RawSample r;
    r.x = xy.x; r.y = xy.y; r.w = __ldg(logit + so + s);
    return r;
  }
  // offsets -> location, logit -> softmax weight
  __device__ __forceinline__ void finish(RawSample& r, int s, int l, const LevelInfo& lv) const {
    const int p = s - l * P;
    const float2 rp = __ldg(reinterpret_cast<const float2*>(ref) + ((ro + l) * R + (R == 1 ? 0 : p)));
    if (scale) {
      const float2 sc = __ldg(reinterpret_cast<const float2*>(scale) + (ro + l));
      r.x = rp.x + r.x * sc.x;
      r.y = rp.y + r.y * sc.y;
    } else {
      r.x = rp.x + __fdividef(r.x, static_cast<float>(lv.W));
      r.y = rp.y + __fdividef(r.y, static_cast<float>(lv.H));
    }
    r.w = __expf(r.w - mx) * inv;
  }
};

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

// ---- forward record: weights and strides resolved by the owning lane ------
// The cell is re-anchored so that all four loads are in bounds: a corner that
// falls outside the map is replaced by a duplicate of a valid one (stride 0)
// carrying weight 0, and the attention weight is folded into the four corner
// weights.  The G lanes then run 4 unconditional 16-byte loads and 4 FMAs per
// channel - no validity tests, no zero fills, no weight arithmetic per lane.
struct __align__(16) FwdRec {
  int off;    // BYTE offset of the anchor corner (head 0, channel 0) inside the batch entry
  int rsx;    // bits 0..30 row stride in BYTES (0 if the second row is a duplicate);
              // bit 31 set if the second column is one pixel to the right (else duplicate).
              // off == kDeadOff: the sample is outside the map (weights 0, loads go to g_zero_row).
  float w1, w2;  // anchor row:   (col a, col b)
  float w3, w4;  // second row:   (col a, col b)
  int pad0, pad1;  // 32-byte records: 16-byte aligned vector access, conflict-free when adjacent
};
constexpr int kDeadOff = -1;  // FwdRec::off of a sample outside the map (byte offsets are < 2^31)

// 256 bytes of zeros: where the four loads of an out-of-map sample are pointed
// (all weights are zero too), so the gather needs no predicates and the
// compiler is free to keep the loads of several samples in flight.
__device__ __align__(16) const float g_zero_row[64] = {};

template <int ELT_BYTES>
__device__ __forceinline__ FwdRec make_fwd_rec(float x, float y, float a, const LevelInfo& lv,
                                               int MD) {
  FwdRec r;
  r.off = kDeadOff; r.rsx = 0; r.w1 = r.w2 = r.w3 = r.w4 = 0.f;
  const float h_im = y * static_cast<float>(lv.H) - 0.5f;
  const float w_im = x * static_cast<float>(lv.W) - 0.5f;
  if (h_im > -1.f && w_im > -1.f && h_im < static_cast<float>(lv.H) &&
      w_im < static_cast<float>(lv.W)) {
    const float hf = floorf(h_im), wf = floorf(w_im);
    int h0 = static_cast<int>(hf), w0 = static_cast<int>(wf);
    const float lh = h_im - hf, lw = w_im - wf;
    // rows: (ra, rb) are the weights of the anchor row and of the row below it
    float ra = 1.f - lh, rb = lh;
    int rs = lv.row_stride;
    if (h0 < 0) { h0 = 0; ra = lh; rb = 0.f; rs = 0; }            // only row 0 (the lower corner) exists
    else if (h0 + 1 > lv.H - 1) { rb = 0.f; rs = 0; }             // only row h0 exists
    float ca = 1.f - lw, cb = lw;
    int right = 1;
    if (w0 < 0) { w0 = 0; ca = lw; cb = 0.f; right = 0; }
    else if (w0 + 1 > lv.W - 1) { cb = 0.f; right = 0; }
    r.off = (lv.start + h0 * lv.W + w0) * MD * ELT_BYTES;
    r.rsx = (rs * ELT_BYTES) | (right ? static_cast<int>(0x80000000u) : 0);
    ra *= a; rb *= a;
    r.w1 = ra * ca; r.w2 = ra * cb; r.w3 = rb * ca; r.w4 = rb * cb;
  }
  return r;
}

}  // namespace msda
