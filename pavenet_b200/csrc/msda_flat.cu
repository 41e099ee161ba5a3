// msda_flat.cu — "flat" kernels for the small-Q shapes of multi-scale deformable
// attention sampling (pose-decoder pose-aware attention, PETR pose attention):
// few (b,q,m) rows, many samples per row (300 queries x 8 heads, L*P = 68..340).
//
// Same maths as msda_fwd.cu / msda_bwd.cu (reference:
// third_party/mmcv/mmcv/ops/csrc/common/cuda/ms_deform_attn_cuda_kernel.cuh:17-131,
// 200-345).  What differs is how the work is cut.  The rows kernels give a block a
// fixed set of rows and split each row over a power-of-two number of lane groups;
// with 2 400 rows that leaves a grid of 2.7 waves of short blocks, each living for a
// chain of dependent DRAM round trips (level table -> locations -> value rows), and
// the last partial wave runs on a third of the machine.  Here the (row, 32-sample
// chunk) space is flattened and cut into one contiguous, equally long piece per warp
// of a persistent grid (two blocks per SM), so every SM finishes at the same time,
// the per-block prologue is paid once, and each lane keeps the value rows of BATCH
// samples in flight.  A row that spans several warps is combined with one
// red.global.add.v4.f32 per warp into the pre-zeroed output (forward); the backward
// needs no combination at all: every sample's gradients are written by the lane that
// owns it, and the fused softmax backward uses
//     sum_t w_t * dL/dw_t  ==  < grad_out[row], out[row] >
// (the attention-weighted sum of the sampled rows IS the forward output), so no
// row-wide reduction is left.
// The forward can also zero-fill the grad_value buffer of the coming backward from
// inside its persistent warps (the kernel is latency bound and leaves most of the
// DRAM write bandwidth idle), which takes the full-tensor clear off the critical path.
#include <type_traits>

#include "msda_kernels.h"

namespace msda {

constexpr int kFlatThreads = 256;
constexpr int kFlatWarps = kFlatThreads / 32;
constexpr int kFlatBlocksPerSM = 2;

__device__ __forceinline__ void red_add_f4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c),
               "f"(d)
               : "memory");
}

// this warp's share of the optional zero-fill, spread over its n_iter chunk iterations (plain
// streaming stores), or handed to the TMA engine up front (bulk stores from a zeroed tile of shared
// memory: no LSU slots, no registers, the fill drains at DRAM write speed underneath the gathers)
constexpr int kZeroTileBytes = 16384;
struct ClearJob {
  uint4* base;       // NULL: nothing to clear
  long long begin;   // this warp's range, in 16-byte units
  long long end;
  long long step;    // units per chunk iteration
  bool tma;
  int policy;        // store policy of the streaming form (knob clear_policy)
  __device__ __forceinline__ void bind(uint4* p, long long n16, long long w, long long W,
                                       long long n_iter, bool tma_, int policy_ = 0) {
    policy = policy_;
    base = p;
    begin = n16 * w / W;
    end = n16 * (w + 1) / W;
    step = (end - begin + n_iter - 1) / n_iter;
    tma = tma_;
  }
  __device__ __forceinline__ void run(long long it, int lane) const {
    if (!base || tma) return;
    const long long a = begin + step * it;
    const long long b = min(a + step, end);
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    if (policy == 0) {
      for (long long i = a + lane; i < b; i += 32) __stcs(base + i, z);
    } else if (policy == 1) {       // default cache policy
      for (long long i = a + lane; i < b; i += 32) base[i] = z;
    } else {                        // L2 evict-last: the zero lines are what the backward's reductions hit first
      uint64_t pol;
      asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
      for (long long i = a + lane; i < b; i += 32)
        asm volatile("st.global.L2::cache_hint.v4.b32 [%0], {%1, %1, %1, %1}, %2;" ::"l"(base + i), "r"(0), "l"(pol) : "memory");
    }
  }
  // TMA form: lane 0 of the warp queues its whole share as bulk stores of the zero tile
  __device__ __forceinline__ void issue_bulk(const uint4* zero_tile, int lane) const {
    if (!base || !tma || lane != 0) return;
    const uint32_t src = static_cast<uint32_t>(__cvta_generic_to_shared(zero_tile));
    constexpr long long kTile16 = kZeroTileBytes / 16;
    for (long long i = begin; i < end; i += kTile16) {
      const uint32_t bytes = static_cast<uint32_t>(min(kTile16, end - i) * 16);
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(base + i),
                   "r"(src), "r"(bytes)
                   : "memory");
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  }
  // before the block may exit (its shared memory is the source of the queued stores)
  __device__ __forceinline__ void drain_bulk(int lane) const {
    if (!base || !tma || lane != 0) return;
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
};

// Stream a tensor into L2 ahead of the gathers: each warp asks for its share with bulk
// prefetches (UBLKPF.L2, 4 KiB each, fire and forget).  The gathers of a small-Q launch touch
// most rows of `value` exactly once from DRAM, 128 bytes at a time in random order, which HBM
// serves at a fraction of its streaming rate; prefetched, the same bytes arrive sequentially.
__device__ __forceinline__ void l2_prefetch_share(const void* base, long long bytes, long long w,
                                                  long long W, int lane) {
  constexpr long long kPiece = 4096;
  const long long n = bytes / kPiece;
  const long long a = n * w / W, b = n * (w + 1) / W;
  const char* p = static_cast<const char*>(base);
  for (long long i = a + lane; i < b; i += 32)
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p + i * kPiece),
                 "r"(static_cast<int>(kPiece))
                 : "memory");
}

// --------------------------------------------------------------------------
// forward
// --------------------------------------------------------------------------
template <int D, typename VT, class SRC, int BATCH, int MINB, int LB = 16>
__global__ void __launch_bounds__(kFlatThreads, MINB)
msda_fwd_flat_kernel(const VT* __restrict__ value, const int64_t* __restrict__ shapes,
                     const int64_t* __restrict__ lsi, SRC src0, float* __restrict__ out, Dims d,
                     int C, long long NC, uint4* __restrict__ clear, long long clear_n16,
                     long long prefetch_bytes, int clear_tma, int phase_align) {
  using VLD = RowLoad<VT, LB>;
  constexpr int VEC = VLD::VEC;
  constexpr int G = D / VEC;     // lanes per row
  constexpr int NG = 32 / G;     // row groups per warp: one chunk = NG groups x G samples = 32 samples
  static_assert(G >= BATCH && G % BATCH == 0, "BATCH must divide the group size");

  __shared__ LevelInfo s_lvl[kMaxSmemLevels];
  __shared__ int4 s_board[kFlatWarps][G * (2 * NG + 1)];
  __shared__ __align__(128) uint4 s_zero[kZeroTileBytes / 16];   // source of the TMA zero-fill

  const int MD = d.M * D;
  for (int l = threadIdx.x; l < d.L; l += blockDim.x) s_lvl[l] = load_level(shapes, lsi, l, MD);
  if (clear && (clear_tma & 1)) {
    for (int i = threadIdx.x; i < kZeroTileBytes / 16; i += blockDim.x) s_zero[i] = make_uint4(0u, 0u, 0u, 0u);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> TMA reads
  }
  __syncthreads();

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int gl = lane & (G - 1);
  const int grp = lane / G;
  const long long W = static_cast<long long>(gridDim.x) * kFlatWarps;
  const long long w = static_cast<long long>(blockIdx.x) * kFlatWarps + warp;
  const long long c0 = NC * w / W, c1 = NC * (w + 1) / W;
  if (prefetch_bytes) l2_prefetch_share(value, prefetch_bytes, w, W, lane);

  ClearJob cj;
  cj.bind(clear, clear_n16, w, W, c1 > c0 ? c1 - c0 : 1, (clear_tma & 1) != 0, clear_tma >> 1);
  cj.issue_bulk(s_zero, lane);
  if (c1 <= c0) {            // more warps than chunks: only the zero-fill share is left
    cj.run(0, lane);
    cj.drain_bulk(lane);
    return;
  }

  const int LP = d.L * d.P;
  const uint32_t lane_b = gl * LB;
  const uint32_t MDb = MD * sizeof(VT);
  int4* board = s_board[warp];
  auto unit_of = [](int j, int grp_, int half) { return j * (2 * NG + 1) + 2 * grp_ + half; };
  const FastDivP level_of(d.P);

  // Phase alignment: a piece usually starts inside a row (at chunk k0 > 0) and ends inside the next
  // one.  Walking the part that starts at chunk 0 first makes every warp of the grid sweep the chunk
  // index -- i.e. the frames / levels, the outer sample index -- upwards at the same time, so the
  // grid's working set in L2 is about one frame of value (+ grad_value) instead of all of them.
  const long long head_end = (phase_align && c0 % C != 0) ? min(c1, (c0 / C + 1) * C) : c0;
  for (int pass = 0; pass < 2; ++pass) {
  long long c = pass == 0 ? head_end : c0;
  const long long ce = pass == 0 ? c1 : head_end;
  while (c < ce) {
    const long long row = c / C;                 // (b*Q + q)*M + m
    const int k_first = static_cast<int>(c - row * C);
    const int k_end = static_cast<int>(min(static_cast<long long>(C), k_first + (ce - c)));
    const int m = static_cast<int>(row % d.M);
    const long long bq = row / d.M;
    const long long b = bq / d.Q;
    const char* vrow = reinterpret_cast<const char*>(value + b * d.S * MD + m * D);

    SRC src = src0;
    src.bind(row, LP, d.M, bq);
    src.template prepass<32>(LP, lane, k_first == 0);   // fused: softmax max / 1/sum of the row

    float acc[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) acc[i] = 0.f;

    RawSample nxt;
    nxt.x = nxt.y = nxt.w = 0.f;
    {
      const int s = 32 * k_first + lane;
      if (s < LP) nxt = src.load(s);
    }
    for (int k = k_first; k < k_end; ++k) {
      const int s = 32 * k + lane;
      {
        FwdRec r;
        r.off = kDeadOff; r.rsx = 0; r.w1 = r.w2 = r.w3 = r.w4 = 0.f;
        if (s < LP) {
          const int l = level_of(s);
          const LevelInfo lv = s_lvl[l];
          RawSample cur = nxt;
          src.finish(cur, s, l, lv);
          r = make_fwd_rec<sizeof(VT)>(cur.x, cur.y, cur.w, lv, MD);
        }
        board[unit_of(gl, grp, 0)] =
            make_int4(r.off, r.rsx, __float_as_int(r.w1), __float_as_int(r.w2));
        *reinterpret_cast<float2*>(&board[unit_of(gl, grp, 1)]) = make_float2(r.w3, r.w4);
        const int sn = s + 32;
        if (k + 1 < k_end && sn < LP) nxt = src.load(sn);
      }
      __syncwarp();
      cj.run(c - c0 + (k - k_first), lane);
      const int nvalid = LP - (32 * k + grp * G);   // samples of this group in this chunk (may be <= 0)
#pragma unroll
      for (int j0 = 0; j0 < G; j0 += BATCH) {
        if (j0 < nvalid) {
          int4 q[BATCH];
          float2 w34[BATCH];
#pragma unroll
          for (int t = 0; t < BATCH; ++t) {
            q[t] = board[unit_of(j0 + t, grp, 0)];
            w34[t] = *reinterpret_cast<const float2*>(&board[unit_of(j0 + t, grp, 1)]);
          }
          float v[BATCH][4][VEC];
#pragma unroll
          for (int t = 0; t < BATCH; ++t) {
            const bool alive = q[t].x != kDeadOff;
            const uint32_t rs = q[t].y & 0x7fffffff;
            const uint32_t xs = (q[t].y >> 31) & MDb;
            const char* sp = alive ? vrow : reinterpret_cast<const char*>(g_zero_row);
            const uint32_t o1 = (alive ? static_cast<uint32_t>(q[t].x) : 0u) + lane_b;
            VLD::load(reinterpret_cast<const VT*>(sp + o1), v[t][0]);
            VLD::load(reinterpret_cast<const VT*>(sp + (o1 + xs)), v[t][1]);
            VLD::load(reinterpret_cast<const VT*>(sp + (o1 + rs)), v[t][2]);
            VLD::load(reinterpret_cast<const VT*>(sp + (o1 + rs + xs)), v[t][3]);
          }
#pragma unroll
          for (int t = 0; t < BATCH; ++t) {
            const float w1 = __int_as_float(q[t].z), w2 = __int_as_float(q[t].w);
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
              acc[i] = fmaf(w1, v[t][0][i], acc[i]);
              acc[i] = fmaf(w2, v[t][1][i], acc[i]);
              acc[i] = fmaf(w34[t].x, v[t][2][i], acc[i]);
              acc[i] = fmaf(w34[t].y, v[t][3][i], acc[i]);
            }
          }
        }
      }
      __syncwarp();
    }
    // the NG groups hold partial rows over disjoint samples
#pragma unroll
    for (int off = G; off < 32; off <<= 1) {
#pragma unroll
      for (int i = 0; i < VEC; ++i) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], off);
    }
    if (grp == 0) {
      float* o = out + row * D + gl * VEC;
      if (k_first == 0 && k_end == C) {          // the whole row is this warp's
#pragma unroll
        for (int i = 0; i < VEC; i += 4)
          *reinterpret_cast<float4*>(o + i) = make_float4(acc[i], acc[i + 1], acc[i + 2], acc[i + 3]);
      } else {                                   // shared with the neighbouring warps: out was zeroed
#pragma unroll
        for (int i = 0; i < VEC; i += 4) red_add_f4(o + i, acc[i], acc[i + 1], acc[i + 2], acc[i + 3]);
      }
    }
    c += k_end - k_first;
  }
  }
  cj.drain_bulk(lane);
}

// --------------------------------------------------------------------------
// backward
// --------------------------------------------------------------------------
// PIPE: the value rows of the next batch of samples are requested before the reductions of the
// current batch are issued (their registers are free once the per-sample sums are formed), so
// loads are in flight while the reductions drain instead of after them.
template <int D, typename VT, typename GT, class IO, int BATCH, int MINB, bool PIPE = false>
__global__ void __launch_bounds__(kFlatThreads, MINB)
msda_bwd_flat_kernel(const VT* __restrict__ value, const int64_t* __restrict__ shapes,
                     const int64_t* __restrict__ lsi, IO io0, const float* __restrict__ grad_out,
                     GT* __restrict__ grad_value, Dims d, int C, long long NC,
                     long long prefetch_bytes, int phase_align) {
  using VL = BwdVec<VT, GT>;
  constexpr int VEC = VL::VEC;
  constexpr int G = D / VEC;
  constexpr int NG = 32 / G;
  static_assert(G >= BATCH && G % BATCH == 0, "BATCH must divide the group size");

  __shared__ LevelInfo s_lvl[kMaxSmemLevels];
  __shared__ int4 s_board[kFlatWarps][G * (2 * NG + 1)];

  const int MD = d.M * D;
  for (int l = threadIdx.x; l < d.L; l += blockDim.x) s_lvl[l] = load_level(shapes, lsi, l, MD);
  __syncthreads();

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int gl = lane & (G - 1);
  const int grp = lane / G;
  const long long W = static_cast<long long>(gridDim.x) * kFlatWarps;
  const long long w = static_cast<long long>(blockIdx.x) * kFlatWarps + warp;
  const long long c0 = NC * w / W, c1 = NC * (w + 1) / W;
  if (prefetch_bytes) l2_prefetch_share(value, prefetch_bytes, w, W, lane);
  if (c1 <= c0) return;

  const int LP = d.L * d.P;
  int4* board = s_board[warp];
  auto unit_of = [](int j, int grp_, int half) { return j * (2 * NG + 1) + 2 * grp_ + half; };
  const FastDivP level_of(d.P);

  // Phase alignment: a piece usually starts inside a row (at chunk k0 > 0) and ends inside the next
  // one.  Walking the part that starts at chunk 0 first makes every warp of the grid sweep the chunk
  // index -- i.e. the frames / levels, the outer sample index -- upwards at the same time, so the
  // grid's working set in L2 is about one frame of value (+ grad_value) instead of all of them.
  const long long head_end = (phase_align && c0 % C != 0) ? min(c1, (c0 / C + 1) * C) : c0;
  for (int pass = 0; pass < 2; ++pass) {
  long long c = pass == 0 ? head_end : c0;
  const long long ce = pass == 0 ? c1 : head_end;
  while (c < ce) {
    const long long row = c / C;
    const int k_first = static_cast<int>(c - row * C);
    const int k_end = static_cast<int>(min(static_cast<long long>(C), k_first + (ce - c)));
    const int m = static_cast<int>(row % d.M);
    const long long bq = row / d.M;
    const long long b = bq / d.Q;
    const long long boff = b * d.S * MD + m * D + gl * VEC;
    const VT* vbase = value + boff;
    GT* gvbase = grad_value + boff;

    IO io = io0;
    io.bind(row, LP, d.M, bq);
    io.src.template prepass<32>(LP, lane, false);   // fused: stats saved by the forward

    float g[VEC];
    {
      const float* gp = grad_out + row * D + gl * VEC;
#pragma unroll
      for (int i = 0; i < VEC; i += 4) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(gp + i));
        g[i] = t.x; g[i + 1] = t.y; g[i + 2] = t.z; g[i + 3] = t.w;
      }
    }
    io.template row_dot<G, VEC>(g, row, D, gl);    // fused: <grad_out[row], out[row]>

    RawSample nxt;
    nxt.x = nxt.y = nxt.w = 0.f;
    {
      const int s = 32 * k_first + lane;
      if (s < LP) nxt = io.src.load(s);
    }
    for (int k = k_first; k < k_end; ++k) {
      const int s = 32 * k + lane;
      float Wf = 0.f, Hf = 0.f, w_true = 0.f;
      int lvl = 0;
      {
        SampleRec r;
        r.off00 = 0; r.meta = 0; r.lh = 0.f; r.lw = 0.f; r.a = 0.f; r.rs = 0;
        if (s < LP) {
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
        const int sn = s + 32;
        if (k + 1 < k_end && sn < LP) nxt = io.src.load(sn);
      }
      __syncwarp();

      float pw[G], px[G], py[G];
#pragma unroll
      for (int j = 0; j < G; ++j) pw[j] = px[j] = py[j] = 0.f;
      const int nvalid = LP - (32 * k + grp * G);
      // one batch = BATCH samples of this lane group: records, value rows, per-sample sums, reductions
      auto read_recs = [&](int j0, int4 (&q)[BATCH], int2 (&ar)[BATCH]) {
#pragma unroll
        for (int t = 0; t < BATCH; ++t) {
          q[t] = board[unit_of(j0 + t, grp, 0)];
          ar[t] = *reinterpret_cast<const int2*>(&board[unit_of(j0 + t, grp, 1)]);
        }
      };
      auto load_rows = [&](const int4 (&q)[BATCH], const int2 (&ar)[BATCH], float (&v)[BATCH][4][VEC]) {
#pragma unroll
        for (int t = 0; t < BATCH; ++t) {
          const int meta = q[t].y;
          const int rs = ar[t].y;
          const VT* p = vbase + q[t].x;
#pragma unroll
          for (int i = 0; i < VEC; ++i) v[t][0][i] = v[t][1][i] = v[t][2][i] = v[t][3][i] = 0.f;
          if (meta & 1) VL::load(p, v[t][0]);
          if (meta & 2) VL::load(p + MD, v[t][1]);
          if (meta & 4) VL::load(p + rs, v[t][2]);
          if (meta & 8) VL::load(p + rs + MD, v[t][3]);
        }
      };
      auto sums = [&](int j0, const int4 (&q)[BATCH], const int2 (&ar)[BATCH],
                      const float (&v)[BATCH][4][VEC]) {
#pragma unroll
        for (int t = 0; t < BATCH; ++t) {
          const float a = __int_as_float(ar[t].x);
          const float lh = __int_as_float(q[t].z), lw = __int_as_float(q[t].w);
          const float hh = 1.f - lh, hw = 1.f - lw;
          const float w1 = hh * hw, w2 = hh * lw, w3 = lh * hw, w4 = lh * lw;
          float sw = 0.f, sx = 0.f, sy = 0.f;
#pragma unroll
          for (int i = 0; i < VEC; ++i) {
            const float tg = g[i] * a;
            const float val = w1 * v[t][0][i] + w2 * v[t][1][i] + w3 * v[t][2][i] + w4 * v[t][3][i];
            const float gw = hh * (v[t][1][i] - v[t][0][i]) + lh * (v[t][3][i] - v[t][2][i]);
            const float gh = hw * (v[t][2][i] - v[t][0][i]) + lw * (v[t][3][i] - v[t][1][i]);
            sw += g[i] * val;
            sx += gw * tg;
            sy += gh * tg;
          }
          pw[j0 + t] = sw; px[j0 + t] = sx; py[j0 + t] = sy;
        }
      };
      auto scatter = [&](const int4 (&q)[BATCH], const int2 (&ar)[BATCH]) {
#pragma unroll
        for (int t = 0; t < BATCH; ++t) {
          const int meta = q[t].y;
          const int rs = ar[t].y;
          const float a = __int_as_float(ar[t].x);
          const float lh = __int_as_float(q[t].z), lw = __int_as_float(q[t].w);
          const float hh = 1.f - lh, hw = 1.f - lw;
          GT* gp = gvbase + q[t].x;
          float tt[VEC];
#define MSDA_SCATTER(BIT, WK, PTR)                                              \
  if (meta & BIT) {                                                             \
    _Pragma("unroll") for (int i = 0; i < VEC; ++i) tt[i] = (WK) * a * g[i];   \
    red_add_row(PTR, tt);                                                       \
  }
          MSDA_SCATTER(1, hh * hw, gp)
          MSDA_SCATTER(2, hh * lw, gp + MD)
          MSDA_SCATTER(4, lh * hw, gp + rs)
          MSDA_SCATTER(8, lh * lw, gp + rs + MD)
#undef MSDA_SCATTER
        }
      };
      if constexpr (!PIPE) {
#pragma unroll
        for (int j0 = 0; j0 < G; j0 += BATCH) {
          if (j0 < nvalid) {
            int4 q[BATCH];
            int2 ar[BATCH];
            float v[BATCH][4][VEC];
            read_recs(j0, q, ar);
            load_rows(q, ar, v);
            sums(j0, q, ar, v);
            scatter(q, ar);
          }
        }
      } else {
        int4 q[BATCH], qn[BATCH];
        int2 ar[BATCH], arn[BATCH];
        float v[BATCH][4][VEC];
        if (nvalid > 0) {
          read_recs(0, q, ar);
          load_rows(q, ar, v);
        }
#pragma unroll
        for (int j0 = 0; j0 < G; j0 += BATCH) {
          if (j0 < nvalid) {
            sums(j0, q, ar, v);                       // v is free after this
            const bool more = j0 + BATCH < G && j0 + BATCH < nvalid;
            if (more) {
              read_recs(j0 + BATCH < G ? j0 + BATCH : 0, qn, arn);
              load_rows(qn, arn, v);                  // in flight while the reductions below drain
            }
            scatter(q, ar);
            if (more) {
#pragma unroll
              for (int t = 0; t < BATCH; ++t) { q[t] = qn[t]; ar[t] = arn[t]; }
            }
          }
        }
      }
      const float tw = group_transpose_reduce<G>(pw, gl);
      const float tx = group_transpose_reduce<G>(px, gl);
      const float ty = group_transpose_reduce<G>(py, gl);
      if (s < LP) io.store(s, lvl, tw, tx, ty, w_true, Wf, Hf);
      __syncwarp();
    }
    c += k_end - k_first;
  }
  }
}

// --------------------------------------------------------------------------
// launchers
// --------------------------------------------------------------------------
// Rows are few enough that the rows kernels would have to split them, and long
// enough that 32-sample chunks are mostly full.
bool flat_preferred(const Dims& d, int G, int sm_count) {
  const int mode = tuning().flat;   // 0 never, 1 heuristic (default), 2 whenever legal
  if (mode == 0) return false;
  const int LP = d.L * d.P;
  const long long units = static_cast<long long>(d.B) * d.Q * d.M;
  const long long C = (LP + 31) / 32;
  if (units * C >= (1LL << 40)) return false;
  if (mode == 2) return true;
  if (LP < 48) return false;
  const long long want_groups = static_cast<long long>(sm_count) * 48 * (32 / G);
  return units * 2 <= want_groups;   // the rows kernel would split every row at least in two
}

// how much of `value` to stream into L2 ahead of the gathers: all of it when it fits
// comfortably (knob l2_prefetch: bit 0 forward, bit 1 backward; l2_prefetch_mb the size limit)
template <typename VT>
static long long prefetch_bytes_for(const Dims& d, int bit) {
  if (!(tuning().l2_prefetch & bit)) return 0;
  const long long bytes = static_cast<long long>(d.B) * d.S * d.M * d.D * sizeof(VT);
  return bytes <= static_cast<long long>(tuning().l2_prefetch_mb) * (1 << 20) ? bytes : 0;
}

template <int D, typename VT, class SRC>
static cudaError_t launch_fwd_flat(const void* value, const int64_t* shapes, const int64_t* lsi,
                                   const SRC& src, float* out, const Dims& d, int sm_count,
                                   void* clear, size_t clear_bytes, cudaStream_t st) {
  const int LP = d.L * d.P;
  const int C = (LP + 31) / 32;
  const long long R = static_cast<long long>(d.B) * d.Q * d.M;
  const long long NC = R * C;
  cudaError_t e = cudaMemsetAsync(out, 0, static_cast<size_t>(R) * D * sizeof(float), st);
  if (e != cudaSuccess) return e;
  constexpr int G = D / Vec16<VT>::VEC;
  constexpr int BATCH = G >= 4 ? 4 : G;
  const long long clear_n16 = static_cast<long long>(clear_bytes / 16);
  const long long pf = prefetch_bytes_for<VT>(d, 1);
  const int clear_tma = (tuning().clear_mode == 2 ? 1 : 0) | (tuning().clear_policy << 1);   // bit 0: TMA bulk stores instead of STG; bits 1..: store policy experiment
  const int order = tuning().flat_order;
#define MSDA_FWD_FLAT_LAUNCH(BATCH_, MINB_)                                                       \
  msda_fwd_flat_kernel<D, VT, SRC, BATCH_, MINB_>                                                 \
      <<<static_cast<unsigned>(sm_count * MINB_), kFlatThreads, 0, st>>>(                         \
          static_cast<const VT*>(value), shapes, lsi, src, out, d, C, NC,                         \
          static_cast<uint4*>(clear), clear_n16, pf, clear_tma, order)
  // tuning sweep (knob flat_fwd_cfg), instantiated for the benchmark's kernel only
  if constexpr (D == 32 && std::is_same<VT, float>::value && std::is_same<SRC, PlainSource>::value) {
    switch (tuning().flat_fwd_cfg) {
      case 1: MSDA_FWD_FLAT_LAUNCH(2, 4); break;
      case 2: MSDA_FWD_FLAT_LAUNCH(2, 3); break;
      case 3: MSDA_FWD_FLAT_LAUNCH((G >= 8 ? 8 : G), 1); break;
      case 4: MSDA_FWD_FLAT_LAUNCH(1, 4); break;
      case 5: MSDA_FWD_FLAT_LAUNCH(4, 3); break;
      case 6:     // 256-bit value loads: 4 lanes per row, 8 rows per warp instruction
        if ((reinterpret_cast<uintptr_t>(value) & 31u) == 0) {
          msda_fwd_flat_kernel<D, VT, SRC, 2, 2, 32>
              <<<static_cast<unsigned>(sm_count * 2), kFlatThreads, 0, st>>>(
                  static_cast<const VT*>(value), shapes, lsi, src, out, d, C, NC,
                  static_cast<uint4*>(clear), clear_n16, pf, clear_tma, order);
          break;
        }
        MSDA_FWD_FLAT_LAUNCH(BATCH, kFlatBlocksPerSM);
        break;
      default: MSDA_FWD_FLAT_LAUNCH(BATCH, kFlatBlocksPerSM); break;
    }
  } else {
    MSDA_FWD_FLAT_LAUNCH(BATCH, kFlatBlocksPerSM);
  }
#undef MSDA_FWD_FLAT_LAUNCH
  note_launches(1);
  note_kernel(std::is_same<SRC, FusedSource>::value ? KF_FWD_FLAT_FUSED : KF_FWD_FLAT);
  return cudaGetLastError();
}

template <int D, typename VT, typename GT, class IO>
static cudaError_t launch_bwd_flat(const void* value, const int64_t* shapes, const int64_t* lsi,
                                   const IO& io, const float* go, void* gv, const Dims& d,
                                   int sm_count, cudaStream_t st) {
  const int LP = d.L * d.P;
  const int C = (LP + 31) / 32;
  const long long NC = static_cast<long long>(d.B) * d.Q * d.M * C;
  constexpr int G = D / BwdVec<VT, GT>::VEC;
  constexpr int BATCH = G >= 2 ? 2 : 1;
  const long long pf = prefetch_bytes_for<VT>(d, 2);
  const int order = tuning().flat_order;
#define MSDA_BWD_FLAT_LAUNCH(BATCH_, MINB_)                                                       \
  msda_bwd_flat_kernel<D, VT, GT, IO, BATCH_, MINB_>                                              \
      <<<static_cast<unsigned>(sm_count * MINB_), kFlatThreads, 0, st>>>(                         \
          static_cast<const VT*>(value), shapes, lsi, io, go, static_cast<GT*>(gv), d, C, NC, pf, order)
  if constexpr (D == 32 && std::is_same<VT, float>::value && std::is_same<IO, PlainIO>::value) {
    switch (tuning().flat_bwd_cfg) {
      case 1: MSDA_BWD_FLAT_LAUNCH(1, 3); break;
      case 2: MSDA_BWD_FLAT_LAUNCH(2, 3); break;
      case 3: MSDA_BWD_FLAT_LAUNCH(4, 1); break;
      case 4: MSDA_BWD_FLAT_LAUNCH(1, 4); break;
      case 5: MSDA_BWD_FLAT_LAUNCH(4, 2); break;
      case 6:     // software-pipelined: next batch's loads ahead of this batch's reductions
        msda_bwd_flat_kernel<D, VT, GT, IO, 2, 2, true>
            <<<static_cast<unsigned>(sm_count * 2), kFlatThreads, 0, st>>>(
                static_cast<const VT*>(value), shapes, lsi, io, go, static_cast<GT*>(gv), d, C, NC, pf, order);
        break;
      case 7:
        msda_bwd_flat_kernel<D, VT, GT, IO, 1, 3, true>
            <<<static_cast<unsigned>(sm_count * 3), kFlatThreads, 0, st>>>(
                static_cast<const VT*>(value), shapes, lsi, io, go, static_cast<GT*>(gv), d, C, NC, pf, order);
        break;
      default: MSDA_BWD_FLAT_LAUNCH(BATCH, kFlatBlocksPerSM); break;
    }
  } else {
    MSDA_BWD_FLAT_LAUNCH(BATCH, kFlatBlocksPerSM);
  }
#undef MSDA_BWD_FLAT_LAUNCH
  note_launches(1);
  note_kernel(IO::kFused ? KF_BWD_FLAT_FUSED : KF_BWD_FLAT);
  return cudaGetLastError();
}

cudaError_t launch_forward_flat(const void* value, const int64_t* shapes, const int64_t* lsi,
                                const PlainSource& src, float* out, const Dims& d, int value_dtype,
                                int sm_count, void* clear, size_t clear_bytes, cudaStream_t st) {
#define MSDA_FLAT_CASE(DD)                                                                      \
  case DD:                                                                                      \
    return value_dtype == MSDA_F32                                                              \
               ? launch_fwd_flat<DD, float, PlainSource>(value, shapes, lsi, src, out, d,       \
                                                         sm_count, clear, clear_bytes, st)      \
               : launch_fwd_flat<DD, __nv_bfloat16, PlainSource>(value, shapes, lsi, src, out,  \
                                                                 d, sm_count, clear,            \
                                                                 clear_bytes, st);
  switch (d.D) {
    MSDA_FLAT_CASE(16)
    MSDA_FLAT_CASE(32)
    MSDA_FLAT_CASE(64)
    default: break;
  }
#undef MSDA_FLAT_CASE
  return cudaErrorNotSupported;
}

cudaError_t launch_forward_flat_fused(const void* value, const int64_t* shapes, const int64_t* lsi,
                                      const FusedSource& src, float* out, const Dims& d,
                                      int value_dtype, int sm_count, void* clear,
                                      size_t clear_bytes, cudaStream_t st) {
#define MSDA_FLAT_CASE(DD)                                                                        \
  case DD:                                                                                        \
    return value_dtype == MSDA_F32                                                                \
               ? launch_fwd_flat<DD, float, FusedSource>(value, shapes, lsi, src, out, d, sm_count, \
                                                         clear, clear_bytes, st)                  \
               : launch_fwd_flat<DD, __nv_bfloat16, FusedSource>(value, shapes, lsi, src, out, d, \
                                                                 sm_count, clear, clear_bytes, st);
  switch (d.D) {
    MSDA_FLAT_CASE(16)
    MSDA_FLAT_CASE(32)
    MSDA_FLAT_CASE(64)
    default: break;
  }
#undef MSDA_FLAT_CASE
  return cudaErrorNotSupported;
}

cudaError_t launch_backward_flat(const void* value, const int64_t* shapes, const int64_t* lsi,
                                 const PlainIO& io, const float* grad_out, void* grad_value,
                                 const Dims& d, int value_dtype, int grad_value_dtype, int sm_count,
                                 cudaStream_t st) {
#define MSDA_FLAT_CASE(DD)                                                                        \
  case DD:                                                                                        \
    if (value_dtype == MSDA_F32)                                                                  \
      return launch_bwd_flat<DD, float, float, PlainIO>(value, shapes, lsi, io, grad_out,         \
                                                        grad_value, d, sm_count, st);             \
    if (grad_value_dtype == MSDA_F32)                                                             \
      return launch_bwd_flat<DD, __nv_bfloat16, float, PlainIO>(value, shapes, lsi, io, grad_out, \
                                                                grad_value, d, sm_count, st);     \
    return launch_bwd_flat<DD, __nv_bfloat16, __nv_bfloat16, PlainIO>(                            \
        value, shapes, lsi, io, grad_out, grad_value, d, sm_count, st);
  switch (d.D) {
    MSDA_FLAT_CASE(16)
    MSDA_FLAT_CASE(32)
    MSDA_FLAT_CASE(64)
    default: break;
  }
#undef MSDA_FLAT_CASE
  return cudaErrorNotSupported;
}

cudaError_t launch_backward_flat_fused(const void* value, const int64_t* shapes,
                                       const int64_t* lsi, const FusedIO& io,
                                       const float* grad_out, float* grad_value, const Dims& d,
                                       int value_dtype, int sm_count, cudaStream_t st) {
  if (!io.out) return cudaErrorNotSupported;
#define MSDA_FLAT_CASE(DD)                                                                         \
  case DD:                                                                                         \
    return value_dtype == MSDA_F32                                                                 \
               ? launch_bwd_flat<DD, float, float, FusedIO>(value, shapes, lsi, io, grad_out,      \
                                                            grad_value, d, sm_count, st)           \
               : launch_bwd_flat<DD, __nv_bfloat16, float, FusedIO>(value, shapes, lsi, io,        \
                                                                    grad_out, grad_value, d,       \
                                                                    sm_count, st);
  switch (d.D) {
    MSDA_FLAT_CASE(16)
    MSDA_FLAT_CASE(32)
    MSDA_FLAT_CASE(64)
    default: break;
  }
#undef MSDA_FLAT_CASE
  return cudaErrorNotSupported;
}

}  // namespace msda
