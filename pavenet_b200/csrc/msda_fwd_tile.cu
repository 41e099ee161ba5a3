// msda_fwd_tile.cu — forward kernel that STAGES LEVEL TILES IN SHARED MEMORY, for the encoder:
// queries are the pixels of the feature pyramid itself (Q == S), 8 heads x 32 channels, P = 4.
//
// Same maths as msda_fwd.cu (reference: ms_deform_attn_cuda_kernel.cuh:17-64, 200-254).  What
// changes is where the value rows come from.  The rows kernel gathers every bilinear corner
// through L1; with 32 raster-consecutive queries per block 31 % of those gathers miss and pay the
// 64 B/clk L2->SM fill path.  Here a block owns a 16 x 16 PATCH of query pixels of one pyramid
// level and one head, and walks the value levels one after the other: for each level it stages
// the window of value rows the patch can reach — the patch's footprint in that level plus a halo
// of kHalo pixels, at most 27 x 27 rows of 128 bytes — into shared memory with cp.async
// (LDGSTS, L1 bypassed), double buffered so that the window of the next level (or of the next
// patch) streams in while the current one is being sampled, and then takes the four corners of
// every sample from the window with conflict-free 16-byte shared-memory loads (8 lanes = one
// 128-byte row = all 32 banks: one wavefront per row, no tag lookup, no miss).
// Learned offsets are not bounded, so a sample whose 2x2 footprint is not completely inside the
// window is fetched from global memory exactly as in the rows kernel; a warp step takes the
// shared-memory path when all four of its lane groups can (out-of-map samples point at a zero
// row kept behind each window), and otherwise a generic-load path that serves both spaces with
// one instruction per corner.  A (patch, level) pair whose window does not fit the buffer (a
// level-3 patch covers most of level 0) is sampled from global memory altogether.
//
// Block = 1024 threads = 128 lane groups of 8; a group owns 2 of the patch's 256 queries and
// keeps their partial output rows in registers across the levels.  Per level a group has
// 2 queries x P = 8 samples = one chunk: lane i of the group resolves the geometry of sample i
// and publishes it on the warp's record board, as in the rows kernel.  Persistent grid, one
// block per SM; (batch entry, patch, head) items are strided over the grid and flattened with
// their levels into one sequence of steps for the staging pipeline; one thread works out the
// next step's window while the block samples the current one.
#include <cuda_pipeline_primitives.h>

#include "msda_kernels.h"

namespace msda {

constexpr int kTileThreads = 1024;
constexpr int kTileWarps = kTileThreads / 32;
constexpr int kTileG = 8;                          // lanes per fp32 row of 32 channels
constexpr int kTileGroups = kTileThreads / kTileG; // 128
constexpr int kPatch = 16;                         // patch side, in query pixels
constexpr int kSlots = kPatch * kPatch / kTileGroups;   // queries per lane group: 2
constexpr int kMaxTileLevels = 8;
constexpr int kHalo = 5;
constexpr int kTileRows = (kPatch + 1 + 2 * kHalo) * (kPatch + 1 + 2 * kHalo);   // 27 x 27 = 729
constexpr int kBufRows = kTileRows + 1;            // + the zero row out-of-map samples read
constexpr int kRowBytes = 128;
constexpr int kBoardUnits = kTileG * (2 * (32 / kTileG) + 1);

// One step of the flattened (item, level) sequence.
struct TileStep {
  int valid;
  int b, m, lq, px0, py0;  // batch entry, head, query level and patch origin (pixels of that level)
  int l;                   // value level sampled in this step
  int x0, y0, ww, wh;      // window of level l staged in shared memory (pixels of level l)
  int tiled;
};

constexpr int kStepTable = 64;          // steps whose descriptors are worked out at once

struct TileSmem {
  LevelInfo lvl[kMaxTileLevels];
  int patch_base[kMaxTileLevels + 1];   // first patch index of each query level, per (b, m)
  TileStep step[kStepTable + 1];        // descriptors of this block's next steps (+ the one after them)
  int4 board[kTileWarps][kBoardUnits];
  // STAGE: followed by float tile[2][kBufRows * 32]
};

__device__ __forceinline__ void make_tile_step(TileStep& t, const TileSmem& sm, long long gstep,
                                               int L, int M, long long n_items) {
  const long long item = gstep / L;
  t.l = static_cast<int>(gstep - item * L);
  t.valid = item < n_items;
  t.b = t.m = t.lq = t.px0 = t.py0 = t.x0 = t.y0 = t.ww = t.wh = t.tiled = 0;
  if (!t.valid) return;
  t.m = static_cast<int>(item % M);
  const long long rest = item / M;
  const int per_b = sm.patch_base[L];
  t.b = static_cast<int>(rest / per_b);
  const int patch = static_cast<int>(rest - static_cast<long long>(t.b) * per_b);
  int lq = 0;
  while (lq + 1 < L && patch >= sm.patch_base[lq + 1]) ++lq;
  t.lq = lq;
  const int npx = (sm.lvl[lq].W + kPatch - 1) / kPatch;
  const int pi = patch - sm.patch_base[lq];
  t.px0 = (pi % npx) * kPatch;
  t.py0 = (pi / npx) * kPatch;
  // window of level l: where samples of this patch land when |offset| <= kHalo pixels of level l
  const float sx = static_cast<float>(sm.lvl[t.l].W) / static_cast<float>(sm.lvl[lq].W);
  const float sy = static_cast<float>(sm.lvl[t.l].H) / static_cast<float>(sm.lvl[lq].H);
  const int px1 = min(t.px0 + kPatch, sm.lvl[lq].W), py1 = min(t.py0 + kPatch, sm.lvl[lq].H);
  int xa = static_cast<int>(floorf((t.px0 + 0.5f) * sx - 0.5f)) - kHalo;
  int xb = static_cast<int>(floorf((px1 - 0.5f) * sx - 0.5f)) + 1 + kHalo;
  int ya = static_cast<int>(floorf((t.py0 + 0.5f) * sy - 0.5f)) - kHalo;
  int yb = static_cast<int>(floorf((py1 - 0.5f) * sy - 0.5f)) + 1 + kHalo;
  xa = max(xa, 0); ya = max(ya, 0);
  xb = min(xb, sm.lvl[t.l].W - 1); yb = min(yb, sm.lvl[t.l].H - 1);
  t.x0 = xa; t.y0 = ya;
  t.ww = xb - xa + 1; t.wh = yb - ya + 1;
  t.tiled = t.ww > 0 && t.wh > 0 && t.ww * t.wh <= kTileRows;
}

// The record a lane publishes for one sample.  Offsets address either the window buffer
// (bit 30 of rsx set; dead samples: the zero row behind the window) or the batch entry of value.
struct TileRec {
  int off;     // byte offset of the anchor corner
  int rsx;     // bits 0..29 row stride in bytes, bit 30 "window", bit 31 "second column exists"
  float w1, w2, w3, w4;
};

__device__ __forceinline__ TileRec make_tile_rec(float x, float y, float a, const LevelInfo& lv,
                                                 int MDb, const TileStep& t) {
  TileRec r;
  r.off = kTileRows * kRowBytes;           // the zero row: strides 0, weights 0
  r.rsx = 0x40000000;
  r.w1 = r.w2 = r.w3 = r.w4 = 0.f;
  const float h_im = y * static_cast<float>(lv.H) - 0.5f;
  const float w_im = x * static_cast<float>(lv.W) - 0.5f;
  if (h_im > -1.f && w_im > -1.f && h_im < static_cast<float>(lv.H) &&
      w_im < static_cast<float>(lv.W)) {
    const float hf = floorf(h_im), wf = floorf(w_im);
    int h0 = static_cast<int>(hf), w0 = static_cast<int>(wf);
    const float lh = h_im - hf, lw = w_im - wf;
    float ra = 1.f - lh, rb = lh;
    int dy = 1, dx = 1;
    if (h0 < 0) { h0 = 0; ra = lh; rb = 0.f; dy = 0; }
    else if (h0 + 1 > lv.H - 1) { rb = 0.f; dy = 0; }
    float ca = 1.f - lw, cb = lw;
    if (w0 < 0) { w0 = 0; ca = lw; cb = 0.f; dx = 0; }
    else if (w0 + 1 > lv.W - 1) { cb = 0.f; dx = 0; }
    ra *= a; rb *= a;
    r.w1 = ra * ca; r.w2 = ra * cb; r.w3 = rb * ca; r.w4 = rb * cb;
    const int right = dx ? static_cast<int>(0x80000000u) : 0;
    if (t.tiled && w0 >= t.x0 && w0 + dx < t.x0 + t.ww && h0 >= t.y0 && h0 + dy < t.y0 + t.wh) {
      r.off = ((h0 - t.y0) * t.ww + (w0 - t.x0)) * kRowBytes;
      r.rsx = (dy ? t.ww * kRowBytes : 0) | 0x40000000 | right;
    } else {
      r.off = (lv.start + h0 * lv.W + w0) * MDb;
      r.rsx = (dy ? lv.row_stride * static_cast<int>(sizeof(float)) : 0) | right;
    }
  }
  return r;
}

// STAGE = true: windows staged in shared memory (fwd_variant 5).  STAGE = false (fwd_variant 6):
// the same patch-per-block, level-by-level walk, but every corner is an ordinary cached global
// load — the window of one level (<= 93 KB) is what the block touches at any time, so L1 itself
// holds it; no staging traffic, no block barriers.
template <int P, bool STAGE>
__global__ void __launch_bounds__(kTileThreads, 1)
msda_fwd_tile_kernel(const float* __restrict__ value, const int64_t* __restrict__ shapes,
                     const int64_t* __restrict__ lsi, const float* __restrict__ loc,
                     const float* __restrict__ aw, float* __restrict__ out, Dims d) {
  constexpr int D = 32, VEC = 4, G = kTileG, NG = 32 / G;
  static_assert(kSlots * P == G, "one chunk of G samples per level and lane group");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  TileSmem& sm = *reinterpret_cast<TileSmem*>(smem_raw);
  constexpr int kTileOff = ((sizeof(TileSmem) + 127) / 128) * 128;
  unsigned char* tile = smem_raw + kTileOff;     // two buffers of kBufRows rows

  const int MD = d.M * D;
  const int MDb = MD * static_cast<int>(sizeof(float));
  const int L = d.L;
  if (threadIdx.x < L) sm.lvl[threadIdx.x] = load_level(shapes, lsi, threadIdx.x, MD);
  if (STAGE && threadIdx.x < 2 * kRowBytes / 16) {        // the two zero rows
    const int buf = threadIdx.x / (kRowBytes / 16), part = threadIdx.x % (kRowBytes / 16);
    *reinterpret_cast<float4*>(tile + (static_cast<size_t>(buf) * kBufRows + kTileRows) * kRowBytes +
                               part * 16) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int l = 0; l < L; ++l) {
      sm.patch_base[l] = acc;
      acc += ((sm.lvl[l].W + kPatch - 1) / kPatch) * ((sm.lvl[l].H + kPatch - 1) / kPatch);
    }
    sm.patch_base[L] = acc;
  }
  __syncthreads();
  const long long n_items = static_cast<long long>(d.B) * sm.patch_base[L] * d.M;
  const long long my_items =
      blockIdx.x < n_items ? (n_items - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const long long n_steps = my_items * L;

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int gl = lane & (G - 1);
  const int grp = lane / G;
  const int gib = threadIdx.x / G;           // lane group within the block, 0..127
  const int LP = L * P;
  const uint32_t lane_b = gl * 16;
  int4* board = sm.board[warp];
  auto unit_of = [](int j, int grp_, int half) { return j * (2 * NG + 1) + 2 * grp_ + half; };

  // step k of this block <-> global step: items blockIdx.x, blockIdx.x + gridDim.x, ...
  auto global_step = [&](long long k) -> long long {
    if (k >= n_steps) return n_items * L;    // past the end: invalid
    const long long it = k / L;
    return (static_cast<long long>(blockIdx.x) + it * gridDim.x) * L + (k - it * L);
  };

  // ---- staging: window rows of one step, global -> shared, 16 bytes per lane, asynchronous ----
  auto stage = [&](const TileStep& t, int buf) {
    if (!STAGE) return;
    if (t.valid && t.tiled) {
      const LevelInfo lv = sm.lvl[t.l];
      const float* src0 =
          value + (static_cast<int64_t>(t.b) * d.S + lv.start) * MD + t.m * D + gl * VEC;
      unsigned char* dst0 = tile + static_cast<size_t>(buf) * kBufRows * kRowBytes + lane_b;
      const int rows = t.ww * t.wh;
      int yy = gib / t.ww, xx = gib - yy * t.ww;
      for (int r = gib; r < rows; r += kTileGroups) {
        const float* src = src0 + (static_cast<int64_t>(t.y0 + yy) * lv.W + (t.x0 + xx)) * MD;
        __pipeline_memcpy_async(dst0 + r * kRowBytes, src, 16);
        xx += kTileGroups;
        while (xx >= t.ww) { xx -= t.ww; ++yy; }
      }
    }
    __pipeline_commit();
  };

  // location and weight of the sample lane gl resolves in a step: query slot gl/P, point gl%P
  auto load_raw = [&](const TileStep& t) -> RawSample {
    RawSample r;
    r.x = r.y = r.w = 0.f;
    if (!t.valid) return r;
    const int j = (gl / P) * kTileGroups + gib;                 // query within the patch
    const int qx = t.px0 + (j & (kPatch - 1)), qy = t.py0 + j / kPatch;
    const LevelInfo lq = sm.lvl[t.lq];
    if (qx < lq.W && qy < lq.H) {
      const int64_t unit =
          (static_cast<int64_t>(t.b) * d.Q + (lq.start + qy * lq.W + qx)) * d.M + t.m;
      const int s = t.l * P + gl % P;
      const float2 xy = ld_stream_f2(loc + (unit * LP + s) * 2);
      r.x = xy.x; r.y = xy.y;
      r.w = ld_stream_f(aw + unit * LP + s);
    } else {
      r.x = -8.f;                                               // no such query: outside every map
    }
    return r;
  };

  float acc[kSlots][VEC];
  RawSample pre;
  pre.x = pre.y = pre.w = 0.f;
  for (long long k0 = 0; k0 < n_steps; k0 += kStepTable) {
    // descriptors of steps k0 .. k0 + kStepTable (inclusive: the prefetch looks one step ahead)
    __syncthreads();
    for (int i = threadIdx.x; i <= kStepTable; i += blockDim.x) {
      make_tile_step(sm.step[i], sm, global_step(k0 + i), L, d.M, n_items);
      if (!STAGE) sm.step[i].tiled = 0;
    }
    __syncthreads();
    stage(sm.step[0], static_cast<int>(k0 & 1));
    pre = load_raw(sm.step[0]);
    const long long k1 = min(n_steps, k0 + kStepTable);
  for (long long k = k0; k < k1; ++k) {
    if (STAGE) {
      __pipeline_wait_prior(0);
      __syncthreads();                                 // window k ready; buffer (k+1)&1 free again
    }
    const TileStep t = sm.step[k - k0];
    const TileStep nxt = sm.step[k - k0 + 1];
    if (k + 1 < k1) stage(nxt, static_cast<int>((k + 1) & 1));
    const RawSample raw = pre;
    if (k + 1 < k1) pre = load_raw(nxt);               // in flight while this step is sampled

    if (t.l == 0) {
#pragma unroll
      for (int s = 0; s < kSlots; ++s)
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[s][i] = 0.f;
    }
    const LevelInfo lv = sm.lvl[t.l];
    const unsigned char* wbase = tile + static_cast<size_t>(k & 1) * kBufRows * kRowBytes;
    const char* vrow =
        reinterpret_cast<const char*>(value + static_cast<int64_t>(t.b) * d.S * MD + t.m * D);

    {   // lane gl resolves sample gl of the chunk
      const TileRec r = make_tile_rec(raw.x, raw.y, raw.w, lv, MDb, t);
      board[unit_of(gl, grp, 0)] =
          make_int4(r.off, r.rsx, __float_as_int(r.w1), __float_as_int(r.w2));
      *reinterpret_cast<float2*>(&board[unit_of(gl, grp, 1)]) = make_float2(r.w3, r.w4);
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < G; ++j) {
      const int4 q = board[unit_of(j, grp, 0)];
      const float2 w34 = *reinterpret_cast<const float2*>(&board[unit_of(j, grp, 1)]);
      const bool in_win = (q.y & 0x40000000) != 0;
      const uint32_t rs = q.y & 0x3fffffff;
      float4 v1, v2, v3, v4;
      if (!STAGE) {
        // cached global loads; nothing is staged, so the window flag can only mean "out of the
        // map": those samples read the zero row (their strides and weights are 0)
        const bool dead = (q.y & 0x40000000) != 0;
        const uint32_t xs = (q.y < 0) ? static_cast<uint32_t>(MDb) : 0u;
        const char* sp = dead ? reinterpret_cast<const char*>(g_zero_row) : vrow;
        const uint32_t o1 = (dead ? 0u : static_cast<uint32_t>(q.x)) + lane_b;
        v1 = __ldg(reinterpret_cast<const float4*>(sp + o1));
        v2 = __ldg(reinterpret_cast<const float4*>(sp + (o1 + xs)));
        v3 = __ldg(reinterpret_cast<const float4*>(sp + (o1 + rs)));
        v4 = __ldg(reinterpret_cast<const float4*>(sp + (o1 + rs + xs)));
      } else if (__all_sync(0xffffffffu, in_win)) {
        // shared-memory path: 32-bit addresses inside the window buffer
        const uint32_t xs = (q.y < 0) ? static_cast<uint32_t>(kRowBytes) : 0u;
        const unsigned char* p = wbase + (static_cast<uint32_t>(q.x) + lane_b);
        v1 = *reinterpret_cast<const float4*>(p);
        v2 = *reinterpret_cast<const float4*>(p + xs);
        v3 = *reinterpret_cast<const float4*>(p + rs);
        v4 = *reinterpret_cast<const float4*>(p + (rs + xs));
      } else {
        // mixed step: generic loads, window (shared) or value tensor (global) per lane group
        const uint32_t xs = (q.y < 0) ? (in_win ? static_cast<uint32_t>(kRowBytes)
                                                : static_cast<uint32_t>(MDb)) : 0u;
        const char* sp = in_win ? reinterpret_cast<const char*>(wbase) : vrow;
        const uint32_t o1 = static_cast<uint32_t>(q.x) + lane_b;
        v1 = *reinterpret_cast<const float4*>(sp + o1);
        v2 = *reinterpret_cast<const float4*>(sp + (o1 + xs));
        v3 = *reinterpret_cast<const float4*>(sp + (o1 + rs));
        v4 = *reinterpret_cast<const float4*>(sp + (o1 + rs + xs));
      }
      const float w1 = __int_as_float(q.z), w2 = __int_as_float(q.w);
      float* a4 = acc[j / P];
      a4[0] = fmaf(w1, v1.x, a4[0]); a4[1] = fmaf(w1, v1.y, a4[1]);
      a4[2] = fmaf(w1, v1.z, a4[2]); a4[3] = fmaf(w1, v1.w, a4[3]);
      a4[0] = fmaf(w2, v2.x, a4[0]); a4[1] = fmaf(w2, v2.y, a4[1]);
      a4[2] = fmaf(w2, v2.z, a4[2]); a4[3] = fmaf(w2, v2.w, a4[3]);
      a4[0] = fmaf(w34.x, v3.x, a4[0]); a4[1] = fmaf(w34.x, v3.y, a4[1]);
      a4[2] = fmaf(w34.x, v3.z, a4[2]); a4[3] = fmaf(w34.x, v3.w, a4[3]);
      a4[0] = fmaf(w34.y, v4.x, a4[0]); a4[1] = fmaf(w34.y, v4.y, a4[1]);
      a4[2] = fmaf(w34.y, v4.z, a4[2]); a4[3] = fmaf(w34.y, v4.w, a4[3]);
    }
    __syncwarp();

    if (t.l == L - 1) {
      const LevelInfo lq = sm.lvl[t.lq];
#pragma unroll
      for (int s = 0; s < kSlots; ++s) {
        const int j = s * kTileGroups + gib;
        const int qx = t.px0 + (j & (kPatch - 1)), qy = t.py0 + j / kPatch;
        if (qx < lq.W && qy < lq.H) {
          const int64_t q = lq.start + qy * lq.W + qx;
          float* o = out + ((static_cast<int64_t>(t.b) * d.Q + q) * d.M + t.m) * D + gl * VEC;
          *reinterpret_cast<float4*>(o) = make_float4(acc[s][0], acc[s][1], acc[s][2], acc[s][3]);
        }
      }
    }
  }
    if (STAGE) __pipeline_wait_prior(0);
  }
}

// Eligible: the encoder's self-attention geometry.  Q == S is what the host can see; that query i
// IS pixel i of the pyramid is the encoder's construction (get_reference_points,
// opera/models/utils/transformer.py:21173-21188) — if it does not hold the results are still
// right (every query is processed exactly once), only the windows are in the wrong place and the
// samples are fetched from global memory.
bool tile_forward_eligible(const Dims& d, int dtype, int value_dtype) {
  return (tuning().fwd_variant == 5 || tuning().fwd_variant == 6) && dtype == MSDA_F32 && value_dtype == MSDA_F32 && d.D == 32 &&
         d.P == 4 && d.L <= kMaxTileLevels && d.Q == d.S && d.Q >= 1024 && d.B < (1 << 20);
}

cudaError_t launch_forward_tile(const void* value, const int64_t* shapes, const int64_t* lsi,
                                const void* loc, const void* aw, void* out, const Dims& d,
                                int sm_count, cudaStream_t st) {
  const bool staged = tuning().fwd_variant == 5;
  const size_t base = ((sizeof(TileSmem) + 127) / 128) * 128;
  const size_t smem = staged ? base + 2ull * kBufRows * kRowBytes : base;
  static bool opted_in[64][2] = {};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= 64 || !opted_in[dev][staged]) {
    e = staged ? cudaFuncSetAttribute(msda_fwd_tile_kernel<4, true>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem))
               : cudaFuncSetAttribute(msda_fwd_tile_kernel<4, false>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64) opted_in[dev][staged] = true;
  }
  if (staged)
    msda_fwd_tile_kernel<4, true><<<static_cast<unsigned>(sm_count), kTileThreads, smem, st>>>(
        static_cast<const float*>(value), shapes, lsi, static_cast<const float*>(loc),
        static_cast<const float*>(aw), static_cast<float*>(out), d);
  else
    msda_fwd_tile_kernel<4, false><<<static_cast<unsigned>(sm_count), kTileThreads, smem, st>>>(
        static_cast<const float*>(value), shapes, lsi, static_cast<const float*>(loc),
        static_cast<const float*>(aw), static_cast<float*>(out), d);
  note_launches(1);
  note_kernel(KF_FWD_TILE);
  return cudaGetLastError();
}

}  // namespace msda
