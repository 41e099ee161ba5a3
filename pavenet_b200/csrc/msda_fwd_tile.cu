// msda_fwd_tile.cu — forward kernel that STAGES LEVEL TILES IN SHARED MEMORY, for the encoder:
// queries are the pixels of the feature pyramid itself (Q == S), 8 heads x 32 channels, P = 4.
//
// Same maths as msda_fwd.cu (reference: ms_deform_attn_cuda_kernel.cuh:17-64, 200-254).  What
// changes is where the value rows come from.  The rows kernel gathers every bilinear corner
// through L1; with 32 raster-consecutive queries per block 31 % of those gathers miss and pay the
// 64 B/clk L2->SM fill path.  Here a block owns a 16 x 16 PATCH of query pixels of one pyramid
// level and one head, and walks the value levels one after the other: for each level it stages
// the window of value rows the patch can reach — the patch's footprint in that level plus a halo
// of R pixels, at most 27 x 27 rows of 128 bytes — into shared memory with cp.async, double
// buffered so that the window of the next level (or of the next patch) streams in while the
// current one is being sampled, and then takes the four corners of every sample from the window
// with conflict-free 16-byte shared-memory loads (8 lanes = one 128-byte row = all 32 banks).
// A sample whose 2x2 footprint is not completely inside the window — learned offsets are not
// bounded — is fetched from global memory exactly as in the rows kernel: both cases are ONE
// generic 16-byte load per corner (the lane's base pointer is either the window or the value
// tensor), so a warp whose four lane groups disagree does not execute two code paths.
// A (patch, level) pair whose window does not fit the buffer (a level-3 patch covers most of
// level 0) is sampled from global memory altogether.
//
// Block = 512 threads = 64 lane groups of 8; a group owns 4 of the patch's 256 queries and keeps
// their four partial output rows in registers across the levels.  Per level a group has
// 4 queries x P = 16 samples = two chunks of 8: lane i of the group resolves the geometry of
// sample i of the chunk and publishes it on the warp's record board, as in the rows kernel.
// Persistent grid, one block per SM; (patch, head, batch entry) items are strided over the grid
// and flattened with their levels into one sequence of steps for the staging pipeline.
#include <cuda_pipeline_primitives.h>

#include "msda_kernels.h"

namespace msda {

constexpr int kTileThreads = 512;
constexpr int kTileWarps = kTileThreads / 32;
constexpr int kTileG = 8;                          // lanes per fp32 row of 32 channels
constexpr int kTileGroups = kTileThreads / kTileG; // 64
constexpr int kPatch = 16;                         // patch side, in query pixels
constexpr int kSlots = kPatch * kPatch / kTileGroups;   // queries per lane group: 4
constexpr int kMaxTileLevels = 8;
constexpr int kHalo = 5;
constexpr int kTileRows = (kPatch + 1 + 2 * kHalo) * (kPatch + 1 + 2 * kHalo);   // 27 x 27 = 729 rows per buffer
constexpr int kBoardUnits = kTileG * (2 * (32 / kTileG) + 1);

struct TileSmem {
  LevelInfo lvl[kMaxTileLevels];
  int patch_base[kMaxTileLevels + 1];   // first patch index of each query level, per (b, m)
  int4 board[kTileWarps][kBoardUnits];
  // followed by float tile[2][kTileRows * 32]
};

// One step of the flattened (item, level) sequence.
struct TileStep {
  bool valid;
  int64_t b;
  int m, lq, px0, py0;     // query level and patch origin (pixels of that level)
  int l;                   // value level sampled in this step
  int x0, y0, ww, wh;      // window of level l staged in shared memory (pixels of level l)
  bool tiled;
};

__device__ __forceinline__ TileStep tile_step(const TileSmem& sm, long long step, int L, int M,
                                              long long n_items) {
  TileStep t;
  const long long item = step / L;
  t.l = static_cast<int>(step - item * L);
  t.valid = item < n_items;
  t.b = 0; t.m = 0; t.lq = 0; t.px0 = t.py0 = 0; t.x0 = t.y0 = 0; t.ww = t.wh = 0; t.tiled = false;
  if (!t.valid) return t;
  t.m = static_cast<int>(item % M);
  const long long rest = item / M;
  const int per_b = sm.patch_base[L];
  t.b = rest / per_b;
  const int patch = static_cast<int>(rest - t.b * per_b);
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
  return t;
}

template <int P>
__global__ void __launch_bounds__(kTileThreads, 1)
msda_fwd_tile_kernel(const float* __restrict__ value, const int64_t* __restrict__ shapes,
                     const int64_t* __restrict__ lsi, const float* __restrict__ loc,
                     const float* __restrict__ aw, float* __restrict__ out, Dims d) {
  constexpr int D = 32, VEC = 4, G = kTileG, NG = 32 / G;
  static_assert(8 % P == 0, "a chunk of 8 samples must hold whole queries");
  constexpr int QPC = 8 / P;                 // queries per chunk
  constexpr int CHUNKS = kSlots / QPC;       // chunks per level and lane group
  static_assert(kSlots % QPC == 0, "");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  TileSmem& sm = *reinterpret_cast<TileSmem*>(smem_raw);
  float* tile = reinterpret_cast<float*>(smem_raw + ((sizeof(TileSmem) + 127) / 128) * 128);

  const int MD = d.M * D;
  const int L = d.L;
  if (threadIdx.x < L) sm.lvl[threadIdx.x] = load_level(shapes, lsi, threadIdx.x, MD);
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
  const long long my_items = blockIdx.x < n_items ? (n_items - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const long long n_steps = my_items * L;

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int gl = lane & (G - 1);
  const int grp = lane / G;
  const int gib = threadIdx.x / G;           // lane group within the block, 0..63
  const int LP = L * P;
  const uint32_t lane_b = gl * 16;
  const uint32_t MDb = MD * sizeof(float);
  int4* board = sm.board[warp];
  auto unit_of = [](int j, int grp_, int half) { return j * (2 * NG + 1) + 2 * grp_ + half; };

  // step k of this block <-> global step: items blockIdx.x, blockIdx.x + gridDim.x, ...
  auto global_step = [&](long long k) -> long long {
    if (k >= n_steps) return n_items * L;    // invalid
    const long long it = k / L;
    return (static_cast<long long>(blockIdx.x) + it * gridDim.x) * L + (k - it * L);
  };

  // ---- staging: window rows of one step, global -> shared, 16 bytes per lane, asynchronous ----
  auto stage = [&](const TileStep& t, int buf) {
    if (t.valid && t.tiled) {
      const LevelInfo lv = sm.lvl[t.l];
      const float* src0 = value + (t.b * d.S + lv.start) * MD + t.m * D + gl * VEC;
      float* dst0 = tile + static_cast<size_t>(buf) * kTileRows * D + gl * VEC;
      const int rows = t.ww * t.wh;
      int yy = gib / t.ww, xx = gib - yy * t.ww;
      for (int r = gib; r < rows; r += kTileGroups) {
        const float* src = src0 + (static_cast<int64_t>(t.y0 + yy) * lv.W + (t.x0 + xx)) * MD;
        __pipeline_memcpy_async(dst0 + r * D, src, 16);
        xx += kTileGroups;
        while (xx >= t.ww) { xx -= t.ww; ++yy; }
      }
    }
    __pipeline_commit();
  };

  // location and weight of the sample lane gl resolves in chunk c of a step
  auto load_raw = [&](const TileStep& t, int c) -> RawSample {
    RawSample r;
    r.x = r.y = r.w = 0.f;
    if (!t.valid) return r;
    const int slot = c * QPC + gl / P;
    const int j = slot * kTileGroups + gib;
    const int qx = t.px0 + (j & (kPatch - 1)), qy = t.py0 + j / kPatch;
    const LevelInfo lq = sm.lvl[t.lq];
    if (qx < lq.W && qy < lq.H) {
      const int64_t unit = (t.b * d.Q + (lq.start + qy * lq.W + qx)) * d.M + t.m;
      const int s = t.l * P + gl % P;
      const float2 xy = ld_stream_f2(loc + (unit * LP + s) * 2);
      r.x = xy.x; r.y = xy.y;
      r.w = ld_stream_f(aw + unit * LP + s);
    }
    return r;
  };

  float acc[kSlots][VEC];
  TileStep cur = tile_step(sm, global_step(0), L, d.M, n_items);
  stage(cur, 0);
  RawSample pre[CHUNKS];
#pragma unroll
  for (int c = 0; c < CHUNKS; ++c) pre[c] = load_raw(cur, c);
  for (long long k = 0; k < n_steps; ++k) {
    __pipeline_wait_prior(0);
    __syncthreads();                                   // window k ready; buffer (k+1)&1 free again
    const TileStep nxt = tile_step(sm, global_step(k + 1), L, d.M, n_items);
    stage(nxt, static_cast<int>((k + 1) & 1));
    const TileStep t = cur;
    cur = nxt;
    RawSample raw[CHUNKS];
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) {
      raw[c] = pre[c];
      pre[c] = load_raw(nxt, c);                       // in flight while this step is sampled
    }

    if (t.l == 0) {
#pragma unroll
      for (int s = 0; s < kSlots; ++s)
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[s][i] = 0.f;
    }
    const LevelInfo lq = sm.lvl[t.lq];
    const LevelInfo lv = sm.lvl[t.l];
    const char* vrow = reinterpret_cast<const char*>(value + t.b * d.S * MD + t.m * D);
    const char* wbase = reinterpret_cast<const char*>(tile + static_cast<size_t>(k & 1) * kTileRows * D);
    const uint32_t wstride = static_cast<uint32_t>(t.ww) * D * sizeof(float);

#pragma unroll
    for (int c = 0; c < CHUNKS; ++c) {
      // ---- lane gl resolves sample gl of the chunk: query slot c*QPC + gl/P, point gl%P ----
      {
        const int slot = c * QPC + gl / P;
        const int p = gl % P;
        const int j = slot * kTileGroups + gib;               // query within the patch
        const int qx = t.px0 + (j & (kPatch - 1)), qy = t.py0 + j / kPatch;
        FwdRec r;
        r.off = kDeadOff; r.rsx = 0; r.w1 = r.w2 = r.w3 = r.w4 = 0.f;
        int space = 0;                                          // 1: the window in shared memory
        (void)p;
        if (qx < lq.W && qy < lq.H) {
          r = make_fwd_rec<sizeof(float)>(raw[c].x, raw[c].y, raw[c].w, lv, MD);
          if (r.off != kDeadOff && t.tiled) {
            // anchor pixel and extents of the (re-anchored) footprint
            const int pix = r.off / static_cast<int>(MDb) - lv.start;
            const int h0 = pix / lv.W, w0 = pix - h0 * lv.W;
            const int dx = (r.rsx < 0) ? 1 : 0;                 // bit 31: second column exists
            const int dy = (r.rsx & 0x7fffffff) ? 1 : 0;
            if (w0 >= t.x0 && w0 + dx < t.x0 + t.ww && h0 >= t.y0 && h0 + dy < t.y0 + t.wh) {
              space = 1;
              r.off = ((h0 - t.y0) * t.ww + (w0 - t.x0)) * D * static_cast<int>(sizeof(float));
              r.rsx = (dy ? static_cast<int>(wstride) : 0) | (dx ? static_cast<int>(0x80000000u) : 0);
            }
          }
        }
        // bit 30 of rsx: the offsets address the shared-memory window
        if (space) r.rsx |= 0x40000000;
        board[unit_of(gl, grp, 0)] =
            make_int4(r.off, r.rsx, __float_as_int(r.w1), __float_as_int(r.w2));
        *reinterpret_cast<float2*>(&board[unit_of(gl, grp, 1)]) = make_float2(r.w3, r.w4);
      }
      __syncwarp();
#pragma unroll
      for (int j = 0; j < G; ++j) {
        const int4 q = board[unit_of(j, grp, 0)];
        const float2 w34 = *reinterpret_cast<const float2*>(&board[unit_of(j, grp, 1)]);
        const bool alive = q.x != kDeadOff;
        const bool in_win = (q.y & 0x40000000) != 0;
        const uint32_t rs = q.y & 0x3fffffff;
        const uint32_t xs = (q.y < 0) ? (in_win ? static_cast<uint32_t>(D * sizeof(float)) : MDb) : 0u;
        const char* sp = !alive ? reinterpret_cast<const char*>(g_zero_row) : (in_win ? wbase : vrow);
        const uint32_t o1 = (alive ? static_cast<uint32_t>(q.x) : 0u) + lane_b;
        // generic 16-byte loads: the window (shared) or the value tensor (global), per lane group
        const float4 v1 = *reinterpret_cast<const float4*>(sp + o1);
        const float4 v2 = *reinterpret_cast<const float4*>(sp + (o1 + xs));
        const float4 v3 = *reinterpret_cast<const float4*>(sp + (o1 + rs));
        const float4 v4 = *reinterpret_cast<const float4*>(sp + (o1 + rs + xs));
        const float w1 = __int_as_float(q.z), w2 = __int_as_float(q.w);
        float* a4 = acc[c * QPC + j / P];
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
    }

    if (t.l == L - 1) {
#pragma unroll
      for (int s = 0; s < kSlots; ++s) {
        const int j = s * kTileGroups + gib;
        const int qx = t.px0 + (j & (kPatch - 1)), qy = t.py0 + j / kPatch;
        if (qx < lq.W && qy < lq.H) {
          const int64_t q = lq.start + qy * lq.W + qx;
          float* o = out + ((t.b * d.Q + q) * d.M + t.m) * D + gl * VEC;
          *reinterpret_cast<float4*>(o) = make_float4(acc[s][0], acc[s][1], acc[s][2], acc[s][3]);
        }
      }
    }
  }
  __pipeline_wait_prior(0);
}

// Eligible: the encoder's self-attention geometry.  Q == S is what the host can see; that query i
// IS pixel i of the pyramid is the encoder's construction (get_reference_points,
// opera/models/utils/transformer.py:21173-21188) — if it does not hold the results are still
// right (every query is processed exactly once), only the windows are in the wrong place and the
// samples are fetched from global memory.
bool tile_forward_eligible(const Dims& d, int dtype, int value_dtype) {
  return tuning().fwd_variant == 5 && dtype == MSDA_F32 && value_dtype == MSDA_F32 && d.D == 32 &&
         d.P == 4 && d.L <= kMaxTileLevels && d.Q == d.S && d.Q >= 1024;
}

cudaError_t launch_forward_tile(const void* value, const int64_t* shapes, const int64_t* lsi,
                                const void* loc, const void* aw, void* out, const Dims& d,
                                int sm_count, cudaStream_t st) {
  const size_t smem = ((sizeof(TileSmem) + 127) / 128) * 128 + 2ull * kTileRows * 32 * sizeof(float);
  static bool opted_in[64] = {};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= 64 || !opted_in[dev]) {
    e = cudaFuncSetAttribute(msda_fwd_tile_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64) opted_in[dev] = true;
  }
  msda_fwd_tile_kernel<4><<<static_cast<unsigned>(sm_count), kTileThreads, smem, st>>>(
      static_cast<const float*>(value), shapes, lsi, static_cast<const float*>(loc),
      static_cast<const float*>(aw), static_cast<float*>(out), d);
  note_launches(1);
  note_kernel(KF_FWD_TILE);
  return cudaGetLastError();
}

}  // namespace msda
