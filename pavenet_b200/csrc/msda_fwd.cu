// msda_fwd.cu — forward kernels of multi-scale deformable attention sampling
// for sm_100a.
//
// What it computes (reference: ms_deform_attn_cuda_kernel.cuh:200-254 with the
// bilinear helper at :17-64; CPU oracle multi_scale_deform_attn.py:92-149):
//   out[b,q,m,c] = sum_{l,p} a[b,q,m,l,p] * bilinear(value_l[b,:,m,c]; loc[b,q,m,l,p])
//
// Two kernel families, both new designs (the reference maps one thread to one
// output scalar and recomputes the sample geometry per channel):
//
//  * rows<D,VT,SPLIT>: a group of G = D*sizeof(VT)/16 lanes owns one (b,q,m)
//    row and moves each bilinear corner with one 16-byte load per lane.  The
//    sample geometry (floor, validity, offsets) is computed once per sample by
//    one lane and handed to the group through a per-warp shared-memory record,
//    so the ALU cost is paid once per sample, not once per channel.  SPLIT
//    groups can share one row (disjoint sample ranges, shuffle-reduced at the
//    end) so small-Q pose-decoder shapes still fill the machine.
//  * generic<T,VT>: any D, fp32/fp64 — the parity path for the shapes the
//    reference's tests use (D = 2, 4, 30, 71, 1025, double precision).
#include <atomic>
#include <type_traits>

#include "msda_kernels.h"

namespace msda {

// --------------------------------------------------------------------------
// generic kernel: one thread per output scalar
// --------------------------------------------------------------------------
template <typename T>
struct Cvt {
  template <typename VT>
  __device__ __forceinline__ static T from(VT v) { return static_cast<T>(v); }
  __device__ __forceinline__ static T from(__nv_bfloat16 v) {
    return static_cast<T>(__bfloat162float(v));
  }
};

template <typename T, typename VT>
__global__ void __launch_bounds__(256)
msda_fwd_generic_kernel(const VT* __restrict__ value, const int64_t* __restrict__ shapes,
                        const int64_t* __restrict__ lsi, const T* __restrict__ loc,
                        const T* __restrict__ aw, T* __restrict__ out, Dims d) {
  const int64_t total = static_cast<int64_t>(d.B) * d.Q * d.M * d.D;
  const int64_t MD = static_cast<int64_t>(d.M) * d.D;
  for (int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(idx % d.D);
    const int64_t unit = idx / d.D;  // (b*Q + q)*M + m
    const int m = static_cast<int>(unit % d.M);
    const int64_t b = unit / d.M / d.Q;
    const VT* vb = value + b * d.S * MD + static_cast<int64_t>(m) * d.D + c;
    const T* lp = loc + unit * d.L * d.P * 2;
    const T* ap = aw + unit * d.L * d.P;
    T acc = 0;
    for (int l = 0; l < d.L; ++l) {
      const int H = static_cast<int>(shapes[2 * l]);
      const int W = static_cast<int>(shapes[2 * l + 1]);
      const VT* vl = vb + lsi[l] * MD;
      for (int p = 0; p < d.P; ++p) {
        const T x = lp[0], y = lp[1], a = ap[0];
        lp += 2; ap += 1;
        const T h_im = y * H - T(0.5), w_im = x * W - T(0.5);
        if (h_im > T(-1) && w_im > T(-1) && h_im < T(H) && w_im < T(W)) {
          const T hf = floor(h_im), wf = floor(w_im);
          const int h0 = static_cast<int>(hf), w0 = static_cast<int>(wf);
          const T lh = h_im - hf, lw = w_im - wf, hh = T(1) - lh, hw = T(1) - lw;
          const bool r0 = h0 >= 0, r1 = h0 + 1 <= H - 1, c0 = w0 >= 0, c1 = w0 + 1 <= W - 1;
          const int64_t o00 = (static_cast<int64_t>(h0) * W + w0) * MD;
          const T v1 = (r0 && c0) ? Cvt<T>::from(vl[o00]) : T(0);
          const T v2 = (r0 && c1) ? Cvt<T>::from(vl[o00 + MD]) : T(0);
          const T v3 = (r1 && c0) ? Cvt<T>::from(vl[o00 + W * MD]) : T(0);
          const T v4 = (r1 && c1) ? Cvt<T>::from(vl[o00 + W * MD + MD]) : T(0);
          const T val = (hh * hw) * v1 + (hh * lw) * v2 + (lh * hw) * v3 + (lh * lw) * v4;
          acc += val * a;
        }
      }
    }
    out[idx] = acc;
  }
}

// --------------------------------------------------------------------------
// rows kernel
// --------------------------------------------------------------------------
constexpr int kRowsThreads = 256;
constexpr int kRowsWarps = kRowsThreads / 32;
// Resident blocks per SM the compiler must allow: 4 (<= 64 registers) for the
// large-Q shapes, which are bound by the L1 data pipe and want warps; 3 (<= 80
// registers, more gathers in flight per lane) for the split small-Q shapes,
// which are latency bound (measured on B200: pose cfg3 58 -> 48 us).
constexpr int fwd_min_blocks(int split) { return split > 1 ? 3 : 4; }
constexpr int kRowsZeroTile16 = 256;   // 4 KiB zero tile for the TMA zero-fill (16-byte units)

// Shared-memory state of one rows block.
template <int D, typename VT, int SPLIT, int LB = 16>
struct FwdRowsSmem {
  static constexpr int VEC = RowLoad<VT, LB>::VEC;
  static constexpr int G = D / VEC;
  static constexpr int NGW = 32 / G;
  static constexpr int WSPLIT = SPLIT < NGW ? SPLIT : NGW;
  static constexpr int XSPLIT = SPLIT / WSPLIT;
  LevelInfo lvl[kMaxSmemLevels];
  int4 board[kRowsWarps][G * (2 * (32 / G) + 1)];
  float part[XSPLIT > 1 ? kRowsWarps : 1][D];   // per-warp partial rows (XSPLIT > 1)
};

// One work item of the rows family: (batch entry b, chunk of consecutive queries, head m).
// All rows of an item belong to ONE head and to neighbouring queries, so when neighbouring
// queries look at neighbouring pixels (the encoder) their bilinear corners are the same
// 128-byte rows and hit in L1.
template <int D, typename VT, int SPLIT, class SRC, int LB = 16>
__device__ __forceinline__ void fwd_rows_item(const VT* __restrict__ value, SRC src,
                                              float* __restrict__ out, const Dims& d,
                                              FwdRowsSmem<D, VT, SPLIT, LB>& sm, int m, int chunk,
                                              int64_t b) {
  using VL = RowLoad<VT, LB>;
  constexpr int VEC = VL::VEC;
  constexpr int G = D / VEC;  // lanes per row
  static_assert(D % VEC == 0 && G >= 1 && G <= 32 && (G & (G - 1)) == 0, "bad D");
  constexpr int NGW = 32 / G;                            // row groups per warp
  constexpr int WSPLIT = SPLIT < NGW ? SPLIT : NGW;      // splits combined by shuffles inside a warp
  constexpr int XSPLIT = SPLIT / WSPLIT;                 // ... and across warps through shared memory
  static_assert(SPLIT * G <= kRowsThreads && (SPLIT & (SPLIT - 1)) == 0, "bad SPLIT");
  auto& s_lvl = sm.lvl;
  auto& s_board = sm.board;
  auto& s_part = sm.part;
  const int MD = d.M * D;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int gl = lane & (G - 1);        // lane within the row group
  const int grp = lane / G;             // group within the warp
  constexpr int QPB = kRowsThreads / G / SPLIT;  // queries per block
  const int gib = threadIdx.x / G;      // group within the block
  const int split = gib % SPLIT;
  int q_idx = chunk * QPB + gib / SPLIT;
  const bool live = q_idx < d.Q;
  if (!live) q_idx = d.Q - 1;           // keep the warp converged; result discarded
  const int64_t unit = (b * d.Q + q_idx) * d.M + m;

  // block-uniform base (batch entry, head) + this lane's 16-byte slot
  const char* vrow = reinterpret_cast<const char*>(value + b * d.S * MD + m * D);
  const uint32_t lane_b = gl * LB;
  const uint32_t MDb = MD * sizeof(VT);
  const int LP = d.L * d.P;
  src.bind(unit, LP, d.M, b * d.Q + q_idx);
  src.template prepass<G>(LP, gl, live && split == 0);  // fused: softmax max / sum of the row

  // sample range of this split, in chunks of G
  const int per = ((LP + SPLIT - 1) / SPLIT + G - 1) / G * G;
  const int s_begin = split * per;
  const int s_end = min(LP, s_begin + per);

  float acc[VEC];
#pragma unroll
  for (int c = 0; c < VEC; ++c) acc[c] = 0.f;

  // Per-warp record board: slot (j, grp) holds the record of sample j of row
  // group grp as two 16-byte halves; half h lives at 16-byte unit
  //   j*(2*NG + 1) + 2*grp + h
  // The odd row pitch makes the 8 lanes of a quarter-warp that store 8
  // different samples, and the NG groups that load the same sample slot, touch
  // distinct bank groups, with compile-time offsets in j.
  constexpr int NG = 32 / G;  // row groups per warp
  int4* board = s_board[warp];
  auto unit_of = [](int j, int grp_, int half) { return j * (2 * NG + 1) + 2 * grp_ + half; };
  const FastDivP level_of(d.P);

  // locations / weights of the chunk after the current one are fetched while
  // the current chunk is being gathered (they come from DRAM)
  RawSample nxt;
  nxt.x = nxt.y = nxt.w = 0.f;
  {
    const int s = s_begin + gl;
    if (s < s_end) nxt = src.load(s);
  }
  for (int s0 = s_begin; s0 < s_begin + per; s0 += G) {
    // --- one lane per sample resolves the geometry and the four weights ---
    {
      const int s = s0 + gl;
      FwdRec r;
      r.off = kDeadOff; r.rsx = 0; r.w1 = r.w2 = r.w3 = r.w4 = 0.f;
      if (s < s_end) {
        const int l = level_of(s);
        const LevelInfo lv = s_lvl[l];
        RawSample cur = nxt;
        src.finish(cur, s, l, lv);
        r = make_fwd_rec<sizeof(VT)>(cur.x, cur.y, cur.w, lv, MD);
      }
      board[unit_of(gl, grp, 0)] =
          make_int4(r.off, r.rsx, __float_as_int(r.w1), __float_as_int(r.w2));
      *reinterpret_cast<float2*>(&board[unit_of(gl, grp, 1)]) = make_float2(r.w3, r.w4);
      const int sn = s + G;
      if (sn < s_end) nxt = src.load(sn);
    }
    __syncwarp();
    // --- the whole group gathers each sample of the chunk ---
#pragma unroll
    for (int j = 0; j < G; ++j) {
      // (broadcasting the record with 6 shuffles instead was measured: 0.241 vs 0.234 ms)
      const int4 q = board[unit_of(j, grp, 0)];
      const float2 w34 = *reinterpret_cast<const float2*>(&board[unit_of(j, grp, 1)]);
      const bool alive = q.x != kDeadOff;
      const uint32_t rs = q.y & 0x7fffffff;
      const uint32_t xs = (q.y >> 31) & MDb;  // one pixel to the right, or 0 for a duplicate
      // out-of-map samples read the zero row (their strides are 0, their weights 0)
      const char* src = alive ? vrow : reinterpret_cast<const char*>(g_zero_row);
      const uint32_t o1 = (alive ? static_cast<uint32_t>(q.x) : 0u) + lane_b;
      float v1[VEC], v2[VEC], v3[VEC], v4[VEC];
      VL::load(reinterpret_cast<const VT*>(src + o1), v1);
      VL::load(reinterpret_cast<const VT*>(src + (o1 + xs)), v2);
      VL::load(reinterpret_cast<const VT*>(src + (o1 + rs)), v3);
      VL::load(reinterpret_cast<const VT*>(src + (o1 + rs + xs)), v4);
      const float w1 = __int_as_float(q.z), w2 = __int_as_float(q.w);
#pragma unroll
      for (int c = 0; c < VEC; ++c) {
        acc[c] = fmaf(w1, v1[c], acc[c]);
        acc[c] = fmaf(w2, v2[c], acc[c]);
        acc[c] = fmaf(w34.x, v3[c], acc[c]);
        acc[c] = fmaf(w34.y, v4[c], acc[c]);
      }
    }
    __syncwarp();
  }

  // --- combine the SPLIT partial rows: adjacent groups of a warp by shuffles ... ---
#pragma unroll
  for (int off = G; off < G * WSPLIT; off <<= 1) {
#pragma unroll
    for (int c = 0; c < VEC; ++c) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], off);
  }
  // --- ... and, when a row is spread over XSPLIT warps, those through shared memory ---
  if (XSPLIT > 1) {
    if (grp == 0) {
#pragma unroll
      for (int c = 0; c < VEC; ++c) s_part[warp][gl * VEC + c] = acc[c];
    }
    __syncthreads();
    if (grp == 0 && (warp % XSPLIT) == 0) {
#pragma unroll
      for (int x = 1; x < XSPLIT; ++x) {
#pragma unroll
        for (int c = 0; c < VEC; ++c) acc[c] += s_part[warp + x][gl * VEC + c];
      }
    }
  }
  if (live && split == 0) {
    float* o = out + unit * D + gl * VEC;
#pragma unroll
    for (int c = 0; c < VEC; c += 4)
      *reinterpret_cast<float4*>(o + c) = make_float4(acc[c], acc[c + 1], acc[c + 2], acc[c + 3]);
  }
}

template <int D, typename VT, int SPLIT, class SRC, int LB = 16, int MINB = (LB == 32 ? 3 : fwd_min_blocks(SPLIT))>
__global__ void __launch_bounds__(kRowsThreads, MINB)
msda_fwd_rows_kernel(const VT* __restrict__ value, const int64_t* __restrict__ shapes,
                     const int64_t* __restrict__ lsi, SRC src, float* __restrict__ out, Dims d,
                     uint4* __restrict__ clear = nullptr, long long clear_n16 = 0) {
  __shared__ FwdRowsSmem<D, VT, SPLIT, LB> sm;
  // optional zero-fill of the coming backward's grad_value (msda_forward_clear, knob clear_mode = 2):
  // each block hands its share to the TMA engine as bulk stores of a zeroed shared-memory tile, so the
  // fill costs no LSU slots and drains into DRAM, which this L1-bound kernel leaves ~85 % idle
  __shared__ __align__(128) uint4 s_zero[kRowsZeroTile16];
  const int MD = d.M * D;
  for (int l = threadIdx.x; l < d.L; l += blockDim.x) sm.lvl[l] = load_level(shapes, lsi, l, MD);
  if (clear) {
    for (int i = threadIdx.x; i < kRowsZeroTile16; i += blockDim.x) s_zero[i] = make_uint4(0u, 0u, 0u, 0u);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (clear && threadIdx.x == 0) {
    const long long a = clear_n16 * blockIdx.x / gridDim.x, b = clear_n16 * (blockIdx.x + 1LL) / gridDim.x;
    const uint32_t src_addr = static_cast<uint32_t>(__cvta_generic_to_shared(s_zero));
    for (long long i = a; i < b; i += kRowsZeroTile16) {
      const uint32_t bytes = static_cast<uint32_t>(min(static_cast<long long>(kRowsZeroTile16), b - i) * 16);
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(clear + i),
                   "r"(src_addr), "r"(bytes)
                   : "memory");
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  }
  constexpr int QPB = kRowsThreads / (D / RowLoad<VT, LB>::VEC) / SPLIT;
  const int n_chunks = (d.Q + QPB - 1) / QPB;
  int blk = blockIdx.x;
  const int m = blk % d.M;
  blk /= d.M;
  fwd_rows_item<D, VT, SPLIT, SRC, LB>(value, src, out, d, sm, m, blk % n_chunks, blk / n_chunks);
  if (clear && threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// Forward variant 2 (large Q): persistent blocks, head-affine.  Every block of an SM serves
// the SAME head -- head = %smid mod M -- pulling (batch entry, query chunk) items of that head
// from a per-head counter in global memory and helping the other heads when its own is done.
// The coarse levels of one head (35 + 134 KB in fp32 for R-50 at 800x1333) then stay resident
// in that SM's L1 instead of competing with the other seven heads' copies.
__device__ int g_head_queue[256][64];   // [slot][head]: a launch zeroes and uses one slot

template <int D, typename VT, class SRC>
__global__ void __launch_bounds__(kRowsThreads, fwd_min_blocks(1))
msda_fwd_rows_affine_kernel(const VT* __restrict__ value, const int64_t* __restrict__ shapes,
                            const int64_t* __restrict__ lsi, SRC src, float* __restrict__ out,
                            Dims d, int slot) {
  __shared__ FwdRowsSmem<D, VT, 1> sm;
  __shared__ int s_item;
  const int MD = d.M * D;
  for (int l = threadIdx.x; l < d.L; l += blockDim.x) sm.lvl[l] = load_level(shapes, lsi, l, MD);
  constexpr int QPB = kRowsThreads / (D / Vec16<VT>::VEC);
  const int n_chunks = (d.Q + QPB - 1) / QPB;
  const int per_head = n_chunks * d.B;
  unsigned smid;
  asm("mov.u32 %0, %%smid;" : "=r"(smid));
  const int home = static_cast<int>(smid % static_cast<unsigned>(d.M));
  int* queue = g_head_queue[slot];
  for (;;) {
    __syncthreads();                      // level table ready / previous item fully consumed
    if (threadIdx.x == 0) {
      int item = -1;
      for (int t = 0, h = home; t < d.M; ++t, h = (h + 1 == d.M ? 0 : h + 1)) {
        const int c = atomicAdd(&queue[h], 1);
        if (c < per_head) {
          item = h * per_head + c;
          break;
        }
      }
      s_item = item;
    }
    __syncthreads();
    const int item = s_item;
    if (item < 0) break;
    const int m = item / per_head;
    const int c = item - m * per_head;
    fwd_rows_item<D, VT, 1, SRC>(value, src, out, d, sm, m, c % n_chunks, c / n_chunks);
  }
}

// --------------------------------------------------------------------------
// launchers
// --------------------------------------------------------------------------
static thread_local int g_fwd_sm_count = 148;   // set by launch_forward* before dispatch
// zero-fill folded into the default rows kernel (clear_mode = 2): set by launch_forward*, consumed by
// launch_rows_split (cleared there so that no later launch repeats it)
static thread_local void* g_fwd_clear = nullptr;
static thread_local size_t g_fwd_clear_bytes = 0;

template <int D, typename VT, int SPLIT, class SRC>
static cudaError_t launch_rows_split(const void* value, const int64_t* shapes, const int64_t* lsi,
                                     const SRC& src, float* out, const Dims& d, cudaStream_t st) {
  constexpr int G = D / Vec16<VT>::VEC;
  constexpr int QPB = kRowsThreads / G / SPLIT;
  // take the pending zero-fill (if any) before anything can return: it must never outlive this call
  uint4* const fold_clear = static_cast<uint4*>(g_fwd_clear);
  const long long fold_n16 = static_cast<long long>(g_fwd_clear_bytes / 16);
  g_fwd_clear = nullptr;
  g_fwd_clear_bytes = 0;
  const int64_t blocks = static_cast<int64_t>(d.B) * ((d.Q + QPB - 1) / QPB) * d.M;
  if (blocks >= (int64_t(1) << 31)) return cudaErrorInvalidConfiguration;
  // variant 3: 256-bit value loads (fp32, 32-byte aligned rows), 8 rows per warp instruction
  if constexpr (SPLIT == 1 && std::is_same<VT, float>::value && D >= 32) {
    if ((tuning().fwd_variant == 3 || tuning().fwd_variant == 4) &&
        (reinterpret_cast<uintptr_t>(value) & 31u) == 0) {
      constexpr int QPBW = kRowsThreads / (D / 8);
      const int64_t wblocks = static_cast<int64_t>(d.B) * ((d.Q + QPBW - 1) / QPBW) * d.M;
      if (tuning().fwd_variant == 4)      // 64 registers, four blocks per SM
        msda_fwd_rows_kernel<D, VT, 1, SRC, 32, 4><<<static_cast<unsigned>(wblocks), kRowsThreads, 0, st>>>(
            static_cast<const VT*>(value), shapes, lsi, src, out, d);
      else
        msda_fwd_rows_kernel<D, VT, 1, SRC, 32, 3><<<static_cast<unsigned>(wblocks), kRowsThreads, 0, st>>>(
            static_cast<const VT*>(value), shapes, lsi, src, out, d);
      note_launches(1);
      note_kernel(std::is_same<SRC, FusedSource>::value ? KF_FWD_ROWS_FUSED : KF_FWD_ROWS);
      return cudaGetLastError();
    }
  }
  if constexpr (SPLIT == 1) {
    const int resident = g_fwd_sm_count * fwd_min_blocks(1);
    if (tuning().fwd_variant == 2 && d.M <= 64 && blocks > 2 * resident) {
      static std::atomic<unsigned> next_slot{0};
      const int slot = static_cast<int>(next_slot.fetch_add(1) % 256u);
      int* q = nullptr;
      cudaError_t e = cudaGetSymbolAddress(reinterpret_cast<void**>(&q), g_head_queue);
      if (e != cudaSuccess) return e;
      e = cudaMemsetAsync(q + slot * 64, 0, 64 * sizeof(int), st);
      if (e != cudaSuccess) return e;
      msda_fwd_rows_affine_kernel<D, VT, SRC><<<static_cast<unsigned>(resident), kRowsThreads, 0, st>>>(
          static_cast<const VT*>(value), shapes, lsi, src, out, d, slot);
      note_launches(1);
      note_kernel(std::is_same<SRC, FusedSource>::value ? KF_FWD_ROWS_FUSED : KF_FWD_ROWS);
      return cudaGetLastError();
    }
  }
  msda_fwd_rows_kernel<D, VT, SPLIT, SRC><<<static_cast<unsigned>(blocks), kRowsThreads, 0, st>>>(
      static_cast<const VT*>(value), shapes, lsi, src, out, d, fold_clear, fold_n16);
  note_launches(1);
  note_kernel(std::is_same<SRC, FusedSource>::value ? KF_FWD_ROWS_FUSED : KF_FWD_ROWS);
  return cudaGetLastError();
}

template <int D, typename VT, class SRC>
static cudaError_t launch_rows(const void* value, const int64_t* shapes, const int64_t* lsi,
                               const SRC& src, float* out, const Dims& d, int split,
                               cudaStream_t st) {
  constexpr int G = D / Vec16<VT>::VEC;
  constexpr int MAXS = (kRowsThreads / G) < 32 ? (kRowsThreads / G) : 32;  // up to a whole block per row
#define MSDA_TRY_SPLIT(S)                                                                          \
  if (split >= S && MAXS >= S)                                                                     \
    return launch_rows_split<D, VT, (MAXS >= S ? S : 1), SRC>(value, shapes, lsi, src, out, d, st);
  MSDA_TRY_SPLIT(32)
  MSDA_TRY_SPLIT(16)
  MSDA_TRY_SPLIT(8)
  MSDA_TRY_SPLIT(4)
  MSDA_TRY_SPLIT(2)
#undef MSDA_TRY_SPLIT
  return launch_rows_split<D, VT, 1, SRC>(value, shapes, lsi, src, out, d, st);
}

// How many groups should share one row: enough that the grid covers the
// machine a few times over, never more than the samples can feed.
int choose_split(const Dims& d, int G, int sm_count) {
  if (tuning().fwd_split > 0) return tuning().fwd_split;
  const int64_t units = static_cast<int64_t>(d.B) * d.Q * d.M;
  // aim for ~48 resident warps per SM, but keep at least two chunks of G samples per group
  const int64_t want_groups = static_cast<int64_t>(sm_count) * 48 * (32 / G);
  const int LP = d.L * d.P;
  int split = 1;
  while (split < 32 && units * split < want_groups && LP / (split * 2) >= 2 * G) split *= 2;
  return split;
}

bool rows_supported(int D, int value_dtype) {
  if (value_dtype != MSDA_F32 && value_dtype != MSDA_BF16) return false;
  return D == 16 || D == 32 || D == 64;
}

// zero-fill of the optional `clear` buffer for the kernel families that do not fold it in
static cudaError_t clear_by_memset(void* clear, size_t clear_bytes, cudaStream_t st) {
  if (!clear || !clear_bytes) return cudaSuccess;
  return cudaMemsetAsync(clear, 0, clear_bytes, st);
}

// the default rows kernel can take the zero-fill along (TMA bulk stores) when it is 16-byte granular
// and no alternative forward variant is selected
static bool rows_fold_clear(const void* clear, size_t clear_bytes) {
  return tuning().clear_mode == 2 && tuning().fwd_variant == 0 && clear && clear_bytes &&
         clear_bytes % 16 == 0 && (reinterpret_cast<uintptr_t>(clear) & 15u) == 0;
}

cudaError_t launch_forward(const void* value, const int64_t* shapes, const int64_t* lsi,
                           const void* loc, const void* aw, void* out, const Dims& d, int dtype,
                           int value_dtype, int sm_count, int force_generic, void* clear,
                           size_t clear_bytes, cudaStream_t st) {
  g_fwd_sm_count = sm_count;
  g_fwd_clear = nullptr;
  g_fwd_clear_bytes = 0;
  if (dtype == MSDA_F32 && !force_generic && d.L <= kMaxSmemLevels &&
      rows_supported(d.D, value_dtype)) {
    PlainSource src;
    src.loc = static_cast<const float*>(loc);
    src.aw = static_cast<const float*>(aw);
    float* outf = static_cast<float*>(out);
    if (tile_forward_eligible(d, dtype, value_dtype) && (reinterpret_cast<uintptr_t>(value) & 15u) == 0) {
      const cudaError_t ce = clear_by_memset(clear, clear_bytes, st);
      if (ce != cudaSuccess) return ce;
      return launch_forward_tile(value, shapes, lsi, loc, aw, out, d, sm_count, st);
    }
    if (flat_preferred(d, d.D / (value_dtype == MSDA_F32 ? 4 : 8), sm_count)) {
      // the persistent kernel folds the zero-fill in when it can be done in 16-byte stores
      const bool fold = clear && clear_bytes % 16 == 0 && (reinterpret_cast<uintptr_t>(clear) & 15u) == 0;
      if (!fold) {
        const cudaError_t ce = clear_by_memset(clear, clear_bytes, st);
        if (ce != cudaSuccess) return ce;
      }
      return launch_forward_flat(value, shapes, lsi, src, outf, d, value_dtype, sm_count,
                                 fold ? clear : nullptr, fold ? clear_bytes : 0, st);
    }
    if (rows_fold_clear(clear, clear_bytes)) {
      g_fwd_clear = clear;
      g_fwd_clear_bytes = clear_bytes;
    } else {
      const cudaError_t ce = clear_by_memset(clear, clear_bytes, st);
      if (ce != cudaSuccess) return ce;
    }
#define MSDA_ROWS_CASE(DD)                                                                   \
  case DD:                                                                                   \
    if (value_dtype == MSDA_F32) {                                                           \
      const int split = choose_split(d, DD / 4, sm_count);                                   \
      return launch_rows<DD, float, PlainSource>(value, shapes, lsi, src, outf, d, split, st); \
    } else {                                                                                 \
      const int split = choose_split(d, DD / 8, sm_count);                                   \
      return launch_rows<DD, __nv_bfloat16, PlainSource>(value, shapes, lsi, src, outf, d, split, st); \
    }
    switch (d.D) {
      MSDA_ROWS_CASE(16)
      MSDA_ROWS_CASE(32)
      MSDA_ROWS_CASE(64)
      default: break;
    }
#undef MSDA_ROWS_CASE
  }
  // generic path
  {
    const cudaError_t ce = clear_by_memset(clear, clear_bytes, st);
    if (ce != cudaSuccess) return ce;
  }
  const int64_t total = static_cast<int64_t>(d.B) * d.Q * d.M * d.D;
  const int64_t want = (total + 255) / 256;
  const unsigned blocks = static_cast<unsigned>(want < (1 << 20) ? want : (1 << 20));
  if (dtype == MSDA_F32 && value_dtype == MSDA_F32) {
    msda_fwd_generic_kernel<float, float><<<blocks, 256, 0, st>>>(
        static_cast<const float*>(value), shapes, lsi, static_cast<const float*>(loc),
        static_cast<const float*>(aw), static_cast<float*>(out), d);
  } else if (dtype == MSDA_F32 && value_dtype == MSDA_BF16) {
    msda_fwd_generic_kernel<float, __nv_bfloat16><<<blocks, 256, 0, st>>>(
        static_cast<const __nv_bfloat16*>(value), shapes, lsi, static_cast<const float*>(loc),
        static_cast<const float*>(aw), static_cast<float*>(out), d);
  } else if (dtype == MSDA_F64 && value_dtype == MSDA_F64) {
    msda_fwd_generic_kernel<double, double><<<blocks, 256, 0, st>>>(
        static_cast<const double*>(value), shapes, lsi, static_cast<const double*>(loc),
        static_cast<const double*>(aw), static_cast<double*>(out), d);
  } else {
    return cudaErrorInvalidValue;
  }
  note_launches(1);
  note_kernel(KF_FWD_GENERIC);
  return cudaGetLastError();
}

// fused prologue: D in {16, 32, 64}, fp32 or bf16 value
cudaError_t launch_forward_fused(const void* value, const int64_t* shapes, const int64_t* lsi,
                                 const FusedSource& src, float* out, const Dims& d, int value_dtype,
                                 int sm_count, void* clear, size_t clear_bytes, cudaStream_t st) {
  g_fwd_sm_count = sm_count;
  if (!rows_supported(d.D, value_dtype) || d.L > kMaxSmemLevels) return cudaErrorNotSupported;
  const int vec = value_dtype == MSDA_F32 ? 4 : 8;
  if (flat_preferred(d, d.D / vec, sm_count)) {
    const bool fold = clear && clear_bytes % 16 == 0 && (reinterpret_cast<uintptr_t>(clear) & 15u) == 0;
    if (!fold) {
      const cudaError_t ce = clear_by_memset(clear, clear_bytes, st);
      if (ce != cudaSuccess) return ce;
    }
    return launch_forward_flat_fused(value, shapes, lsi, src, out, d, value_dtype, sm_count,
                                     fold ? clear : nullptr, fold ? clear_bytes : 0, st);
  }
  const cudaError_t ce = clear_by_memset(clear, clear_bytes, st);
  if (ce != cudaSuccess) return ce;
  const int split = choose_split(d, d.D / vec, sm_count);
#define MSDA_FUSED_CASE(DD)                                                                        \
  case DD:                                                                                         \
    return value_dtype == MSDA_F32                                                                 \
               ? launch_rows<DD, float, FusedSource>(value, shapes, lsi, src, out, d, split, st)   \
               : launch_rows<DD, __nv_bfloat16, FusedSource>(value, shapes, lsi, src, out, d, split, st);
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
