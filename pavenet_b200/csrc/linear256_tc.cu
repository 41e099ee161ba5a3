// linear256_tc.cu — the fp32 Linear layers around the sampling op on Blackwell's 5th
// generation tensor cores, with fp32-level accuracy: value_proj, output_proj,
// sampling_offsets, attention_weights (widths 128 / 256), the 256 <-> 1024 feed-forward
// pair of the transformer layers, and the input-, weight- and bias-gradients of each.
//
//   Y[rows, out] = epilogue(X[rows, in] * W^T)        W is (out, in), row-major
//
// after the sampling kernels, these fp32 GEMMs are
// the largest cost of an attention-module call and of a transformer layer (cuBLAS runs
// them as SIMT sgemm because PyTorch keeps TF32 off by default, and so does the
// reference).  Plain TF32 would break the 1e-4 parity contract, so the product is
// computed as a 3xTF32 split:  x = x_hi + x_lo, w = w_hi + w_lo (hi = top 19 bits),
//   x*w ~= x_hi*w_hi + x_hi*w_lo + x_lo*w_hi        (error ~2^-21 per product)
// accumulated in fp32 in tensor memory.
//
// Structure (one CTA = one 128-row x NT-column tile, NT = min(out, 256)):
//   TMA (cp.async.bulk.tensor, 64- or 128-byte swizzle) streams X and the pre-split W in
//   K-chunks of 16 or 32 floats through a 2-stage shared-memory ring; the threads
//   split the X chunk into hi / lo in place; one thread issues
//   tcgen05.mma.kind::tf32 (M=128, N=NT, K=8) x 3 per K-step into a 128 x NT fp32
//   accumulator in TMEM; tcgen05.commit -> mbarrier releases the stage; the
//   epilogue reads TMEM with tcgen05.ld and applies, in this order, bias, ReLU, a
//   gate (the backward of ReLU + dropout), dropout, the padding mask and a residual,
//   then stores fp32 or bf16 rows — for value_proj the (B, S, M, D) layout the
//   sampling kernels read, so no further pass touches the projected value.
#include <atomic>

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "msda_kernels.h"

namespace msda {
namespace {

constexpr int kBM = 128;           // rows per CTA
constexpr int kStages = 2;
constexpr int kUmmaK = 8;          // tf32 MMA K
// in features KIN are 128, 256 or 1024; out features are cut into column tiles of
// NT = 128 or 256 (template parameters below).

// BK = floats per K chunk = one swizzle span: 32 (128-byte swizzle, 192 KiB of shared
// memory, one CTA per SM) or 16 (64-byte swizzle, 96 KiB, two CTAs per SM so one tile's
// epilogue overlaps the other's main loop).
template <int BK, int KIN, int NT>
struct Cfg {
  static constexpr int kBK = BK;
  static constexpr int kChunks = KIN / BK;
  static constexpr uint32_t kABytes = kBM * BK * 4;
  static constexpr uint32_t kBBytes = NT * BK * 4;
  static constexpr uint32_t kStageBytes = 2 * kABytes + 2 * kBBytes;   // A_hi, A_lo, B_hi, B_lo
  static constexpr uint32_t kTxBytes = kABytes + 2 * kBBytes;          // what TMA delivers per stage
  static constexpr uint32_t kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 128 /*barriers*/;
  static constexpr uint32_t kSwizzleBytes = BK * 4;                    // 128 or 64
  static constexpr uint64_t kLayoutType = BK == 32 ? 2 : 4;            // SWIZZLE_128B / SWIZZLE_64B
  static constexpr uint64_t kSBO = 8 * kSwizzleBytes;                  // bytes between 8-row atoms
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(bar), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar,
                                            int c_inner, int c_outer) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c_inner), "r"(c_outer)
      : "memory");
}
// K-major swizzled operand: 8-row atoms (8 x swizzle span bytes), stacked densely.
template <class C>
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
  return static_cast<uint64_t>((smem_addr >> 4) & 0x3fff) | (uint64_t(1) << 16) /*LBO (unused)*/ |
         ((C::kSBO >> 4) << 32) /*SBO*/ | (uint64_t(1) << 46) /*sm100 descriptor*/ |
         (C::kLayoutType << 61);
}
// instruction descriptor: D fp32, A/B tf32, M = 128, N = n; mn_major sets both operands MN-major
__host__ __device__ constexpr uint32_t idesc_tf32(int n, bool mn_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(n >> 3) << 17) |
         (static_cast<uint32_t>(kBM >> 4) << 24) | (mn_major ? ((1u << 15) | (1u << 16)) : 0u);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

// W -> (W_hi, W_lo), once per call (at most 65 536 elements)
__global__ void split_weight_kernel(const float* __restrict__ w, float* __restrict__ w_hi,
                                    float* __restrict__ w_lo, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const float x = w[i], hi = tf32_hi(x);
    w_hi[i] = hi;
    w_lo[i] = x - hi;
  }
}

// Counter-based dropout: element (row, col) of a (rows, n_total) matrix is kept iff a
// 32-bit mix of its linear index and the call's seed is >= threshold = p * 2^32.  The
// backward regenerates the same decision from the seed (dropout_backward kernel below).
__device__ __forceinline__ bool dropout_keep(uint32_t idx, uint32_t seed_lo, uint32_t seed_hi,
                                             uint32_t threshold) {
  uint32_t x = (idx ^ seed_lo) * 0x9E3779B1u;
  x ^= x >> 16;
  x = (x + seed_hi) * 0x85EBCA6Bu;
  x ^= x >> 13;
  x *= 0xC2B2AE35u;
  x ^= x >> 16;
  return x >= threshold;
}

// seed = host part + optional device part
__device__ __forceinline__ void resolve_seed(uint32_t lo, uint32_t hi, const unsigned long long* ptr,
                                             uint32_t* out_lo, uint32_t* out_hi) {
  unsigned long long s = (static_cast<unsigned long long>(hi) << 32) | lo;
  if (ptr != nullptr) s += *ptr;
  *out_lo = static_cast<uint32_t>(s);
  *out_hi = static_cast<uint32_t>(s >> 32);
}

template <typename OT>
__device__ __forceinline__ void store_chunk(OT* row_ptr, const float (&v)[32]);
template <>
__device__ __forceinline__ void store_chunk<__nv_bfloat16>(__nv_bfloat16* p, const float (&v)[32]) {
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    uint4 u;
    __nv_bfloat162 t;
    t = __floats2bfloat162_rn(v[j], v[j + 1]);     u.x = *reinterpret_cast<uint32_t*>(&t);
    t = __floats2bfloat162_rn(v[j + 2], v[j + 3]); u.y = *reinterpret_cast<uint32_t*>(&t);
    t = __floats2bfloat162_rn(v[j + 4], v[j + 5]); u.z = *reinterpret_cast<uint32_t*>(&t);
    t = __floats2bfloat162_rn(v[j + 6], v[j + 7]); u.w = *reinterpret_cast<uint32_t*>(&t);
    *reinterpret_cast<uint4*>(p + j) = u;
  }
}

// One thread = one output row: reads its NT accumulator columns from tensor memory 32 at a
// time (taddr = the row's lane quarter + the accumulator's first column) and applies the epilogue.
//
// fp32 outputs leave through TMA: a thread writing its own 1 KB row with 16-byte stores touches 32
// different lines per warp instruction (32 L1 wavefronts each; measured: the store phase cost as
// much as the MMAs).  Instead each warp parks its 32 x 32 chunk in a 4 KB shared-memory tile laid
// out in the 128-byte swizzle (16-byte unit j of row r at position j ^ (r & 7): conflict-free for
// row-per-lane writes, and exactly what a SWIZZLE_128B tensor map expects) and lane 0 issues one
// cp.async.bulk.tensor store per chunk, double-buffered; rows past the end are clipped by the map.
template <typename OT, int NT>
__device__ __forceinline__ void epilogue_row(uint32_t taddr_row, int r, int rows, int n0, int n_total,
                                             const LinearEpilogue& ep, OT* __restrict__ y,
                                             const CUtensorMap* map_y, uint8_t* warp_tiles, int warp_row0) {
  constexpr bool kTmaStore = sizeof(OT) == 4;
  const int lane = threadIdx.x & 31;
  const bool in_range = r < rows;
  const bool masked = in_range && ep.row_mask != nullptr && ep.mask_mode != 0 && ep.row_mask[r] != 0;
  const int64_t row_off = static_cast<int64_t>(in_range ? r : 0) * n_total + n0;
  OT* yrow = y + row_off;
  const float* gate = ep.gate ? ep.gate + row_off : nullptr;
  const float* res = ep.residual ? ep.residual + row_off : nullptr;
  const uint32_t drop_idx0 = static_cast<uint32_t>(row_off);
  uint32_t seed_lo = 0, seed_hi = 0;
  if (ep.dropout_threshold != 0u) resolve_seed(ep.seed_lo, ep.seed_hi, ep.seed_ptr, &seed_lo, &seed_hi);
#pragma unroll 1
  for (int c0 = 0; c0 < NT; c0 += 32) {
    uint32_t u[32];
    const uint32_t taddr = taddr_row + c0;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, "
        "%14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
          "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]),
          "=r"(u[15]), "=r"(u[16]), "=r"(u[17]), "=r"(u[18]), "=r"(u[19]), "=r"(u[20]), "=r"(u[21]),
          "=r"(u[22]), "=r"(u[23]), "=r"(u[24]), "=r"(u[25]), "=r"(u[26]), "=r"(u[27]), "=r"(u[28]),
          "=r"(u[29]), "=r"(u[30]), "=r"(u[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float b = ep.bias ? __ldg(ep.bias + n0 + c0 + j) : 0.f;
      float acc = __uint_as_float(u[j]) + b;
      if (masked) acc = (ep.mask_mode == 1) ? 0.f : b;
      v[j] = acc;
    }
    if (ep.relu) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
    }
    if (gate != nullptr && in_range) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 g = __ldcs(reinterpret_cast<const float4*>(gate + c0 + j));
        v[j] = g.x > 0.f ? v[j] * ep.gate_scale : 0.f;
        v[j + 1] = g.y > 0.f ? v[j + 1] * ep.gate_scale : 0.f;
        v[j + 2] = g.z > 0.f ? v[j + 2] * ep.gate_scale : 0.f;
        v[j + 3] = g.w > 0.f ? v[j + 3] * ep.gate_scale : 0.f;
      }
    }
    if (ep.dropout_threshold != 0u) {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        v[j] = dropout_keep(drop_idx0 + c0 + j, seed_lo, seed_hi, ep.dropout_threshold)
                   ? v[j] * ep.dropout_scale : 0.f;
    }
    if (res != nullptr && in_range) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 g = __ldcs(reinterpret_cast<const float4*>(res + c0 + j));
        v[j] += g.x; v[j + 1] += g.y; v[j + 2] += g.z; v[j + 3] += g.w;
      }
    }
    if constexpr (kTmaStore) {
      const int buf = (c0 >> 5) & 1;
      if (c0 >= 64) {                      // the store issued two chunks ago must have read this tile
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        __syncwarp();
      }
      float4* tile = reinterpret_cast<float4*>(warp_tiles + buf * 4096);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        tile[lane * 8 + (j ^ (lane & 7))] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                     ::"l"(map_y), "r"(smem_u32(tile)), "r"(n0 + c0), "r"(warp_row0)
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    } else {
      if (in_range) store_chunk<OT>(yrow + c0, v);
    }
  }
  if constexpr (kTmaStore) {
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // before the tiles go away
    __syncwarp();
  }
}

// mask_mode: 0 none; 1 masked rows are written as zeros (mask applied after the
// projection, multi_scale_deform_attn.py:369-371); 2 masked rows are written as
// the bias (mask applied to the input before it, transformer.py:1706-1711).
// (LinearEpilogue is declared in msda_kernels.h.)
constexpr int kThreads = 160;   // warps 0-3: split, MMA issue (thread 0), epilogue; warp 4: TMA producer

template <typename OT, int BK, int KIN, int NT>
__global__ void __launch_bounds__(kThreads, BK == 32 ? 1 : 2)
linear256_tf32x3_kernel(const __grid_constant__ CUtensorMap map_x,
                        const __grid_constant__ CUtensorMap map_whi,
                        const __grid_constant__ CUtensorMap map_wlo,
                        const __grid_constant__ CUtensorMap map_y, const LinearEpilogue ep,
                        OT* __restrict__ y, int rows, int n_total) {
  using C = Cfg<BK, KIN, NT>;
  constexpr int kBK = C::kBK, kChunks = C::kChunks;
  constexpr uint32_t kIdesc = idesc_tf32(NT, false);
  constexpr uint32_t kABytes = C::kABytes, kBBytes = C::kBBytes, kStageBytes = C::kStageBytes,
                     kTxBytes = C::kTxBytes;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;          // SWIZZLE_128B wants 1024-byte alignment
  uint8_t* base_ptr = smem_raw + (base - raw);
  const uint32_t bars = base + kStages * kStageBytes;    // full[2], mma_done[2], tmem slot
  const uint32_t full0 = bars, done0 = bars + 16, tmem_slot = bars + 32;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(base_ptr + kStages * kStageBytes + 32);

  const int tid = threadIdx.x, warp = tid >> 5;
  // column tiles of one row tile are neighbours in the grid, so they share X through L2
  const int n_tiles = n_total / NT;
  const int row0 = (blockIdx.x / n_tiles) * kBM;
  const int n0 = (blockIdx.x % n_tiles) * NT;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                 "n"(NT));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(done0 + 8 * s, 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_whi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_wlo) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_y) : "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = *tmem_slot_ptr;

  auto stage_addr = [&](int s) { return base + s * kStageBytes; };
  auto issue_loads = [&](int chunk, int s) {
    const uint32_t bar = full0 + 8 * s, a = stage_addr(s);
    mbar_expect_tx(bar, kTxBytes);
    tma_load_2d(a, &map_x, bar, chunk * kBK, row0);                        // A (hi, split in place)
    tma_load_2d(a + 2 * kABytes, &map_whi, bar, chunk * kBK, n0);           // B_hi
    tma_load_2d(a + 2 * kABytes + kBBytes, &map_wlo, bar, chunk * kBK, n0); // B_lo
  };
  if (warp == 4) {
    // ---- TMA producer: keeps the ring full; a stage is refilled as soon as the MMAs that
    // read it have retired (tcgen05.commit -> mma_done) ----
    if (tid == 128) {
      for (int c = 0; c < kChunks; ++c) {
        const int s = c & 1;
        if (c >= kStages) mbar_wait(done0 + 8 * s, ((c - kStages) >> 1) & 1);
        issue_loads(c, s);
      }
    }
  } else {
    for (int kc = 0; kc < kChunks; ++kc) {
      const int s = kc & 1;
      const uint32_t parity = (kc >> 1) & 1;
      mbar_wait(full0 + 8 * s, parity);
      // split the X chunk: hi stays where TMA put it, lo goes to the twin buffer
      // (same offsets, so the swizzle pattern is preserved)
      {
        float4* a_hi = reinterpret_cast<float4*>(base_ptr + s * kStageBytes);
        float4* a_lo = reinterpret_cast<float4*>(base_ptr + s * kStageBytes + kABytes);
#pragma unroll
        for (int i = tid; i < static_cast<int>(kABytes / 16); i += 128) {
          const float4 x = a_hi[i];
          const float4 h = make_float4(tf32_hi(x.x), tf32_hi(x.y), tf32_hi(x.z), tf32_hi(x.w));
          a_hi[i] = h;
          a_lo[i] = make_float4(x.x - h.x, x.y - h.y, x.z - h.z, x.w - h.w);
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> tensor-core reads
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      asm volatile("bar.sync 1, 128;" ::: "memory");                 // the 4 consumer warps only
      if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a = stage_addr(s);
#pragma unroll
        for (int k = 0; k < kBK / kUmmaK; ++k) {
          const uint32_t koff = k * kUmmaK * 4;   // bytes inside the swizzle span
          const uint64_t a_hi = umma_desc<C>(a + koff), a_lo = umma_desc<C>(a + kABytes + koff);
          const uint64_t b_hi = umma_desc<C>(a + 2 * kABytes + koff);
          const uint64_t b_lo = umma_desc<C>(a + 2 * kABytes + kBBytes + koff);
          umma_tf32(tmem_d, a_lo, b_hi, kIdesc, (kc | k) != 0);   // small terms first
          umma_tf32(tmem_d, a_hi, b_lo, kIdesc, 1);
          umma_tf32(tmem_d, a_hi, b_hi, kIdesc, 1);
        }
        umma_commit(done0 + 8 * s);   // arrives when the MMAs above have read the stage
      }
    }
    // the last commit covers every MMA issued before it
    mbar_wait(done0 + 8 * ((kChunks - 1) & 1), ((kChunks - 1) >> 1) & 1);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }

  if (warp < 4) {
    // epilogue: warp w owns TMEM lanes 32w .. 32w+31 = rows of the tile
    // (the pipeline stages are free by now: every load was consumed and every MMA has retired)
    epilogue_row<OT, NT>(tmem_d + (static_cast<uint32_t>(warp * 32) << 16), row0 + tid, rows, n0, n_total, ep, y,
                         &map_y, base_ptr + warp * 8192, row0 + warp * 32);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(NT));
  }
}

// ---------------------------------------------------------------------------
// The same product with 256 rows per CTA (tuning knob, off by default).
//
// The 128-row kernel above streams the split weight (hi + lo: 2 * NT * KIN * 4 bytes, 512 KB at
// 256 x 256) from L2 into shared memory once per 128-row tile; at 66 669 rows that is 333 MB of
// L2 -> SM traffic for 68 MB of input, and the kernel runs at the rate the L2 delivers it
// (measured: 5.6 TB/s, 0.072 ms before the TMA-store epilogue).  Here one CTA owns 256 rows: two M = 128 accumulators side
// by side in tensor memory (2 * NT columns, all 512 at NT = 256), so every weight chunk feeds
// twice the MMAs -- 200 MB for the same problem.  One CTA per SM (192 KB: three 64 KB stages of
// A_hi, A_lo (256 x 16 floats each), B_hi, B_lo), 8 consumer warps (split, epilogue: warps
// 0-3 rows 0..127, warps 4-7 rows 128..255), thread 0 issues, warp 8 is the TMA producer.
constexpr int kBM2 = 256, kStages2 = 3, kBK2 = 16, kConsumers2 = 256, kThreads2 = kConsumers2 + 32;

template <int NT>
struct Cfg2 {
  static constexpr uint32_t kABytes = kBM2 * kBK2 * 4;                 // 16 KB (hi or lo)
  static constexpr uint32_t kBBytes = NT * kBK2 * 4;
  static constexpr uint32_t kStageBytes = 2 * kABytes + 2 * kBBytes;
  static constexpr uint32_t kTxBytes = kABytes + 2 * kBBytes;
  static constexpr uint32_t kSmemBytes = kStages2 * kStageBytes + 1024 + 128;
};

template <typename OT, int KIN, int NT>
__global__ void __launch_bounds__(kThreads2, 1)
linear_m256_tf32x3_kernel(const __grid_constant__ CUtensorMap map_x,
                          const __grid_constant__ CUtensorMap map_whi,
                          const __grid_constant__ CUtensorMap map_wlo,
                          const __grid_constant__ CUtensorMap map_y, const LinearEpilogue ep,
                          OT* __restrict__ y, int rows, int n_total) {
  using C = Cfg<kBK2, KIN, NT>;          // swizzle / descriptor constants of the 16-float chunk
  using C2 = Cfg2<NT>;
  constexpr int kChunks = KIN / kBK2;
  constexpr uint32_t kIdesc = idesc_tf32(NT, false);
  constexpr uint32_t kABytes = C2::kABytes, kBBytes = C2::kBBytes, kStageBytes = C2::kStageBytes;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw);
  const uint32_t bars = base + kStages2 * kStageBytes;   // full[3], mma_done[3], tmem slot
  const uint32_t full0 = bars, done0 = bars + 32, tmem_slot = bars + 64;
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(base_ptr + kStages2 * kStageBytes + 64);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int n_tiles = n_total / NT;
  const int row0 = (blockIdx.x / n_tiles) * kBM2;
  const int n0 = (blockIdx.x % n_tiles) * NT;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                 "n"(2 * NT));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    for (int s = 0; s < kStages2; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(done0 + 8 * s, 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_whi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_wlo) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_y) : "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = *tmem_slot_ptr;

  if (warp == kConsumers2 / 32) {
    if (tid == kConsumers2) {
      for (int c = 0; c < kChunks; ++c) {
        const int s = c % kStages2;
        if (c >= kStages2) mbar_wait(done0 + 8 * s, ((c / kStages2) - 1) & 1);
        const uint32_t bar = full0 + 8 * s, a = base + s * kStageBytes;
        mbar_expect_tx(bar, C2::kTxBytes);
        tma_load_2d(a, &map_x, bar, c * kBK2, row0);                           // A: 256 rows (hi, split in place)
        tma_load_2d(a + 2 * kABytes, &map_whi, bar, c * kBK2, n0);             // B_hi
        tma_load_2d(a + 2 * kABytes + kBBytes, &map_wlo, bar, c * kBK2, n0);   // B_lo
      }
    }
  } else {
    for (int kc = 0; kc < kChunks; ++kc) {
      const int s = kc % kStages2;
      mbar_wait(full0 + 8 * s, (kc / kStages2) & 1);
      {
        float4* a_hi = reinterpret_cast<float4*>(base_ptr + s * kStageBytes);
        float4* a_lo = reinterpret_cast<float4*>(base_ptr + s * kStageBytes + kABytes);
#pragma unroll
        for (int i = tid; i < static_cast<int>(kABytes / 16); i += kConsumers2) {
          const float4 x = a_hi[i];
          const float4 h = make_float4(tf32_hi(x.x), tf32_hi(x.y), tf32_hi(x.z), tf32_hi(x.w));
          a_hi[i] = h;
          a_lo[i] = make_float4(x.x - h.x, x.y - h.y, x.z - h.z, x.w - h.w);
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a = base + s * kStageBytes;
#pragma unroll
        for (int k = 0; k < kBK2 / kUmmaK; ++k) {
          const uint32_t koff = k * kUmmaK * 4;
          const uint64_t b_hi = umma_desc<C>(a + 2 * kABytes + koff);
          const uint64_t b_lo = umma_desc<C>(a + 2 * kABytes + kBBytes + koff);
#pragma unroll
          for (int h = 0; h < 2; ++h) {                       // rows 128h .. 128h+127 of the tile
            const uint32_t aoff = h * (kABytes / 2) + koff;
            const uint64_t a_hi = umma_desc<C>(a + aoff), a_lo = umma_desc<C>(a + kABytes + aoff);
            const uint32_t d = tmem_d + h * NT;
            umma_tf32(d, a_lo, b_hi, kIdesc, (kc | k) != 0);
            umma_tf32(d, a_hi, b_lo, kIdesc, 1);
            umma_tf32(d, a_hi, b_hi, kIdesc, 1);
          }
        }
        umma_commit(done0 + 8 * s);
      }
    }
    constexpr int last = kChunks - 1;
    mbar_wait(done0 + 8 * (last % kStages2), (last / kStages2) & 1);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int h = warp >> 2, q = warp & 3;                    // accumulator, TMEM lane quarter
    epilogue_row<OT, NT>(tmem_d + (static_cast<uint32_t>(q * 32) << 16) + h * NT,
                         row0 + h * 128 + q * 32 + (tid & 31), rows, n0, n_total, ep, y,
                         &map_y, base_ptr + warp * 8192, row0 + h * 128 + q * 32);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(2 * NT));
  }
}

// ---------------------------------------------------------------------------
// weight gradient:  dW[o][i] = sum_r dY[r][o] * X[r][i]      (256 x 256, K = rows)
//
// Split-K over the rows: each CTA owns a contiguous slice of rows, accumulates a
// full 256 x 256 fp32 partial in TMEM (two M=128 halves x N=256 = all 512
// columns), and reduces it into dW with red.global.add.v4.f32.  Both operands
// are activations laid out (rows, 256), i.e. MN-major for this product: TMA
// fetches 16-row slabs as eight 32-float-wide column blocks (one 3-D box), which
// is the canonical MN-major operand layout for 32-bit types (128-byte rows swizzled in
// 32-byte units, 4 k-rows per atom, atoms along M/N at LBO, along K at SBO); both
// slabs are split into hi / lo in place for the 3xTF32 product.
// ---------------------------------------------------------------------------
constexpr int kWgRows = 16;                                   // rows (K) per stage
constexpr int kWgStages = 3;
constexpr uint32_t kWgAtomStride = kWgRows * 128;             // bytes between 32-float column blocks

template <int MOUT, int NIN>
struct WgCfg {
  static constexpr uint32_t kSlabA = kWgRows * MOUT * 4;      // dY slab, hi or lo
  static constexpr uint32_t kSlabB = kWgRows * NIN * 4;       // X slab, hi or lo
  static constexpr uint32_t kStageBytes = 2 * kSlabA + 2 * kSlabB;   // dY_hi, dY_lo, X_hi, X_lo
  static constexpr uint32_t kSmemBytes = kWgStages * kStageBytes + 1024 + 128;
  static constexpr int kHalves = MOUT / 128;                  // M = 128 accumulators
  static constexpr int kTmemCols = kHalves * NIN;             // 256 or 512
};

// 32-bit MN-major operands have exactly one legal shared-memory layout: 128-byte rows
// swizzled in 32-byte units, in atoms of 4 k-rows (CUTLASS: "for mn-major tf32 operands,
// SW128_32B is the only available smem layout"); TMA produces it with
// CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.  LBO: next 32-float column block; SBO: next 4 k-rows.
__device__ __forceinline__ uint64_t umma_desc_mn(uint32_t smem_addr) {
  return static_cast<uint64_t>((smem_addr >> 4) & 0x3fff) | (uint64_t(kWgAtomStride >> 4) << 16) /*LBO*/ |
         (uint64_t(512 >> 4) << 32) /*SBO*/ | (uint64_t(1) << 46) | (uint64_t(1) << 61) /*SWIZZLE_128B_BASE32B*/;
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0,
                                            int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// split one 16-row slab into hi / lo in place; rows with row_mask != 0 are dropped
__device__ __forceinline__ void split_slab(uint8_t* hi_base, uint32_t slab_bytes, int tid, int row0, int rows,
                                           const uint8_t* row_mask) {
  float4* hi = reinterpret_cast<float4*>(hi_base);
  float4* lo = reinterpret_cast<float4*>(hi_base + slab_bytes);
#pragma unroll 4
  for (int j = tid; j < static_cast<int>(slab_bytes / 16); j += 128) {
    float4 x = hi[j];
    if (row_mask != nullptr) {
      const int r = row0 + ((j >> 3) & (kWgRows - 1));             // 8 float4 per 128-byte row
      if (r < rows && row_mask[r]) x = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const float4 h = make_float4(tf32_hi(x.x), tf32_hi(x.y), tf32_hi(x.z), tf32_hi(x.w));
    hi[j] = h;
    lo[j] = make_float4(x.x - h.x, x.y - h.y, x.z - h.z, x.w - h.w);
  }
}

// zero_dy / zero_x: rows with row_mask != 0 contribute nothing through dY (mask applied
// after the projection) or through X (mask applied to the input before it)
template <int MOUT, int NIN>
__global__ void __launch_bounds__(kThreads, 1)
linear256_wgrad_kernel(const __grid_constant__ CUtensorMap map_dy, const __grid_constant__ CUtensorMap map_x,
                       const uint8_t* __restrict__ row_mask, int zero_dy, int zero_x,
                       float* __restrict__ dw, int rows, int rows_per_cta, int n_tiles, int in_total) {
  // MOUT x NIN is one tile of dW; blockIdx.y walks the tiles of a wider weight
  // (column blocks of dY select the tile's output rows, column blocks of X its inputs)
  using W = WgCfg<MOUT, NIN>;
  const int m_t = blockIdx.y / n_tiles, n_t = blockIdx.y % n_tiles;
  dw += static_cast<int64_t>(m_t) * MOUT * in_total + n_t * NIN;
  constexpr uint32_t kSlabA = W::kSlabA, kSlabB = W::kSlabB, kStageBytes = W::kStageBytes;
  constexpr uint32_t kIdescMN = idesc_tf32(NIN, true);
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw);
  const uint32_t bars = base + kWgStages * kStageBytes;
  const uint32_t full0 = bars, done0 = bars + 32, tmem_slot = bars + 64;
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(base_ptr + kWgStages * kStageBytes + 64);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int row_begin = blockIdx.x * rows_per_cta;
  const int row_end = min(rows, row_begin + rows_per_cta);
  const int n_stages = (row_end - row_begin + kWgRows - 1) / kWgRows;   // >= 1 by construction

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                 "n"(W::kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    for (int s = 0; s < kWgStages; ++s) {
      mbar_init(full0 + 8 * s, 1);
      mbar_init(done0 + 8 * s, 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_dy) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = *tmem_slot_ptr;

  if (warp == 4) {
    if (tid == 128) {
      for (int c = 0; c < n_stages; ++c) {
        const int s = c % kWgStages;
        if (c >= kWgStages) mbar_wait(done0 + 8 * s, ((c / kWgStages) - 1) & 1);
        const uint32_t bar = full0 + 8 * s, a = base + s * kStageBytes;
        mbar_expect_tx(bar, kSlabA + kSlabB);
        tma_load_3d(a, &map_dy, bar, 0, row_begin + c * kWgRows, m_t * (MOUT / 32));
        tma_load_3d(a + 2 * kSlabA, &map_x, bar, 0, row_begin + c * kWgRows, n_t * (NIN / 32));
      }
    }
  } else {
    for (int c = 0; c < n_stages; ++c) {
      const int s = c % kWgStages;
      mbar_wait(full0 + 8 * s, (c / kWgStages) & 1);
      {
        uint8_t* st = base_ptr + s * kStageBytes;
        const int row0 = row_begin + c * kWgRows;
        split_slab(st, kSlabA, tid, row0, rows, zero_dy ? row_mask : nullptr);
        split_slab(st + 2 * kSlabA, kSlabB, tid, row0, rows, zero_x ? row_mask : nullptr);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a = base + s * kStageBytes;
#pragma unroll
        for (int kg = 0; kg < kWgRows / kUmmaK; ++kg) {                    // 8 rows per MMA
          const uint32_t koff = kg * 1024;
          const uint64_t b_hi = umma_desc_mn(a + 2 * kSlabA + koff);
          const uint64_t b_lo = umma_desc_mn(a + 2 * kSlabA + kSlabB + koff);
#pragma unroll
          for (int h = 0; h < W::kHalves; ++h) {                           // output rows o in [128h, 128h+128)
            const uint32_t aoff = h * 4 * kWgAtomStride + koff;
            const uint64_t a_hi = umma_desc_mn(a + aoff), a_lo = umma_desc_mn(a + kSlabA + aoff);
            const uint32_t d = tmem_d + h * NIN;
            umma_tf32(d, a_lo, b_hi, kIdescMN, (c | kg) != 0);
            umma_tf32(d, a_hi, b_lo, kIdescMN, 1);
            umma_tf32(d, a_hi, b_hi, kIdescMN, 1);
          }
        }
        umma_commit(done0 + 8 * s);
      }
    }
    const int last = n_stages - 1;
    mbar_wait(done0 + 8 * (last % kWgStages), (last / kWgStages) & 1);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // epilogue: lane = output row o within the half, 32 input columns per TMEM load
#pragma unroll 1
    for (int h = 0; h < W::kHalves; ++h) {
      float* drow = dw + static_cast<int64_t>(h * 128 + tid) * in_total;
#pragma unroll 1
      for (int c0 = 0; c0 < NIN; c0 += 32) {
        uint32_t u[32];
        const uint32_t taddr = tmem_d + (static_cast<uint32_t>(warp * 32) << 16) + h * NIN + c0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, "
            "%14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
              "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]),
              "=r"(u[15]), "=r"(u[16]), "=r"(u[17]), "=r"(u[18]), "=r"(u[19]), "=r"(u[20]), "=r"(u[21]),
              "=r"(u[22]), "=r"(u[23]), "=r"(u[24]), "=r"(u[25]), "=r"(u[26]), "=r"(u[27]), "=r"(u[28]),
              "=r"(u[29]), "=r"(u[30]), "=r"(u[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(drow + c0 + j),
                       "f"(__uint_as_float(u[j])), "f"(__uint_as_float(u[j + 1])),
                       "f"(__uint_as_float(u[j + 2])), "f"(__uint_as_float(u[j + 3]))
                       : "memory");
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(W::kTmemCols));
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

// cudaFuncSetAttribute once per kernel and device (later launches may be inside a stream capture,
// where the fewer host-side driver calls the better)
struct SmemOptIn {
  std::atomic<unsigned long long> done{0};   // bit d: set on device d (forward and autograd threads both launch)
  template <typename K>
  cudaError_t ensure(K kernel, uint32_t smem) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 64 && ((done.load(std::memory_order_acquire) >> dev) & 1ull)) return cudaSuccess;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e == cudaSuccess && dev < 64) done.fetch_or(1ull << dev, std::memory_order_release);
    return e;
  }
};

// (rows x width) fp32 row-major matrix, boxes of bk floats x box_rows rows, swizzle span = bk floats
bool make_map(CUtensorMap* map, const float* ptr, int rows, int width, int box_rows, int bk) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  const cuuint64_t dims[2] = {static_cast<cuuint64_t>(width), static_cast<cuuint64_t>(rows)};
  const cuuint64_t strides[1] = {static_cast<cuuint64_t>(width) * 4};
  const cuuint32_t box[2] = {static_cast<cuuint32_t>(bk), static_cast<cuuint32_t>(box_rows)};
  const cuuint32_t estr[2] = {1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, bk == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// the fp32 output (rows x width) for the epilogue's TMA stores: 32 x 32 boxes, 128-byte swizzle
// (bf16 outputs are stored directly; they still get a valid map so the kernel signature is one)
bool make_map_y(CUtensorMap* map, const void* ptr, int rows, int width, bool is_f32) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  const cuuint64_t elt = is_f32 ? 4 : 2;
  const cuuint64_t dims[2] = {static_cast<cuuint64_t>(width), static_cast<cuuint64_t>(rows)};
  const cuuint64_t strides[1] = {static_cast<cuuint64_t>(width) * elt};
  const cuuint32_t box[2] = {32, 32};
  const cuuint32_t estr[2] = {1, 1};
  return fn(map, is_f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
            const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            is_f32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
            CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// (rows x width) fp32 viewed as {32 floats, rows, width/32 column blocks}: one box = a 16-row slab of
// `tile` columns laid out block-major, each block a run of 128-byte rows (MN-major SW128_32B operand)
bool make_map_mn(CUtensorMap* map, const float* ptr, int rows, int width, int tile) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  const cuuint64_t dims[3] = {32, static_cast<cuuint64_t>(rows), static_cast<cuuint64_t>(width / 32)};
  const cuuint64_t strides[2] = {static_cast<cuuint64_t>(width) * 4, 128};
  const cuuint32_t box[3] = {32, static_cast<cuuint32_t>(kWgRows), static_cast<cuuint32_t>(tile / 32)};
  const cuuint32_t estr[3] = {1, 1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(ptr), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <typename OT, int BK, int KIN, int NT>
cudaError_t launch_variant(const float* x, const float* w_hi, const float* w_lo, const LinearEpilogue& ep,
                           void* y, int rows, int n_total, cudaStream_t st) {
  CUtensorMap mx, mhi, mlo, my;
  if (!make_map(&mx, x, rows, KIN, kBM, BK) || !make_map(&mhi, w_hi, n_total, KIN, NT, BK) ||
      !make_map(&mlo, w_lo, n_total, KIN, NT, BK) || !make_map_y(&my, y, rows, n_total, sizeof(OT) == 4))
    return cudaErrorNotSupported;
  constexpr uint32_t smem = Cfg<BK, KIN, NT>::kSmemBytes;
  static SmemOptIn opt_in;
  const cudaError_t e = opt_in.ensure(linear256_tf32x3_kernel<OT, BK, KIN, NT>, smem);
  if (e != cudaSuccess) return e;
  const unsigned grid = static_cast<unsigned>((rows + kBM - 1) / kBM) * (n_total / NT);
  linear256_tf32x3_kernel<OT, BK, KIN, NT><<<grid, kThreads, smem, st>>>(
      mx, mhi, mlo, my, ep, static_cast<OT*>(y), rows, n_total);
  return cudaGetLastError();
}

template <typename OT, int KIN, int NT>
cudaError_t launch_m256(const float* x, const float* w_hi, const float* w_lo, const LinearEpilogue& ep,
                        void* y, int rows, int n_total, cudaStream_t st) {
  CUtensorMap mx, mhi, mlo, my;
  if (!make_map(&mx, x, rows, KIN, kBM2, kBK2) || !make_map(&mhi, w_hi, n_total, KIN, NT, kBK2) ||
      !make_map(&mlo, w_lo, n_total, KIN, NT, kBK2) || !make_map_y(&my, y, rows, n_total, sizeof(OT) == 4))
    return cudaErrorNotSupported;
  constexpr uint32_t smem = Cfg2<NT>::kSmemBytes;
  static SmemOptIn opt_in;
  const cudaError_t e = opt_in.ensure(linear_m256_tf32x3_kernel<OT, KIN, NT>, smem);
  if (e != cudaSuccess) return e;
  const unsigned grid = static_cast<unsigned>((rows + kBM2 - 1) / kBM2) * (n_total / NT);
  linear_m256_tf32x3_kernel<OT, KIN, NT><<<grid, kThreads2, smem, st>>>(
      mx, mhi, mlo, my, ep, static_cast<OT*>(y), rows, n_total);
  return cudaGetLastError();
}

template <int KIN, int NT>
cudaError_t launch_shape(const float* x, const float* w_hi, const float* w_lo, const LinearEpilogue& ep,
                         void* y, int rows, int n_total, int out_dtype, cudaStream_t st) {
  // 256 rows per CTA: a tuning knob (PAVENET_MSDA_LINEAR_BM=256).  Measured on B200 it does not beat
  // two resident 128-row CTAs per SM (66 669 x 256 x 256: 0.062 vs 0.058 ms; 256 -> 1024: 0.211 vs 0.189 ms).
  if (tuning().linear_bm == 256) {
    if (out_dtype == MSDA_BF16)
      return launch_m256<__nv_bfloat16, KIN, NT>(x, w_hi, w_lo, ep, y, rows, n_total, st);
    return launch_m256<float, KIN, NT>(x, w_hi, w_lo, ep, y, rows, n_total, st);
  }
  // the 32-float K-chunk variant exists for the 256-wide projections only (tuning knob)
  if (KIN == 256 && tuning().linear_bk == 32) {
    if (out_dtype == MSDA_BF16)
      return launch_variant<__nv_bfloat16, 32, 256, NT>(x, w_hi, w_lo, ep, y, rows, n_total, st);
    return launch_variant<float, 32, 256, NT>(x, w_hi, w_lo, ep, y, rows, n_total, st);
  }
  if (out_dtype == MSDA_BF16)
    return launch_variant<__nv_bfloat16, 16, KIN, NT>(x, w_hi, w_lo, ep, y, rows, n_total, st);
  return launch_variant<float, 16, KIN, NT>(x, w_hi, w_lo, ep, y, rows, n_total, st);
}

template <int MT, int NT>
cudaError_t launch_wgrad_shape(const float* dy, const float* x, const uint8_t* row_mask, int mask_mode,
                               float* dw, int rows, int in_total, int out_total, int sm_count,
                               cudaStream_t st) {
  CUtensorMap mdy, mx;
  if (!make_map_mn(&mdy, dy, rows, out_total, MT) || !make_map_mn(&mx, x, rows, in_total, NT))
    return cudaErrorNotSupported;
  cudaError_t e = cudaMemsetAsync(dw, 0, sizeof(float) * out_total * in_total, st);
  if (e != cudaSuccess) return e;
  constexpr uint32_t smem = WgCfg<MT, NT>::kSmemBytes;
  static SmemOptIn opt_in;
  e = opt_in.ensure(linear256_wgrad_kernel<MT, NT>, smem);
  if (e != cudaSuccess) return e;
  const int n_tiles = in_total / NT, tiles = (out_total / MT) * n_tiles;
  // split-K CTAs per tile: one per SM whatever the tile count -- the tensor core accumulates with
  // truncation, so the error grows with the length of a CTA's chain; short chains keep it at 3e-6
  const int splits = sm_count;
  int rows_per_cta = (rows + splits - 1) / splits;
  rows_per_cta = (rows_per_cta + kWgRows - 1) / kWgRows * kWgRows;
  const dim3 grid((rows + rows_per_cta - 1) / rows_per_cta, tiles);
  linear256_wgrad_kernel<MT, NT><<<grid, kThreads, smem, st>>>(
      mdy, mx, row_mask, mask_mode == 1, mask_mode == 2, dw, rows, rows_per_cta, n_tiles, in_total);
  note_launches(1);
  note_kernel(KF_LINEAR_WGRAD);
  return cudaGetLastError();
}

// bias gradient: out[c] = sum over rows of dy[r][c] (rows with row_mask != 0 skipped).
// torch's reduction over dim 0 of a (rows, 256) tensor runs at ~0.8 TB/s; this streams it once.
// With a dropout threshold the pass is also the backward of the epilogue's dropout: it
// regenerates each keep decision, writes dy_out = dy * keep * scale and sums that instead.
template <int WIDTH>
__global__ void __launch_bounds__(256)
colsum_kernel(const float* __restrict__ dy, const uint8_t* __restrict__ row_mask, float* __restrict__ out,
              float* __restrict__ dy_out, uint32_t threshold, float scale, uint32_t seed_lo_host,
              uint32_t seed_hi_host, const unsigned long long* __restrict__ seed_ptr, int rows,
              int rows_per_block) {
  constexpr int CG = WIDTH / 4;                        // column groups of 4 floats
  constexpr int RL = CG >= 256 ? 1 : 256 / CG;         // row lanes per pass
  constexpr int PASSES = CG > 256 ? CG / 256 : 1;      // column passes (WIDTH 1024: 1)
  static_assert(PASSES == 1, "WIDTH up to 1024");
  __shared__ float4 s_part[RL][CG < 256 ? CG : 256];
  const int cg = threadIdx.x % CG, lane_r = threadIdx.x / CG;
  const int r0 = blockIdx.x * rows_per_block, r1 = min(rows, r0 + rows_per_block);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  uint32_t seed_lo = 0, seed_hi = 0;
  if (threshold != 0u) resolve_seed(seed_lo_host, seed_hi_host, seed_ptr, &seed_lo, &seed_hi);
  for (int r = r0 + lane_r; r < r1; r += RL) {
    if (row_mask != nullptr && row_mask[r]) continue;
    const int64_t off = static_cast<int64_t>(r) * WIDTH + 4 * cg;
    float4 v = __ldcs(reinterpret_cast<const float4*>(dy + off));
    if (threshold != 0u) {
      const uint32_t i = static_cast<uint32_t>(off);
      v.x = dropout_keep(i, seed_lo, seed_hi, threshold) ? v.x * scale : 0.f;
      v.y = dropout_keep(i + 1, seed_lo, seed_hi, threshold) ? v.y * scale : 0.f;
      v.z = dropout_keep(i + 2, seed_lo, seed_hi, threshold) ? v.z * scale : 0.f;
      v.w = dropout_keep(i + 3, seed_lo, seed_hi, threshold) ? v.w * scale : 0.f;
      *reinterpret_cast<float4*>(dy_out + off) = v;
    }
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  if (out == nullptr) return;
  if (RL > 1) {
    s_part[lane_r][cg] = acc;
    __syncthreads();
    if (lane_r == 0) {
      for (int k = 1; k < RL; ++k) {
        const float4 v = s_part[k][cg];
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
    }
  }
  if (lane_r == 0) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(out + 4 * cg), "f"(acc.x), "f"(acc.y),
                 "f"(acc.z), "f"(acc.w)
                 : "memory");
  }
}

}  // namespace

static bool width_ok(int n) { return n == 128 || n == 256 || n == 1024; }

bool linear_shape_supported(int in_features, int out_features) {
  // every pair of {128, 256, 1024} except the square 128 and 1024 (nothing in the model uses them)
  return width_ok(in_features) && width_ok(out_features) &&
         !(in_features == out_features && in_features != 256);
}

// y = epilogue(x w^T); returns cudaSuccess, cudaErrorNotSupported (shape, or no driver entry
// point), or the launch error
cudaError_t launch_linear256(const float* x, const float* w, const LinearEpilogue& ep, void* y, int rows,
                             int in_features, int out_features, int out_dtype, float* scratch,
                             cudaStream_t st) {
  if (!linear_shape_supported(in_features, out_features)) return cudaErrorNotSupported;
  const int n = in_features * out_features;
  float* w_hi = scratch;
  float* w_lo = scratch + n;
  split_weight_kernel<<<(n + 255) / 256, 256, 0, st>>>(w, w_hi, w_lo, n);
  note_launches(2);
  note_kernel(KF_LINEAR);
  const bool wide = out_features != 128;      // column tiles of 256, or one of 128
  switch (in_features) {
    case 128:
      return launch_shape<128, 256>(x, w_hi, w_lo, ep, y, rows, out_features, out_dtype, st);
    case 256:
      return wide ? launch_shape<256, 256>(x, w_hi, w_lo, ep, y, rows, out_features, out_dtype, st)
                  : launch_shape<256, 128>(x, w_hi, w_lo, ep, y, rows, out_features, out_dtype, st);
    default:
      return wide ? launch_shape<1024, 256>(x, w_hi, w_lo, ep, y, rows, out_features, out_dtype, st)
                  : launch_shape<1024, 128>(x, w_hi, w_lo, ep, y, rows, out_features, out_dtype, st);
  }
}

// dW (out x in, fp32) = dY^T X; dW is overwritten
cudaError_t launch_linear256_wgrad(const float* dy, const float* x, const uint8_t* row_mask,
                                   int mask_mode, float* dw, int rows, int in_features, int out_features,
                                   int sm_count, cudaStream_t st) {
  if (!linear_shape_supported(in_features, out_features)) return cudaErrorNotSupported;
  const bool m128 = out_features == 128, n128 = in_features == 128;
  if (m128)
    return launch_wgrad_shape<128, 256>(dy, x, row_mask, mask_mode, dw, rows, in_features, out_features,
                                        sm_count, st);
  if (n128)
    return launch_wgrad_shape<256, 128>(dy, x, row_mask, mask_mode, dw, rows, in_features, out_features,
                                        sm_count, st);
  return launch_wgrad_shape<256, 256>(dy, x, row_mask, mask_mode, dw, rows, in_features, out_features,
                                      sm_count, st);
}

// out[width] = column sums of dy (optionally of dropout-backward(dy), also written to dy_out)
cudaError_t launch_colsum256(const float* dy, const uint8_t* row_mask, float* out, float* dy_out,
                             uint32_t threshold, float scale, uint32_t seed_lo, uint32_t seed_hi,
                             const unsigned long long* seed_ptr, int rows, int width, int sm_count,
                             cudaStream_t st) {
  if (!width_ok(width)) return cudaErrorNotSupported;
  if (out != nullptr) {
    const cudaError_t e = cudaMemsetAsync(out, 0, sizeof(float) * width, st);
    if (e != cudaSuccess) return e;
  }
  const int blocks_wanted = sm_count * 8;
  int rows_per_block = (rows + blocks_wanted - 1) / blocks_wanted;
  rows_per_block = (rows_per_block + 7) / 8 * 8;
  const unsigned grid = (rows + rows_per_block - 1) / rows_per_block;
#define MSDA_COLSUM(W_) \
  colsum_kernel<W_><<<grid, 256, 0, st>>>(dy, row_mask, out, dy_out, threshold, scale, seed_lo, seed_hi, \
                                          seed_ptr, rows, rows_per_block)
  if (width == 128) MSDA_COLSUM(128);
  else if (width == 256) MSDA_COLSUM(256);
  else MSDA_COLSUM(1024);
#undef MSDA_COLSUM
  note_launches(1);
  note_kernel(KF_COLSUM);
  return cudaGetLastError();
}

}  // namespace msda
