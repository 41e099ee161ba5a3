// msda_bwd_priv.cu — backward variant 3 of the large-Q shapes: the coarsest pyramid level of grad_value is
// PRIVATISED in shared memory by a block that owns one (batch entry, head) and a strided set of query chunks.
//
// This is the "(b)" variant the round-1 review asked to be built and measured in the kernel: on the encoder
// (config 2) a quarter of all row reductions — 8.5 M of 34.1 M — land on the 273 pixels x 8 heads x 3 frames of
// the coarsest level, ~1 300 per row.  Here they are accumulated on the SM and leave it once per block and row.
// sm_100a has no floating-point shared-memory atomic (atomicAdd(float) on shared memory is an ATOMS.CAST.SPIN
// loop); a lane adds its 16-byte slice of a row with LDS.128 + 4 FADD + one native 128-bit compare-and-swap
// (ATOMS.CAS.128), retried when another lane got in between.  Rows of the other levels go to global memory as in
// the default kernel (red.global.add.v4.f32).
//
// Maths and record layout are those of msda_bwd_rows_kernel (msda_bwd.cu; reference:
// third_party/mmcv/mmcv/ops/csrc/common/cuda/ms_deform_attn_cuda_kernel.cuh:66-131, 256-345).  Specialised to
// D = 32, fp32 value and gradients, the plain (unfused) interface, rows not split — the benchmark's kernel.
// Selected with msda_set_option("bwd_variant", 3); measured in profiles/r02_kernel_variant_ab.txt.
#include "msda_kernels.h"
#include "msda_bwd_io.cuh"

namespace msda {

namespace {
constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int D = 32, VEC = 4, G = 8, NG = 4;

// tile[addr] += v for one 16-byte slice, atomically with respect to every other lane of the block
__device__ __forceinline__ void smem_add_cas128(uint32_t saddr, const float (&v)[4]) {
  float4 old;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(old.x), "=f"(old.y), "=f"(old.z), "=f"(old.w)
               : "r"(saddr));
  while (true) {
    const unsigned long long c0 = (unsigned long long)__float_as_uint(old.x) | ((unsigned long long)__float_as_uint(old.y) << 32);
    const unsigned long long c1 = (unsigned long long)__float_as_uint(old.z) | ((unsigned long long)__float_as_uint(old.w) << 32);
    const unsigned long long n0 = (unsigned long long)__float_as_uint(old.x + v[0]) | ((unsigned long long)__float_as_uint(old.y + v[1]) << 32);
    const unsigned long long n1 = (unsigned long long)__float_as_uint(old.z + v[2]) | ((unsigned long long)__float_as_uint(old.w + v[3]) << 32);
    unsigned long long o0, o1;
    asm volatile("{\n .reg .b128 c, n, o;\n mov.b128 c, {%3, %4};\n mov.b128 n, {%5, %6};\n"
                 " atom.shared.cas.b128 o, [%2], c, n;\n mov.b128 {%0, %1}, o;\n}\n"
                 : "=l"(o0), "=l"(o1)
                 : "r"(saddr), "l"(c0), "l"(c1), "l"(n0), "l"(n1)
                 : "memory");
    if (o0 == c0 && o1 == c1) break;
    old.x = __uint_as_float((uint32_t)o0); old.y = __uint_as_float((uint32_t)(o0 >> 32));
    old.z = __uint_as_float((uint32_t)o1); old.w = __uint_as_float((uint32_t)(o1 >> 32));
  }
}

extern __shared__ __align__(16) float s_priv[];   // the privatised level: rows x 32 floats

__global__ void __launch_bounds__(kThreads, 3)
msda_bwd_rows_priv_kernel(const float* __restrict__ value, const int64_t* __restrict__ shapes,
                          const int64_t* __restrict__ lsi, PlainIO io, const float* __restrict__ grad_out,
                          float* __restrict__ grad_value, Dims d, int blocks_per_pair, int tile_bytes) {
  __shared__ LevelInfo s_lvl[kMaxSmemLevels];
  __shared__ int4 s_board[kWarps][G * (2 * NG + 1)];

  const int MD = d.M * D;
  for (int l = threadIdx.x; l < d.L; l += blockDim.x) s_lvl[l] = load_level(shapes, lsi, l, MD);
  __syncthreads();

  // the coarsest (last) level is privatised when it fits the tile the launch provides
  const int pl = d.L - 1;
  const int prows = s_lvl[pl].H * s_lvl[pl].W;
  const bool priv = prows * D * static_cast<int>(sizeof(float)) <= tile_bytes;
  const int pstart = s_lvl[pl].start, pW = s_lvl[pl].W;
  if (priv)
    for (int i = threadIdx.x; i < prows * (D / 4); i += blockDim.x)
      reinterpret_cast<float4*>(s_priv)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();
  const uint32_t tile = static_cast<uint32_t>(__cvta_generic_to_shared(s_priv));

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int gl = lane & (G - 1);
  const int grp = lane / G;

  // block -> (batch entry, head, slot); the block walks the query chunks slot, slot + K, slot + 2K, ...
  constexpr int QPB = kThreads / G;
  const int n_chunks = (d.Q + QPB - 1) / QPB;
  int blk = blockIdx.x;
  const int m = blk % d.M;
  blk /= d.M;
  const int slot = blk % blocks_per_pair;
  const int64_t b = blk / blocks_per_pair;
  const int gib = threadIdx.x / G;
  const int64_t boff = b * d.S * MD + m * D + gl * VEC;
  const float* vbase = value + boff;
  float* gvbase = grad_value + boff;
  const int LP = d.L * d.P;
  int4* board = s_board[warp];
  auto unit_of = [](int j, int grp_, int half) { return j * (2 * NG + 1) + 2 * grp_ + half; };
  const FastDivP level_of(d.P);

  for (int chunk = slot; chunk < n_chunks; chunk += blocks_per_pair) {
    int q_idx = chunk * QPB + gib;
    const bool live = q_idx < d.Q;
    if (!live) q_idx = d.Q - 1;
    const int64_t unit = (b * d.Q + q_idx) * d.M + m;
    io.bind(unit, LP, d.M, b * d.Q + q_idx);

    float g[VEC];
    {
      const float4 t = __ldcs(reinterpret_cast<const float4*>(grad_out + unit * D + gl * VEC));
      g[0] = t.x; g[1] = t.y; g[2] = t.z; g[3] = t.w;
    }
    const int per = (LP + G - 1) / G * G;
    const int s_end = live ? LP : 0;

    RawSample nxt;
    nxt.x = nxt.y = nxt.w = 0.f;
    if (gl < s_end) nxt = io.src.load(gl);
    for (int s0 = 0; s0 < per; s0 += G) {
      const int s = s0 + gl;
      float Wf = 0.f, Hf = 0.f, w_true = 0.f;
      int lvl = 0;
      {
        SampleRec r;
        r.off00 = 0; r.meta = 0; r.lh = 0.f; r.lw = 0.f; r.a = 0.f; r.rs = 0;
        int pix = 0;
        if (s < s_end) {
          lvl = level_of(s);
          const LevelInfo lv = s_lvl[lvl];
          RawSample cur = nxt;
          w_true = cur.w;
          r.a = cur.w;
          Wf = static_cast<float>(lv.W);
          Hf = static_cast<float>(lv.H);
          r.rs = lv.row_stride;
          make_sample(cur.x, cur.y, r.a, lv, lvl, MD, r.off00, r.meta, r.lh, r.lw);
          if (priv && lvl == pl) pix = r.off00 / MD - pstart;   // pixel index of corner (h0, w0) inside the level
        }
        board[unit_of(gl, grp, 0)] =
            make_int4(r.off00, r.meta, __float_as_int(r.lh), __float_as_int(r.lw));
        board[unit_of(gl, grp, 1)] = make_int4(__float_as_int(r.a), r.rs, pix, 0);
        const int sn = s + G;
        if (sn < s_end) nxt = io.src.load(sn);
      }
      __syncwarp();

      float pw[G], px[G], py[G];
#pragma unroll
      for (int j = 0; j < G; ++j) {
        const int4 q = board[unit_of(j, grp, 0)];
        const int4 ar = board[unit_of(j, grp, 1)];
        const int meta = q.y;
        const float a = __int_as_float(ar.x);
        const float lh = __int_as_float(q.z), lw = __int_as_float(q.w);
        const float hh = 1.f - lh, hw = 1.f - lw;
        const int rs = ar.y;
        const float* p = vbase + q.x;
        float* gp = gvbase + q.x;
        float v1[VEC], v2[VEC], v3[VEC], v4[VEC];
#pragma unroll
        for (int c = 0; c < VEC; ++c) v1[c] = v2[c] = v3[c] = v4[c] = 0.f;
        if (meta & 1) Vec16<float>::load(p, v1);
        if (meta & 2) Vec16<float>::load(p + MD, v2);
        if (meta & 4) Vec16<float>::load(p + rs, v3);
        if (meta & 8) Vec16<float>::load(p + rs + MD, v4);
        const float w1 = hh * hw, w2 = hh * lw, w3 = lh * hw, w4 = lh * lw;
        float sw = 0.f, sx = 0.f, sy = 0.f;
#pragma unroll
        for (int c = 0; c < VEC; ++c) {
          const float tg = g[c] * a;
          const float val = w1 * v1[c] + w2 * v2[c] + w3 * v3[c] + w4 * v4[c];
          const float gw = hh * (v2[c] - v1[c]) + lh * (v4[c] - v3[c]);
          const float gh = hw * (v3[c] - v1[c]) + lw * (v4[c] - v2[c]);
          sw += g[c] * val;
          sx += gw * tg;
          sy += gh * tg;
        }
        pw[j] = sw; px[j] = sx; py[j] = sy;
        float t[VEC];
        if (priv && (meta >> 4) == pl && (meta & 15)) {
          // this sample's corners live in the privatised tile: accumulate on the SM
          const uint32_t row00 = tile + static_cast<uint32_t>(ar.z) * (D * 4) + gl * 16;
#define MSDA_PRIV(BIT, WK, DPIX)                                              \
  if (meta & BIT) {                                                           \
    _Pragma("unroll") for (int c = 0; c < VEC; ++c) t[c] = (WK) * a * g[c];  \
    smem_add_cas128(row00 + (DPIX) * (D * 4), t);                             \
  }
          MSDA_PRIV(1, w1, 0)
          MSDA_PRIV(2, w2, 1)
          MSDA_PRIV(4, w3, pW)
          MSDA_PRIV(8, w4, pW + 1)
#undef MSDA_PRIV
        } else {
#define MSDA_SCATTER(BIT, WK, PTR)                                            \
  if (meta & BIT) {                                                           \
    _Pragma("unroll") for (int c = 0; c < VEC; ++c) t[c] = (WK) * a * g[c];  \
    red_add_row(PTR, t);                                                      \
  }
          MSDA_SCATTER(1, w1, gp)
          MSDA_SCATTER(2, w2, gp + MD)
          MSDA_SCATTER(4, w3, gp + rs)
          MSDA_SCATTER(8, w4, gp + rs + MD)
#undef MSDA_SCATTER
        }
      }
      const float tw = group_transpose_reduce<G>(pw, gl);
      const float tx = group_transpose_reduce<G>(px, gl);
      const float ty = group_transpose_reduce<G>(py, gl);
      if (s < s_end) io.store(s, lvl, tw, tx, ty, w_true, Wf, Hf);
      __syncwarp();
    }
  }

  // flush the tile: one vector reduction per row that received anything
  if (priv) {
    __syncthreads();
    float* lbase = grad_value + b * d.S * MD + static_cast<int64_t>(pstart) * MD + m * D + gl * VEC;
    for (int row = gib; row < prows; row += QPB) {
      const float4 t4 = reinterpret_cast<const float4*>(s_priv)[row * (D / 4) + gl];
      if (t4.x != 0.f || t4.y != 0.f || t4.z != 0.f || t4.w != 0.f) {
        const float t[4] = {t4.x, t4.y, t4.z, t4.w};
        red_add_row(lbase + static_cast<int64_t>(row) * MD, t);
      }
    }
  }
}
}  // namespace

// launcher: variant 3, fp32 / D = 32 / plain interface only (the caller checks)
cudaError_t launch_backward_priv(const void* value, const int64_t* shapes, const int64_t* lsi,
                                 const PlainIO& io, const float* grad_out, void* grad_value,
                                 const Dims& d, int sm_count, cudaStream_t st) {
  const int tile_kb = tuning().agg_tile_kb;
  const int tile_bytes = tile_kb * 1024;
  static bool attr_set[64] = {};     // the opt-in to > 48 KB of dynamic shared memory is per device
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = -1;
  if (dev < 0 || !attr_set[dev]) {
    const cudaError_t e = cudaFuncSetAttribute(msda_bwd_rows_priv_kernel,
                                               cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return e;
    if (dev >= 0) attr_set[dev] = true;
  }
  // resident blocks per SM that the tile leaves room for (static shared memory ~ 10 KB per block)
  int per_sm = (220 * 1024) / (tile_bytes + 11 * 1024);
  if (per_sm > 3) per_sm = 3;
  if (per_sm < 1) per_sm = 1;
  const int pairs = d.B * d.M;
  int k = (sm_count * per_sm) / pairs;                      // blocks per (batch entry, head): one resident wave, no tail
  const int n_chunks = (d.Q + 31) / 32;
  if (k > n_chunks) k = n_chunks;
  if (k < 1) k = 1;
  const long long blocks = static_cast<long long>(pairs) * k;
  msda_bwd_rows_priv_kernel<<<static_cast<unsigned>(blocks), kThreads, tile_bytes, st>>>(
      static_cast<const float*>(value), shapes, lsi, io, grad_out, static_cast<float*>(grad_value), d, k,
      tile_bytes);
  note_launches(1);
  note_kernel(KF_BWD_ROWS);
  return cudaGetLastError();
}

}  // namespace msda
