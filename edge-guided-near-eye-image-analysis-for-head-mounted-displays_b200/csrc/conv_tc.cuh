// Implicit-GEMM convolution on tcgen05 tensor cores (sm_100a).
//
//   out[n, y, x, :] = epilogue( sum_taps sum_k  A[n, y+dy, x+dx, k] * Wt[tap][:, k] )
//
// * GEMM view: a CTA tile is S (1 or 2) sub-tiles of M = 128 output pixels, each `sr` rows x `bw`
//   columns of one frame (bw = 16 -> 8 rows, bw = 8 -> 16 rows), stacked in y; N = n_tile output
//   channels (<= 256); K = 32-channel chunks x taps.
// * A operand: per (chunk, horizontal tap offset dx) ONE 4-D TMA box {32 ch, bw px, tr + 2*dmax rows,
//   1 frame} is loaded; every vertical tap offset dy of that dx - and both sub-tiles - read it
//   through UMMA descriptors that start (dy + dmax) * bw (+ 128 per sub-tile) rows into the box, so
//   a 3x3 layer fetches 3 boxes per chunk instead of 9 (the L2 -> SM path is what bounds the plain
//   tap-per-load scheme).  TMA out-of-bounds zero fill is the convolution padding for every
//   dilation (1/2/4/8/12); a chunk table (source buffer, channel offset, frame offset) implements
//   channel concatenation without copies.
// * B operand: weight tiles {32 k, n_tile rows} stream through their own ring from the repacked
//   [tap][cout_pad][kpad] matrix; one tile feeds S sub-tiles.  Both operands use the 64-byte
//   swizzled K-major layout.
// * Precision: nsplit==3 issues hi*hi + lo*hi + hi*lo bf16 MMAs into one fp32 TMEM accumulator.
// * Warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2-9 = epilogue
//   (tcgen05.ld -> bias/activation/affine -> split-bf16 NHWC stores [+ InstanceNorm sum/sumsq
//   atomics], or the fused MSBlock tail).  Rings of full/empty mbarriers couple producer and MMA;
//   tmem_full/tmem_empty couple MMA and epilogue (two accumulator buffers when they fit in the
//   512 TMEM columns); persistent CTAs stride over the tile list.
#pragma once
#include <type_traits>

#include "common.cuh"

#define TC_MAX_ALOADS 8

struct TcParams {
  CUtensorMap a_map[2][EGN_MAX_SRC];  // [plane hi/lo][source]
  CUtensorMap w_map[2];               // [plane hi/lo]
  ConvGeom g;
  ConvEpi e;
  int nsplit;                          // 1: hi*hi only, 3: split product
  int n_tile, n_blocks, tiles_x, tiles_y, total_tiles;
  int bw_log2, sr, S, tr, dmax, box_rows;
  int xshare, box_w, dxmax;            // xshare: ONE (bw + 2*dxmax)-pixel-wide box per chunk serves every horizontal tap offset
  int na, nw;                          // A / W ring depths
  int wide_b;                          // 1: hi*hi and hi*lo issue as ONE MMA of N = 2*n_tile against the stacked [W_hi; W_lo] tile (A_hi read once)
  int w_box;                           // 1: a W slot holds every tap of an activation box (one barrier round trip per box)
  int w_slot_taps;                     // taps per W slot (1 unless w_box)
  int w_frames;                        // > 0: PER-FRAME weights [frame][tap][cout_pad][kpad] (InstanceNorm folded into the layer, engine.cuh); w_map is 3-D
  int prod_mode;                       // precision probe (EGN_PRODUCTS): 0 = hi*hi + lo*hi + hi*lo, 1 = drop lo*hi (activations act as bf16), 2 = drop hi*lo (weights act as bf16)
  int tap_triples;                     // 1: taps are ordered round-robin over three accumulator groups (phase-lattice MSBlock tail)
  int w_res;                           // 1: every (chunk, tap) weight tile of the layer stays resident in shared memory (loaded once per CTA)
  uint32_t a_plane_bytes, a_box_bytes, w_plane_bytes;
  int acc_bufs;                        // TMEM accumulator buffers (1 or 2)
  int n_aloads;
  int a_C[EGN_MAX_SRC];                // channels of every source buffer
  int l2_prefetch;                     // producer prefetches its next tile's activation boxes into L2
  // upsample-add layers: the half-resolution operand patch of a tile ((tr/2 + 2) x (bw/2 + 2) pixels x n_tile
  // channels, both planes) is staged in shared memory by the epilogue warps with cp.async, one tile ahead
  int up_rows, up_cols;                // patch rows / columns (half-resolution pixels)
  uint32_t up_pix_stride, up_plane_stride, up_patch_bytes;
  int dbg;                             // timing experiments only: 1 = skip epilogue body, 2 = skip MMAs, 4 = skip activation TMA loads, 8 = skip weight TMA loads,
                                       // 32 = skip the epilogue's global stores, 64 = skip its TMEM reads
  int8_t aload_dx[TC_MAX_ALOADS];
  uint8_t aload_tap0[TC_MAX_ALOADS], aload_ntaps[TC_MAX_ALOADS];
  int* err_flag;
  long long* timing;                   // optional [grid][10] cycle counters (EGN_TC_TIMING), see tc_launch
};

// Per-tap / per-chunk / per-box schedule entries, staged in shared memory once per CTA so the
// single-thread producer and MMA issuer never index kernel parameters dynamically (an indexed
// constant-bank load costs hundreds of cycles on the issue path).
struct TcTapStep { uint32_t a_off, d_col, flags, pad; };   // flags: 1 first tap of its box, 2 last tap of its box, 4 first tap of its accumulator group
struct TcChunk { int src, c0, noff, srcC; };   // srcC: channels of the source buffer (phase-lattice loads)
struct TcLoad { int dx, tap0, ntaps, pad; };
#define TC_SCHED_BYTES (32 * 16 + EGN_MAX_CHUNKS * 16 + TC_MAX_ALOADS * 16)

#define TC_THREADS 320
#define TC_EPI_WARPS 8
#define TC_ACC_STRIDE 256   // TMEM columns between the two accumulator buffers
#define TC_UP_PAIRS 6        // (pixel, 8-channel) pairs of the staged half-resolution patch per epilogue thread

namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// Bounded wait: a protocol bug must surface as a trap, never as a hung GPU.
__device__ __forceinline__ uint32_t mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  return done;
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int* err_flag, int code) {
  if (mbar_try(bar, parity)) return;              // common case: no clock read, no loop
  uint32_t done = 0;
  const long long t0 = clock64();
  for (;;) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) return;
    if (clock64() - t0 > 2000000000LL) break;      // ~1 s: no pipeline event takes that long
  }
  if (err_flag) atomicExch(err_flag, code);
  __threadfence_system();
  __trap();
}

__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint32_t bar, uint32_t dst,
                                            int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

__device__ __forceinline__ void tma_load_5d(const CUtensorMap* map, uint32_t bar, uint32_t dst,
                                            int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

__device__ __forceinline__ void tma_load_3d(const CUtensorMap* map, uint32_t bar, uint32_t dst, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint32_t bar, uint32_t dst,
                                            int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}

// Pulls a box into L2 only (no shared-memory destination, no barrier).
__device__ __forceinline__ void tma_prefetch_4d(const CUtensorMap* map, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global [%0, {%1, %2, %3, %4}];"
               ::"l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

__device__ __forceinline__ void prefetch_map(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// K-major, 64-byte swizzle: rows of 64 B (32 bf16), 8-row atoms 512 B apart (SBO), version 1.
// `sbo`: bytes between consecutive 8-row groups.  The start address need not be aligned to the 512-byte
// swizzle atom and sbo may be any multiple of 64: the XOR pattern is a function of the absolute
// shared-memory address (tools/desc_shift_probe.cu), which the shared activation box relies on.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t sbo = 512) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;                 // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(sbo >> 4) << 32;        // stride byte offset
  d |= (uint64_t)1 << 46;                 // descriptor version
  d |= (uint64_t)4 << 61;                 // SWIZZLE_64B
  return d;
}

__device__ __forceinline__ void mma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc,
                                         uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void mma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
               ::"r"(bar) : "memory");
}

__device__ __forceinline__ void fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t r[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}

__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// All MMAs of one (chunk, tap): NSUB sub-tiles x 2 K-steps x (1 or 3) split products, ordered so
// that consecutive instructions target different accumulators.  Operands arrive as ready-made
// descriptors: shared-memory addresses sit in the low 14 bits of a descriptor in 16-byte units, so the
// caller forms a tap's descriptors by ADDING (offset >> 4) to the box / slot descriptor (one 64-bit add
// each instead of rebuilding them) - with a single sub-tile per tile the issuing thread is otherwise
// the bottleneck (~45 uniform instructions for 4 MMAs).
template <int NSUB, int NPL, int WIDE>
__device__ __forceinline__ void issue_tap(uint64_t dA_hi, uint32_t a_plane16, uint64_t dW_hi, uint32_t w_plane16,
                                          uint32_t d_tmem, uint32_t sub_cols, uint32_t idesc, uint32_t idesc_wide,
                                          uint32_t first, uint32_t a_sub16, int pm = 0) {
  const uint64_t dA_lo = dA_hi + a_plane16, dW_lo = dW_hi + w_plane16;
#pragma unroll
  for (int k = 0; k < EGN_KC / 16; ++k) {
    const uint64_t koff = (uint64_t)((k * 32) >> 4);        // 16 bf16 = 32 bytes along K
    const uint32_t acc = (k == 0) ? (first ^ 1u) : 1u;
    if (NPL == 2 && WIDE) {
      // A_hi x [W_hi; W_lo] -> columns [0, n) += hi*hi, [n, 2n) += hi*lo ; A_lo x W_hi -> columns [0, n)
#pragma unroll
      for (int s = 0; s < NSUB; ++s)
        mma_bf16(d_tmem + s * sub_cols, dA_hi + koff + (uint64_t)(s * a_sub16), dW_hi + koff, idesc_wide, acc);
      if (pm != 1) {
#pragma unroll
        for (int s = 0; s < NSUB; ++s)
          mma_bf16(d_tmem + s * sub_cols, dA_lo + koff + (uint64_t)(s * a_sub16), dW_hi + koff, idesc, 1u);
      }
    } else {
#pragma unroll
      for (int s = 0; s < NSUB; ++s)
        mma_bf16(d_tmem + s * sub_cols, dA_hi + koff + (uint64_t)(s * a_sub16), dW_hi + koff, idesc, acc);
      if (NPL == 2) {
        if (pm != 1) {
#pragma unroll
          for (int s = 0; s < NSUB; ++s)
            mma_bf16(d_tmem + s * sub_cols, dA_lo + koff + (uint64_t)(s * a_sub16), dW_hi + koff, idesc, 1u);
        }
        if (pm != 2) {
#pragma unroll
          for (int s = 0; s < NSUB; ++s)
            mma_bf16(d_tmem + s * sub_cols, dA_hi + koff + (uint64_t)(s * a_sub16), dW_lo + koff, idesc, 1u);
        }
      }
    }
  }
}

// Three taps that feed three DIFFERENT accumulator groups (the phase-lattice MSBlock tail orders its taps
// round-robin over the dilation groups), wide layout, one sub-tile: the MMAs of the three taps are
// interleaved so that consecutive instructions never target the same TMEM columns.
__device__ __forceinline__ void issue_triple_wide(const uint64_t (&dA)[3], uint32_t a_plane16, const uint64_t (&dW)[3],
                                                  const uint32_t (&dcol)[3], uint32_t idesc, uint32_t idesc_wide,
                                                  const uint32_t (&first)[3]) {
#pragma unroll
  for (int k = 0; k < EGN_KC / 16; ++k) {
    const uint64_t koff = (uint64_t)((k * 32) >> 4);
#pragma unroll
    for (int g = 0; g < 3; ++g)
      mma_bf16(dcol[g], dA[g] + koff, dW[g] + koff, idesc_wide, (k == 0) ? (first[g] ^ 1u) : 1u);
#pragma unroll
    for (int g = 0; g < 3; ++g)
      mma_bf16(dcol[g], dA[g] + a_plane16 + koff, dW[g] + koff, idesc, 1u);
  }
}

// Two fp32 -> packed bf16x2 hi and lo words (hi = rn(x), lo = rn(x - hi)).
__device__ __forceinline__ void split_pack2(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  const float ra = a - __uint_as_float(hi << 16);
  const float rb = b - __uint_as_float(hi & 0xffff0000u);
  const __nv_bfloat162 l = __floats2bfloat162_rn(ra, rb);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

// 16 consecutive channels of one pixel to both planes as one 32-byte store each (full sectors).
#ifndef ST16
#define ST16 "st.global.v8.b32"
#endif
__device__ __forceinline__ void store16(bf16* hi, bf16* lo, size_t idx, const float v[16]) {
  uint32_t h[8], l[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) split_pack2(v[2 * i], v[2 * i + 1], h[i], l[i]);
  asm volatile(ST16 " [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"l"(hi + idx), "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]), "r"(h[4]), "r"(h[5]), "r"(h[6]), "r"(h[7]) : "memory");
  asm volatile(ST16 " [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"l"(lo + idx), "r"(l[0]), "r"(l[1]), "r"(l[2]), "r"(l[3]), "r"(l[4]), "r"(l[5]), "r"(l[6]), "r"(l[7]) : "memory");
}

__device__ __forceinline__ void load16(const bf16* hi, const bf16* lo, size_t idx, float v[16]) {
  uint32_t h[8], l[8];
  asm volatile("ld.global.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(h[0]), "=r"(h[1]), "=r"(h[2]), "=r"(h[3]), "=r"(h[4]), "=r"(h[5]), "=r"(h[6]), "=r"(h[7]) : "l"(hi + idx));
  asm volatile("ld.global.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(l[0]), "=r"(l[1]), "=r"(l[2]), "=r"(l[3]), "=r"(l[4]), "=r"(l[5]), "=r"(l[6]), "=r"(l[7]) : "l"(lo + idx));
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    v[2 * i] = __uint_as_float(h[i] << 16) + __uint_as_float(l[i] << 16);
    v[2 * i + 1] = __uint_as_float(h[i] & 0xffff0000u) + __uint_as_float(l[i] & 0xffff0000u);
  }
}

// Sums 32 per-lane values across the warp in 31 shuffles; lane L returns the total of v[L].
__device__ __forceinline__ float warp_reduce32x32(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = upper ? v[i] : v[i + off];
      const float keep = upper ? v[i + off] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

}  // namespace tc

// UP: the layer adds the bilinear upsample of a half-resolution tensor in its epilogue (decoder 1x1
// convolutions); a separate instantiation keeps the registers of that path out of every other layer.
template <bool UP>
__global__ void __launch_bounds__(TC_THREADS, 1) conv_tc_kernel(const __grid_constant__ TcParams p) {
  using namespace tc;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;      // swizzled tiles: 1024-B aligned slots
  uint8_t* smem = smem_raw + (base - raw_addr);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);   // warp-uniform by construction
  const int lane = threadIdx.x & 31;
  const int nplanes = p.nsplit == 1 ? 1 : 2;
  const uint32_t a_slot_bytes = nplanes * p.a_plane_bytes;
  const uint32_t w_tap_bytes = nplanes * p.w_plane_bytes;
  const uint32_t w_slot_bytes = p.w_slot_taps * w_tap_bytes;
  const int na = p.na, nw = p.nw;

  const uint32_t w_base = base + na * a_slot_bytes;
  const uint32_t bar_base = w_base + nw * w_slot_bytes;
  auto a_full = [&](int s) { return bar_base + 8u * s; };
  auto a_empty = [&](int s) { return bar_base + 8u * (na + s); };
  auto w_full = [&](int s) { return bar_base + 8u * (2 * na + s); };
  auto w_empty = [&](int s) { return bar_base + 8u * (2 * na + nw + s); };
  auto tfull_bar = [&](int b) { return bar_base + 8u * (2 * na + 2 * nw + b); };
  auto tempty_bar = [&](int b) { return bar_base + 8u * (2 * na + 2 * nw + 2 + b); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + (bar_base - base) + 8 * (2 * na + 2 * nw + 4));

  if (warp == 0 && lane == 0) {
    prefetch_map(&p.w_map[0]);
    prefetch_map(&p.a_map[0][0]);
    for (int s = 0; s < na; ++s) { mbar_init(a_full(s), 1); mbar_init(a_empty(s), 1); }
    for (int s = 0; s < nw; ++s) { mbar_init(w_full(s), 1); mbar_init(w_empty(s), 1); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull_bar(b), 1);
      mbar_init(tempty_bar(b), TC_EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // epilogue constants (bias per accumulator group, optional post-activation affine, MSBlock score
  // weights) are read from shared memory by the epilogue warps
  float* stat_buf = reinterpret_cast<float*>(smem + (bar_base - base) + 8 * (2 * na + 2 * nw + 4) + 16);
  float* c_bias = stat_buf + 2 * 2 * 4 * 32;       // [groups * cout_pad] <= 768
  float* c_scale = c_bias + 768;                   // [cout_pad] <= 512   (MSBlock: score_w[64])
  float* c_shift = c_scale + 512;                  // [cout_pad] <= 512
  TcTapStep* s_tap = reinterpret_cast<TcTapStep*>(c_shift + 512);
  TcChunk* s_chunk = reinterpret_cast<TcChunk*>(s_tap + 32);
  TcLoad* s_load = reinterpret_cast<TcLoad*>(s_chunk + EGN_MAX_CHUNKS);
  const uint32_t up_base = smem_u32(s_load + TC_MAX_ALOADS);   // two staged half-resolution patches (UP layers only)
  if ((int)threadIdx.x < p.g.ntaps) {
    const int t = threadIdx.x;
    TcTapStep st;
    st.a_off = p.xshare ? (uint32_t)((p.g.tap_dy[t] + p.dmax) * p.box_w + p.g.tap_dx[t] + p.dxmax) * 64u
                        : (uint32_t)((p.g.tap_dy[t] + p.dmax) << p.bw_log2) * 64u;
    st.d_col = (uint32_t)(p.g.tap_grp[t] * p.n_tile * (p.wide_b ? 2 : 1));
    st.flags = 0; st.pad = 0;
    for (int l = 0; l < p.n_aloads; ++l) {
      if (t == p.aload_tap0[l]) st.flags |= 1u;
      if (t == p.aload_tap0[l] + p.aload_ntaps[l] - 1) st.flags |= 2u;
    }
    bool seen = false;
    for (int u = 0; u < t; ++u) seen |= (p.g.tap_grp[u] == p.g.tap_grp[t]);
    if (!seen) st.flags |= 4u;
    s_tap[t] = st;
  }
  if ((int)threadIdx.x >= 64 && (int)threadIdx.x < 64 + p.g.nchunks) {
    const int c = threadIdx.x - 64;
    s_chunk[c] = TcChunk{(int)p.g.chunk_src[c], (int)p.g.chunk_c0[c], p.g.chunk_noff[c], p.a_C[p.g.chunk_src[c]]};
  }
  if ((int)threadIdx.x >= 128 && (int)threadIdx.x < 128 + p.n_aloads) {
    const int l = threadIdx.x - 128;
    s_load[l] = TcLoad{(int)p.aload_dx[l], (int)p.aload_tap0[l], (int)p.aload_ntaps[l], 0};
  }
  for (int i = threadIdx.x; i < p.g.groups * p.g.cout_pad; i += blockDim.x) c_bias[i] = p.e.bias[i];
  if (p.e.post_scale)
    for (int i = threadIdx.x; i < p.g.cout_pad; i += blockDim.x) { c_scale[i] = p.e.post_scale[i]; c_shift[i] = p.e.post_shift[i]; }
  if (p.e.mode == CONV_MSBLOCK)
    for (int i = threadIdx.x; i < 64; i += blockDim.x) c_scale[i] = p.e.score_w[i];
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int bw = 1 << p.bw_log2;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int as = 0, ws = 0;
      uint32_t aph = 0, wph = 0;
      const int nchunks = p.g.nchunks, n_aloads = p.n_aloads, n_blocks = p.n_blocks, tiles_x = p.tiles_x, tiles_y = p.tiles_y;
      const uint32_t a_tx = nplanes * p.a_box_bytes, w_tx = nplanes * p.w_plane_bytes;
      const uint32_t a_plane = p.a_plane_bytes, w_plane = p.w_plane_bytes;
      const int cout_pad = p.g.cout_pad, n_tile = p.n_tile, dbg = p.dbg;
      const bool ptm = p.timing != nullptr;
      long long ptm_a = 0, ptm_w = 0, ptm_c0 = 0;
      const long long ptm_start = ptm ? clock64() : 0;
      if (p.w_res) {
        // small layers: all weight tiles [chunk][tap] are loaded once and stay in shared memory
        const int ntaps_all = p.g.ntaps;
        if (dbg & 8) {
          mbar_arrive(w_full(0));
        } else {
          mbar_expect_tx(w_full(0), w_tx * ntaps_all * nchunks);
          for (int c = 0; c < nchunks; ++c)
            for (int t = 0; t < ntaps_all; ++t) {
              const uint32_t sW = w_base + (uint32_t)(c * ntaps_all + t) * w_tap_bytes;
              tma_load_2d(&p.w_map[0], w_full(0), sW, c * EGN_KC, t * cout_pad);
              if (nplanes == 2) tma_load_2d(&p.w_map[1], w_full(0), sW + w_plane, c * EGN_KC, t * cout_pad);
            }
        }
      }
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int nb = tile % n_blocks;
        int rest = tile / n_blocks;
        const int tx = rest % tiles_x;
        rest /= tiles_x;
        const int ty = rest % tiles_y;
        const int n = rest / tiles_y;
        const int x0 = tx * bw;
        const int y0 = ty * p.tr - p.dmax;
        const int wrow0 = nb * n_tile;
        // phase lattice: "frame" n is (image, row phase, column phase); the 5-D map addresses
        // (phase_x * C + channel, lattice x, lattice y, phase_y, image)
        const int phd = p.g.phase;
        int n_img = n, ph_y = 0, ph_x = 0;
        if (phd) { const int dd = phd * phd; n_img = n / dd; const int ph = n - n_img * dd; ph_y = ph / phd; ph_x = ph - ph_y * phd; }
        for (int c = 0; c < nchunks; ++c) {
          const TcChunk ck = s_chunk[c];
          const int nn = n_img + ck.noff;
          const CUtensorMap* map_hi = &p.a_map[0][ck.src];
          const CUtensorMap* map_lo = &p.a_map[1][ck.src];
          for (int l = 0; l < n_aloads; ++l) {
            const TcLoad ld = s_load[l];
            if (ptm) ptm_c0 = clock64();
            mbar_wait(a_empty(as), aph ^ 1u, p.err_flag, 1);
            if (ptm) ptm_a += clock64() - ptm_c0;
            const uint32_t sA = base + as * a_slot_bytes;
            if (dbg & 4) {
              mbar_arrive(a_full(as));
            } else {
              mbar_expect_tx(a_full(as), a_tx);
              if (phd) {
                tma_load_5d(map_hi, a_full(as), sA, ph_x * ck.srcC + ck.c0, x0 + ld.dx, y0, ph_y, nn);
                if (nplanes == 2) tma_load_5d(map_lo, a_full(as), sA + a_plane, ph_x * ck.srcC + ck.c0, x0 + ld.dx, y0, ph_y, nn);
              } else {
                tma_load_4d(map_hi, a_full(as), sA, ck.c0, x0 + ld.dx, y0, nn);
                if (nplanes == 2) tma_load_4d(map_lo, a_full(as), sA + a_plane, ck.c0, x0 + ld.dx, y0, nn);
              }
            }
            if (++as == na) { as = 0; aph ^= 1u; }
            if (p.w_res) continue;
            if (p.w_box) {
              // runs of up to w_slot_taps taps share one weight slot / barrier round trip
              for (int t0 = 0; t0 < ld.ntaps; t0 += p.w_slot_taps) {
                const int nrun = min(p.w_slot_taps, ld.ntaps - t0);
                if (ptm) ptm_c0 = clock64();
                mbar_wait(w_empty(ws), wph ^ 1u, p.err_flag, 2);
                if (ptm) ptm_w += clock64() - ptm_c0;
                const uint32_t sW = w_base + ws * w_slot_bytes;
                if (dbg & 8) {
                  mbar_arrive(w_full(ws));
                } else {
                  mbar_expect_tx(w_full(ws), w_tx * nrun);
                  for (int j = 0; j < nrun; ++j) {
                    const int wrow = (ld.tap0 + t0 + j) * cout_pad + wrow0;
                    if (p.w_frames) {
                      tma_load_3d(&p.w_map[0], w_full(ws), sW + j * w_tap_bytes, c * EGN_KC, wrow, n);
                      if (nplanes == 2) tma_load_3d(&p.w_map[1], w_full(ws), sW + j * w_tap_bytes + w_plane, c * EGN_KC, wrow, n);
                    } else {
                      tma_load_2d(&p.w_map[0], w_full(ws), sW + j * w_tap_bytes, c * EGN_KC, wrow);
                      if (nplanes == 2) tma_load_2d(&p.w_map[1], w_full(ws), sW + j * w_tap_bytes + w_plane, c * EGN_KC, wrow);
                    }
                  }
                }
                if (++ws == nw) { ws = 0; wph ^= 1u; }
              }
            } else {
              for (int t = ld.tap0; t < ld.tap0 + ld.ntaps; ++t) {
                if (ptm) ptm_c0 = clock64();
                mbar_wait(w_empty(ws), wph ^ 1u, p.err_flag, 2);
                if (ptm) ptm_w += clock64() - ptm_c0;
                const uint32_t sW = w_base + ws * w_slot_bytes;
                const int wrow = t * cout_pad + wrow0;
                if (dbg & 8) {
                  mbar_arrive(w_full(ws));
                } else {
                  mbar_expect_tx(w_full(ws), w_tx);
                  if (p.w_frames) {
                    tma_load_3d(&p.w_map[0], w_full(ws), sW, c * EGN_KC, wrow, n);
                    if (nplanes == 2) tma_load_3d(&p.w_map[1], w_full(ws), sW + w_plane, c * EGN_KC, wrow, n);
                  } else {
                    tma_load_2d(&p.w_map[0], w_full(ws), sW, c * EGN_KC, wrow);
                    if (nplanes == 2) tma_load_2d(&p.w_map[1], w_full(ws), sW + w_plane, c * EGN_KC, wrow);
                  }
                }
                if (++ws == nw) { ws = 0; wph ^= 1u; }
              }
            }
          }
        }
      }
      if (ptm) {
        long long* o = p.timing + (size_t)blockIdx.x * 14;
        o[6] = ptm_a; o[7] = ptm_w; o[8] = clock64() - ptm_start;
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    // The whole warp walks the schedule with warp-uniform control flow (so addresses and
    // descriptors live in uniform registers and the MMAs issue back to back); per tap it reads one
    // shared-memory entry, waits on the operand barriers (usually already complete) and one
    // elected lane issues NSUB x 2 x (1 or 3) MMAs plus the commits that free the slots.
    {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) |
                             ((uint32_t)(p.n_tile >> 3) << 17) | ((128u >> 4) << 24);
      const uint32_t idesc_wide = (1u << 4) | (1u << 7) | (1u << 10) |
                                  ((uint32_t)((2 * p.n_tile) >> 3) << 17) | ((128u >> 4) << 24);
      const uint32_t sub_cols = (uint32_t)(p.g.groups * p.n_tile * (p.wide_b ? 2 : 1));
      // shared box: consecutive 8-pixel groups of a sub-tile are consecutive image rows, box_w pixels apart
      const uint32_t a_sbo = p.xshare ? (uint32_t)p.box_w * 64u : 512u;
      const uint32_t a_sub16 = (p.xshare ? (uint32_t)(p.sr * p.box_w) * 64u : 8192u) >> 4;
      const int nchunks = p.g.nchunks, ntaps = p.g.ntaps;
      const uint32_t a_plane16 = p.a_plane_bytes >> 4, w_plane16 = p.w_plane_bytes >> 4, w_tap16 = w_tap_bytes >> 4;
      const int dbg = p.dbg, pm = p.prod_mode;
      int as = 0, ws = 0;
      uint32_t aph = 0, wph = 0;
      int use = 0;                                   // tiles issued so far by this CTA
      const bool tm = p.timing != nullptr;
      long long tm_te = 0, tm_a = 0, tm_w = 0, tm_i = 0, tm_k = 0, tm_c0 = 0, tm_x = 0, tm_h = 0, tm_l = 0;
      const long long tm_start = tm ? clock64() : 0;
      if (p.w_res) {
        mbar_wait(w_full(0), 0u, p.err_flag, 5);
        fence_after();
      }
      // the issuing warp only needs the tile's row index (the last tile row may hold fewer sub-tiles): it is advanced
      // incrementally, the integer divisions of a full decode would sit on the critical path of every tile
      const int row_len = p.n_blocks * p.tiles_x;                 // tiles per tile row
      const int step_q = (int)gridDim.x / row_len, step_r = (int)gridDim.x % row_len, step_ty = step_q % p.tiles_y;
      int dec_r = (int)blockIdx.x % row_len, ty = ((int)blockIdx.x / row_len) % p.tiles_y;
      const int nsub_last = min(p.S, (p.g.H - (p.tiles_y - 1) * p.tr + p.sr - 1) / p.sr);
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++use) {
        if (tm) tm_l = clock64();
        const int nsub = ty == p.tiles_y - 1 ? nsub_last : p.S;
        {
          dec_r += step_r;
          int adv = step_ty;
          if (dec_r >= row_len) { dec_r -= row_len; ++adv; }
          ty += adv;
          if (ty >= p.tiles_y) ty -= p.tiles_y;
          if (ty >= p.tiles_y) ty -= p.tiles_y;
        }
        const int ab = p.acc_bufs == 2 ? (use & 1) : 0;
        const uint32_t aphase = (p.acc_bufs == 2 ? (use >> 1) : use) & 1u;
        if (tm) { tm_c0 = clock64(); tm_h += tm_c0 - tm_l; }
        mbar_wait(tempty_bar(ab), aphase ^ 1u, p.err_flag, 3);
        if (tm) tm_te += clock64() - tm_c0;
        fence_after();
        const uint32_t d_base = tmem_base + ab * TC_ACC_STRIDE;
        auto run_tile = [&](auto nsub_c, auto npl_c, auto wide_c) {
          constexpr int NSUB = decltype(nsub_c)::value;
          constexpr int NPL = decltype(npl_c)::value;
          constexpr int WIDE = decltype(wide_c)::value;
          if (p.w_res) {
            // resident weights: no weight barriers at all, every tap of a box issues in one go
            const int n_aloads = p.n_aloads;
            for (int c = 0; c < nchunks; ++c) {
              for (int l = 0; l < n_aloads; ++l) {
                const TcLoad ld = s_load[l];
                if (tm) tm_c0 = clock64();
                mbar_wait(a_full(as), aph, p.err_flag, 4);
                if (tm) { const long long now = clock64(); tm_a += now - tm_c0; tm_c0 = now; }
                fence_after();
                const uint64_t dA = make_desc(base + as * a_slot_bytes, a_sbo);
                uint64_t dW = make_desc(w_base + (uint32_t)(c * ntaps + ld.tap0) * w_tap_bytes);
                if (p.tap_triples && NSUB == 1 && WIDE) {
                  // taps come in triples over the three accumulator groups: interleave their MMAs
                  if (elect_one()) {
                    if (!(dbg & 2)) {
                      for (int j = 0; j < ld.ntaps; j += 3, dW += 3 * w_tap16) {
                        const TcTapStep t0 = s_tap[ld.tap0 + j], t1 = s_tap[ld.tap0 + j + 1], t2 = s_tap[ld.tap0 + j + 2];
                        const uint64_t tA[3] = {dA + (t0.a_off >> 4), dA + (t1.a_off >> 4), dA + (t2.a_off >> 4)};
                        const uint64_t tW[3] = {dW, dW + w_tap16, dW + 2 * w_tap16};
                        const uint32_t tD[3] = {d_base + t0.d_col, d_base + t1.d_col, d_base + t2.d_col};
                        const uint32_t tF[3] = {(c == 0 && (t0.flags & 4u)) ? 1u : 0u, (c == 0 && (t1.flags & 4u)) ? 1u : 0u,
                                                (c == 0 && (t2.flags & 4u)) ? 1u : 0u};
                        issue_triple_wide(tA, a_plane16, tW, tD, idesc, idesc_wide, tF);
                      }
                    }
                    if (tm) { const long long now = clock64(); tm_i += now - tm_c0; tm_c0 = now; }
                    mma_commit(a_empty(as));
                    if (tm) tm_k += clock64() - tm_c0;
                  }
                } else
                if (elect_one()) {
                  if (!(dbg & 2)) {
                    TcTapStep nxt = s_tap[ld.tap0];
                    for (int j = 0; j < ld.ntaps; ++j, dW += w_tap16) {
                      const TcTapStep cur = nxt;
                      if (j + 1 < ld.ntaps) nxt = s_tap[ld.tap0 + j + 1];
                      const uint32_t first = (c == 0 && (cur.flags & 4u)) ? 1u : 0u;
                      issue_tap<NSUB, NPL, WIDE>(dA + (cur.a_off >> 4), a_plane16, dW, w_plane16, d_base + cur.d_col, sub_cols, idesc, idesc_wide, first, a_sub16, pm);
                    }
                  }
                  if (tm) { const long long now = clock64(); tm_i += now - tm_c0; tm_c0 = now; }
                  mma_commit(a_empty(as));
                  if (tm) tm_k += clock64() - tm_c0;
                }
                __syncwarp();
                if (++as == na) { as = 0; aph ^= 1u; }
              }
            }
            return;
          }
          if (p.w_box) {
            // one weight-barrier round trip per run of up to three taps instead of one per tap
            const int n_aloads = p.n_aloads;
            for (int c = 0; c < nchunks; ++c) {
              for (int l = 0; l < n_aloads; ++l) {
                const TcLoad ld = s_load[l];
                if (tm) tm_c0 = clock64();
                mbar_wait(a_full(as), aph, p.err_flag, 4);
                if (tm) tm_a += clock64() - tm_c0;
                const uint64_t dA = make_desc(base + as * a_slot_bytes, a_sbo);
                for (int t0 = 0; t0 < ld.ntaps; t0 += p.w_slot_taps) {
                  const int nrun = min(p.w_slot_taps, ld.ntaps - t0);
                  const bool last_run = t0 + nrun >= ld.ntaps;
                  if (tm) tm_c0 = clock64();
                  mbar_wait(w_full(ws), wph, p.err_flag, 5);
                  if (tm) { const long long now = clock64(); tm_w += now - tm_c0; tm_c0 = now; }
                  fence_after();
                  uint64_t dW = make_desc(w_base + ws * w_slot_bytes);
                  if (elect_one()) {
                    if (!(dbg & 2)) {
                      TcTapStep nxt = s_tap[ld.tap0 + t0];
                      for (int j = 0; j < nrun; ++j, dW += w_tap16) {
                        const TcTapStep cur = nxt;
                        if (j + 1 < nrun) nxt = s_tap[ld.tap0 + t0 + j + 1];
                        const uint32_t first = (c == 0 && (cur.flags & 4u)) ? 1u : 0u;
                        issue_tap<NSUB, NPL, WIDE>(dA + (cur.a_off >> 4), a_plane16, dW, w_plane16, d_base + cur.d_col, sub_cols, idesc, idesc_wide, first, a_sub16, pm);
                      }
                    }
                    if (tm) { const long long now = clock64(); tm_i += now - tm_c0; tm_c0 = now; }
                    mma_commit(w_empty(ws));
                    if (last_run) mma_commit(a_empty(as));
                    if (tm) tm_k += clock64() - tm_c0;
                  }
                  __syncwarp();
                  if (++ws == nw) { ws = 0; wph ^= 1u; }
                }
                if (++as == na) { as = 0; aph ^= 1u; }
              }
            }
            return;
          }
          for (int c = 0; c < nchunks; ++c) {
            uint64_t dA = 0;
            TcTapStep nxt = s_tap[0];
            for (int t = 0; t < ntaps; ++t) {
              const TcTapStep cur = nxt;
              if (t + 1 < ntaps) nxt = s_tap[t + 1];
              if (cur.flags & 1u) {
                if (tm) tm_c0 = clock64();
                mbar_wait(a_full(as), aph, p.err_flag, 4);
                if (tm) tm_a += clock64() - tm_c0;
                dA = make_desc(base + as * a_slot_bytes, a_sbo);
              }
              if (tm) tm_c0 = clock64();
              mbar_wait(w_full(ws), wph, p.err_flag, 5);
              if (tm) { const long long now = clock64(); tm_w += now - tm_c0; tm_c0 = now; }
              fence_after();
              const uint64_t dW = make_desc(w_base + ws * w_slot_bytes);
              const uint32_t first = (c == 0 && (cur.flags & 4u)) ? 1u : 0u;
              if (elect_one()) {
                if (!(dbg & 2))
                  issue_tap<NSUB, NPL, WIDE>(dA + (cur.a_off >> 4), a_plane16, dW, w_plane16, d_base + cur.d_col, sub_cols, idesc, idesc_wide, first, a_sub16, pm);
                if (tm) { const long long now = clock64(); tm_i += now - tm_c0; tm_c0 = now; }
                mma_commit(w_empty(ws));             // frees the weight slot when these MMAs retire
                if (cur.flags & 2u) mma_commit(a_empty(as));   // last tap of this box
                if (tm) tm_k += clock64() - tm_c0;
              }
              __syncwarp();
              if (++ws == nw) { ws = 0; wph ^= 1u; }
              if (cur.flags & 2u) {
                if (++as == na) { as = 0; aph ^= 1u; }
              }
            }
          }
        };
        using std::integral_constant;
        typedef integral_constant<int, 0> I0;
        typedef integral_constant<int, 1> I1;
        typedef integral_constant<int, 2> I2;
        typedef integral_constant<int, 3> I3;
        typedef integral_constant<int, 4> I4;
        if (nplanes == 2 && p.wide_b) {
          if (nsub == 4) run_tile(I4{}, I2{}, I1{});
          else if (nsub == 2) run_tile(I2{}, I2{}, I1{});
          else if (nsub == 3) run_tile(I3{}, I2{}, I1{});
          else run_tile(I1{}, I2{}, I1{});
        } else if (nplanes == 2) {
          if (nsub == 4) run_tile(I4{}, I2{}, I0{});
          else if (nsub == 2) run_tile(I2{}, I2{}, I0{});
          else if (nsub == 3) run_tile(I3{}, I2{}, I0{});
          else run_tile(I1{}, I2{}, I0{});
        } else {
          if (nsub == 4) run_tile(I4{}, I1{}, I0{});
          else if (nsub == 2) run_tile(I2{}, I1{}, I0{});
          else if (nsub == 3) run_tile(I3{}, I1{}, I0{});
          else run_tile(I1{}, I1{}, I0{});
        }
        if (tm) tm_l = clock64();
        if (elect_one()) mma_commit(tfull_bar(ab));  // accumulators complete -> epilogue
        __syncwarp();
        if (tm) tm_x += clock64() - tm_l;
      }
      if (tm) {
        // counters are summed over lanes by the elected thread only for tm_i / tm_k
        long long* o = p.timing + (size_t)blockIdx.x * 14;
        if (lane == 0) { o[0] = clock64() - tm_start; o[1] = tm_te; o[2] = tm_a; o[3] = tm_w; o[5] = use; o[10] = tm_x; o[11] = tm_h; }
        if (tm_i) { o[4] = tm_i; o[9] = tm_k; }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..9)
    const int quarter = warp & 3;                  // TMEM lane quarter this warp may read
    const int half = (warp - 2) >> 2;              // two warps share a quarter and split the work
    const int m = quarter * 32 + lane;             // accumulator row = pixel within the sub-tile
    const int mr = m >> p.bw_log2, mc = m & (bw - 1);
    int stat_flip = 0;
    int use = 0;
    long long e_wait = 0, e_busy = 0;
    // upsample-add layers (decoder 1x1 convolutions): the (tr/2 + 2) x (bw/2 + 2) half-resolution pixels a tile
    // blends (indices clamped like F.interpolate, align_corners=False) x n_tile channels are staged in shared
    // memory one tile ahead: cp.async of the hi and lo 16-byte pieces (8 channels) side by side, which the SAME
    // thread turns into 8 fp32 values in place once they have landed (up_convert) - the blend then reads plain
    // fp32 and every half-resolution value is converted once instead of once per output pixel that uses it.
    // Each thread owns up to TC_UP_PAIRS (pixel, 8-channel) pairs, fixed for the whole launch.
    int up_desc[TC_UP_PAIRS];
    int up_n = 0;
    if (UP) {
      const int c8n = p.n_tile >> 3;
      const int npairs = p.up_rows * p.up_cols * c8n;
#pragma unroll
      for (int j = 0; j < TC_UP_PAIRS; ++j) {
        const int i = (int)threadIdx.x - 64 + j * (TC_EPI_WARPS * 32);
        up_desc[j] = 0;
        if (i < npairs) {
          const int c8 = i % c8n, pp = i / c8n;
          up_desc[j] = (pp / p.up_cols) | ((pp % p.up_cols) << 8) | (c8 << 16);
          up_n = j + 1;
        }
      }
    }
    auto up_stage = [&](int t, int buf) {
      const int nb2 = t % p.n_blocks;
      int rest2 = t / p.n_blocks;
      const int tx2 = rest2 % p.tiles_x;
      rest2 /= p.tiles_x;
      const int ty2 = rest2 % p.tiles_y;
      const int n2 = rest2 / p.tiles_y;
      const int Hi = p.g.H >> 1, Wi = p.g.W >> 1;
      const int y0 = ((ty2 * p.tr) >> 1) - 1, x0 = ((tx2 * bw) >> 1) - 1;
      const uint32_t dst0 = up_base + (uint32_t)buf * p.up_patch_bytes;
      const size_t cbase = (size_t)p.e.up_coff + (size_t)nb2 * p.n_tile;
      const size_t fbase = (size_t)n2 * Hi * Wi;
#pragma unroll
      for (int j = 0; j < TC_UP_PAIRS; ++j) {
        if (j < up_n) {
          const int row = up_desc[j] & 255, col = (up_desc[j] >> 8) & 255, c8 = up_desc[j] >> 16;
          const int yy = min(max(y0 + row, 0), Hi - 1), xx = min(max(x0 + col, 0), Wi - 1);
          const size_t src = (fbase + (size_t)yy * Wi + xx) * p.e.up_C + cbase + c8 * 8;
          const uint32_t dst = dst0 + (uint32_t)(row * p.up_cols + col) * p.up_pix_stride + c8 * 32;
          cp_async16(dst, p.e.up_hi + src);
          cp_async16(dst + 16, p.e.up_lo + src);
        }
      }
      cp_async_commit();
    };
    auto up_convert = [&](int buf) {
      const uint32_t dst0 = up_base + (uint32_t)buf * p.up_patch_bytes;
#pragma unroll
      for (int j = 0; j < TC_UP_PAIRS; ++j) {
        if (j < up_n) {
          const int row = up_desc[j] & 255, col = (up_desc[j] >> 8) & 255, c8 = up_desc[j] >> 16;
          const uint32_t dst = dst0 + (uint32_t)(row * p.up_cols + col) * p.up_pix_stride + c8 * 32;
          const uint4 h = lds128(dst), l = lds128(dst + 16);
          const uint32_t hh[4] = {h.x, h.y, h.z, h.w}, ll[4] = {l.x, l.y, l.z, l.w};
          uint32_t f[8];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            f[2 * i] = __float_as_uint(__uint_as_float(hh[i] << 16) + __uint_as_float(ll[i] << 16));
            f[2 * i + 1] = __float_as_uint(__uint_as_float(hh[i] & 0xffff0000u) + __uint_as_float(ll[i] & 0xffff0000u));
          }
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(f[0]), "r"(f[1]), "r"(f[2]), "r"(f[3]) : "memory");
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + 16), "r"(f[4]), "r"(f[5]), "r"(f[6]), "r"(f[7]) : "memory");
        }
      }
    };
    if (UP && (int)blockIdx.x < p.total_tiles) up_stage(blockIdx.x, 0);
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++use) {
      const int nb = tile % p.n_blocks;
      int rest = tile / p.n_blocks;
      const int tx = rest % p.tiles_x;
      rest /= p.tiles_x;
      const int ty = rest % p.tiles_y;
      const int n = rest / p.tiles_y;
      const int nsub = min(p.S, (p.g.H - ty * p.tr + p.sr - 1) / p.sr);
      const int ab = p.acc_bufs == 2 ? (use & 1) : 0;
      const uint32_t aphase = (p.acc_bufs == 2 ? (use >> 1) : use) & 1u;
      const int px = tx * bw + mc;

      // fused MSBlock tail with wide MMAs (phase lattice): this warp owns 16 of the 32 channels of its 32 pixels;
      // `o` and the running score are requested before the accumulator is awaited
      const bool tail2 = !UP && p.e.mode == CONV_MSBLOCK && p.wide_b;
      int n_img = n, ph_y = 0, ph_x = 0, phd = 1;
      if (!UP && p.g.phase) {
        phd = p.g.phase;
        const int dd = phd * phd;
        n_img = n / dd;
        const int ph = n - n_img * dd;
        ph_y = ph / phd; ph_x = ph - ph_y * phd;
      }
      auto tail_pix = [&](int s, bool& valid) -> size_t {
        const int qy = ty * p.tr + s * p.sr + mr;
        valid = (qy < p.g.H) && (px < p.g.W);
        return ((size_t)n_img * (p.g.H * phd) + (qy * phd + ph_y)) * (size_t)(p.g.W * phd) + (px * phd + ph_x);
      };
      float o_pre[16];
      float2 sc_pre = make_float2(0.f, 0.f);
      if (tail2) {
        bool valid;
        const size_t pix = tail_pix(0, valid);
#pragma unroll
        for (int i = 0; i < 16; ++i) o_pre[i] = 0.f;
        if (valid) {
          load16(p.e.o_hi, p.e.o_lo, pix * p.e.o_C + p.e.o_coff + (half << 4), o_pre);
          if (half == 0 && p.e.score_accum) sc_pre = reinterpret_cast<const float2*>(p.e.score)[pix];
        }
      }
      long long et0 = 0;
      if (p.timing) et0 = clock64();
      mbar_wait(tfull_bar(ab), aphase, p.err_flag, 6);
      if (p.timing) { const long long now = clock64(); e_wait += now - et0; et0 = now; }
      fence_after();
      const uint32_t tbase = tmem_base + ((uint32_t)(quarter * 32) << 16) + ab * TC_ACC_STRIDE;
      if (UP) {
        // this tile's half-resolution patch has landed (everybody's copies), and everybody is done reading the
        // other buffer (they are past the previous tile): refill it with the next tile's patch
        cp_async_wait_all();
        up_convert(use & 1);
        asm volatile("bar.sync 3, 256;" ::: "memory");
        if (tile + (int)gridDim.x < p.total_tiles) up_stage(tile + gridDim.x, (use + 1) & 1);
      }
      const uint32_t up_buf = up_base + (uint32_t)(use & 1) * p.up_patch_bytes;

      if (p.dbg & 1) {
      } else if (p.e.mode == CONV_STORE || p.e.mode == CONV_LOGITS) {
        const int per_sub = p.n_tile >> 4;         // 16-column groups per sub-tile
        const bool has_post = p.e.post_scale != nullptr;
        for (int g = half; g < per_sub; g += 2) {   // a warp keeps its channel groups across sub-tiles
          const int c0 = g << 4;
          const int cb = nb * p.n_tile + c0;      // first output channel of this 16-column group
          float bias[16];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 b4 = reinterpret_cast<const float4*>(c_bias + cb)[i];
            bias[4 * i] = b4.x; bias[4 * i + 1] = b4.y; bias[4 * i + 2] = b4.z; bias[4 * i + 3] = b4.w;
          }
          float sacc[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) sacc[i] = 0.f;
          for (int s0 = 0; s0 < nsub; s0 += 2) {
            uint32_t r[2][16];
            const uint32_t sstride = p.wide_b ? 2 * p.n_tile : p.n_tile;
            if (p.dbg & 64) {                      // timing experiment: no TMEM reads
#pragma unroll
              for (int i = 0; i < 16; ++i) { r[0][i] = 0; r[1][i] = 0; }
            } else {
            tmem_ld16(tbase + s0 * sstride + c0, r[0]);
            if (s0 + 1 < nsub) tmem_ld16(tbase + (s0 + 1) * sstride + c0, r[1]);
            }
            if (p.wide_b && !(p.dbg & 64)) {       // add the hi*lo partial sums held in columns [n, 2n)
              uint32_t q[2][16];
              tmem_ld16(tbase + s0 * sstride + p.n_tile + c0, q[0]);
              if (s0 + 1 < nsub) tmem_ld16(tbase + (s0 + 1) * sstride + p.n_tile + c0, q[1]);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                r[0][i] = __float_as_uint(__uint_as_float(r[0][i]) + __uint_as_float(q[0][i]));
                r[1][i] = __float_as_uint(__uint_as_float(r[1][i]) + __uint_as_float(q[1][i]));
              }
            }
            tmem_ld_wait();
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              if (s0 + u < nsub) {
                const int py = ty * p.tr + (s0 + u) * p.sr + mr;
                const bool valid = (py < p.g.H) && (px < p.g.W);
                const size_t pix = ((size_t)n * p.g.H + py) * p.g.W + px;
                float v[16];
                if (!UP && p.e.bias_fc) {
                  // InstanceNorm folded into this layer: the mean term is a bias that depends on which taps fall
                  // inside the frame, i.e. on the pixel's border class (3 x 3 classes), per frame
                  const int cls = (py == 0 ? 0 : (py >= p.g.H - 1 ? 2 : 1)) * 3 + (px == 0 ? 0 : (px >= p.g.W - 1 ? 2 : 1));
                  const float4* bp = reinterpret_cast<const float4*>(p.e.bias_fc + ((size_t)n * 9 + cls) * p.g.cout_pad + cb);
#pragma unroll
                  for (int i = 0; i < 4; ++i) {
                    const float4 b4 = __ldg(bp + i);
                    bias[4 * i] = b4.x; bias[4 * i + 1] = b4.y; bias[4 * i + 2] = b4.z; bias[4 * i + 3] = b4.w;
                  }
                }
                if (UP) {
                  // + bilinear x2 upsample of the half-resolution tensor (same arithmetic as common.cuh upsample_add)
                  // from the staged patch: for o = 2i the taps are (i-1, i) with weights (0.25, 0.75), for o = 2i+1
                  // (i, i+1) with (0.75, 0.25), indices clamped to the frame
                  const int Hi = p.g.H >> 1, Wi = p.g.W >> 1;
                  const int iy = py >> 1, jx = px >> 1;
                  const int y0 = ((ty * p.tr) >> 1) - 1, x0 = ((tx * bw) >> 1) - 1;
                  const int ya = min((py & 1) ? iy : max(iy - 1, 0), Hi - 1), yb = min((py & 1) ? iy + 1 : iy, Hi - 1);
                  const int xa = min((px & 1) ? jx : max(jx - 1, 0), Wi - 1), xb = min((px & 1) ? jx + 1 : jx, Wi - 1);
                  const int ra = min(ya - y0, p.up_rows - 1), rb = min(yb - y0, p.up_rows - 1);
                  const int ca = min(xa - x0, p.up_cols - 1), cc = min(xb - x0, p.up_cols - 1);
                  const float wya = (py & 1) ? 0.75f : 0.25f, wyb = 1.f - wya;
                  const float wxa = (px & 1) ? 0.75f : 0.25f, wxb = 1.f - wxa;
                  const float w00 = wya * wxa, w01 = wya * wxb, w10 = wyb * wxa, w11 = wyb * wxb;   // exact (multiples of 1/16)
                  const uint32_t chb = up_buf + (uint32_t)c0 * 4u;
                  const uint32_t a00 = chb + (uint32_t)(ra * p.up_cols + ca) * p.up_pix_stride, a01 = chb + (uint32_t)(ra * p.up_cols + cc) * p.up_pix_stride;
                  const uint32_t a10 = chb + (uint32_t)(rb * p.up_cols + ca) * p.up_pix_stride, a11 = chb + (uint32_t)(rb * p.up_cols + cc) * p.up_pix_stride;
#pragma unroll
                  for (int q4 = 0; q4 < 4; ++q4) {
                    const uint4 t00 = lds128(a00 + q4 * 16), t01 = lds128(a01 + q4 * 16), t10 = lds128(a10 + q4 * 16), t11 = lds128(a11 + q4 * 16);
                    const uint32_t e00[4] = {t00.x, t00.y, t00.z, t00.w}, e01[4] = {t01.x, t01.y, t01.z, t01.w};
                    const uint32_t e10[4] = {t10.x, t10.y, t10.z, t10.w}, e11[4] = {t11.x, t11.y, t11.z, t11.w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                      const int i = q4 * 4 + k;
                      const float up = fmaf(w00, __uint_as_float(e00[k]), fmaf(w01, __uint_as_float(e01[k]),
                                       fmaf(w10, __uint_as_float(e10[k]), w11 * __uint_as_float(e11[k]))));
                      v[i] = apply_act((__uint_as_float(r[u][i]) + bias[i]) + up, p.e.act);
                    }
                  }
                } else {
#pragma unroll
                  for (int i = 0; i < 16; ++i) v[i] = apply_act(__uint_as_float(r[u][i]) + bias[i], p.e.act);
                }
                if (has_post) {
#pragma unroll
                  for (int i = 0; i < 16; ++i) v[i] = fmaf(v[i], c_scale[cb + i], c_shift[cb + i]);
                }
                if (p.e.mode == CONV_LOGITS) {
                  if (valid && cb == 0) {
                    float* dst = p.e.logits + (size_t)n * p.e.logits_c * p.g.H * p.g.W + (size_t)py * p.g.W + px;
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                      if (i < p.e.logits_c) dst[(size_t)i * p.g.H * p.g.W] = v[i];
                  }
                } else if (valid && !(p.dbg & 32)) store16(p.e.out_hi, p.e.out_lo, pix * p.e.out_C + p.e.out_coff + cb, v);
                if (!UP && p.e.pool_hi && cb < p.e.pool_ch) {
                  // fused 2x2 / stride 2 max-pool: the window partners are lanes ^1 (x) and ^bw (y) of this warp (tile
                  // origins are even, H and W are even, so a window is entirely valid or entirely outside the frame);
                  // rounding is monotonic, so pooling before the split equals pooling the stored hi + lo values
                  float m[16];
#pragma unroll
                  for (int i = 0; i < 16; ++i) {
                    const float a = fmaxf(v[i], __shfl_xor_sync(0xffffffffu, v[i], 1));
                    m[i] = fmaxf(a, __shfl_xor_sync(0xffffffffu, a, bw));
                  }
                  if (valid && !((mc | mr) & 1)) {
                    const size_t pp = ((size_t)n * (p.g.H >> 1) + (py >> 1)) * (size_t)(p.g.W >> 1) + (px >> 1);
                    store16(p.e.pool_hi, p.e.pool_lo, pp * p.e.pool_C + cb, m);
                  }
                }
                if (p.e.stats) {
                  const float mk = valid ? 1.f : 0.f;
#pragma unroll
                  for (int i = 0; i < 16; ++i) {
                    const float w = v[i] * mk;
                    sacc[i] += w;
                    sacc[16 + i] = fmaf(w, w, sacc[16 + i]);
                  }
                }
              }
            }
          }
          if (p.e.stats) {
            // InstanceNorm statistics of what was stored: warp butterfly, then the four lane-quarter
            // warps of this half combine through shared memory -> 32 atomics per (tile, group)
            const float tot = warp_reduce32x32(sacc, lane);
            float* sb = stat_buf + ((half * 2 + (stat_flip & 1)) * 4) * 32;
            sb[quarter * 32 + lane] = tot;
            asm volatile("bar.sync %0, 128;" ::"r"(1 + half) : "memory");
            if (quarter == 0) {
              const float t4 = sb[lane] + sb[32 + lane] + sb[64 + lane] + sb[96 + lane];
              const int ch = cb + (lane & 15);
              atomicAdd(p.e.stats + ((size_t)n * p.e.stats_C + p.e.stats_coff + ch) * 2 + (lane >> 4), (double)t4);
            }
            ++stat_flip;
          }
        }
      } else if (tail2) {
        // fused MSBlock tail (bdcn_new.py:49-55 + conv*_down/score_dsn* collapsed), wide layout: group g
        // holds hi*hi + lo*hi in columns [64 g, 64 g + 32) and hi*lo in [64 g + 32, 64 g + 64) of every sub-tile.
        // The two warps of a lane quarter split the 32 channels; the upper half hands its two partial dots
        // to the lower one through shared memory (double-buffered by tile parity).
        const int c0 = half << 4;
        for (int s = 0; s < nsub; ++s) {
          bool valid;
          const size_t pix = tail_pix(s, valid);
          float v[16];
          float2 sc = sc_pre;
          if (s == 0) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = o_pre[i];
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = 0.f;
            sc = make_float2(0.f, 0.f);
            if (valid) {
              load16(p.e.o_hi, p.e.o_lo, pix * p.e.o_C + p.e.o_coff + c0, v);
              if (half == 0 && p.e.score_accum) sc = reinterpret_cast<const float2*>(p.e.score)[pix];
            }
          }
          const uint32_t tb = tbase + (uint32_t)s * (uint32_t)(p.g.groups * 64);
#pragma unroll
          for (int g = 0; g < 3; ++g) {
            uint32_t r[16], q[16];
            tmem_ld16(tb + g * 64 + c0, r);
            tmem_ld16(tb + g * 64 + 32 + c0, q);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i)
              v[i] += fmaxf((__uint_as_float(r[i]) + __uint_as_float(q[i])) + c_bias[g * p.g.cout_pad + c0 + i], 0.f);
          }
          float s0 = 0.f, s1 = 0.f;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            s0 = fmaf(v[i], c_scale[c0 + i], s0);
            s1 = fmaf(v[i], c_scale[32 + c0 + i], s1);
          }
          float2* xb = reinterpret_cast<float2*>(stat_buf) + ((stat_flip & 1) * 4 + quarter) * 32 + lane;
          if (half == 1) *xb = make_float2(s0, s1);
          asm volatile("bar.sync %0, 64;" ::"r"(4 + quarter) : "memory");
          if (half == 0 && valid) {
            const float2 o2 = *xb;
            sc.x += s0 + o2.x;
            sc.y += s1 + o2.y;
            reinterpret_cast<float2*>(p.e.score)[pix] = sc;
          }
          ++stat_flip;
        }
      } else {
        // fused MSBlock tail (bdcn_new.py:49-55 + conv*_down/score_dsn* collapsed)
        for (int s = half; s < nsub; s += 2) {
          const int py = ty * p.tr + s * p.sr + mr;
          const bool valid = (py < p.g.H) && (px < p.g.W);
          const size_t pix = ((size_t)n * p.g.H + py) * p.g.W + px;
          float s0 = 0.f, s1 = 0.f;
#pragma unroll
          for (int c0 = 0; c0 < 32; c0 += 16) {
            uint32_t r0[16], r1[16], r2[16];
            tmem_ld16(tbase + (s * 3 + 0) * 32 + c0, r0);
            tmem_ld16(tbase + (s * 3 + 1) * 32 + c0, r1);
            tmem_ld16(tbase + (s * 3 + 2) * 32 + c0, r2);
            float o[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = 0.f;
            if (valid) load16(p.e.o_hi, p.e.o_lo, pix * p.e.o_C + p.e.o_coff + c0, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int ch = c0 + i;
              float v = o[i];
              v += fmaxf(__uint_as_float(r0[i]) + c_bias[ch], 0.f);
              v += fmaxf(__uint_as_float(r1[i]) + c_bias[p.g.cout_pad + ch], 0.f);
              v += fmaxf(__uint_as_float(r2[i]) + c_bias[2 * p.g.cout_pad + ch], 0.f);
              s0 = fmaf(v, c_scale[ch], s0);
              s1 = fmaf(v, c_scale[32 + ch], s1);
            }
          }
          if (valid) {
            float2* dst = reinterpret_cast<float2*>(p.e.score) + pix;
            float2 cur = p.e.score_accum ? *dst : make_float2(0.f, 0.f);
            cur.x += s0;
            cur.y += s1;
            *dst = cur;
          }
        }
      }
      fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(ab));
      if (p.timing) e_busy += clock64() - et0;
    }
    if (p.timing && warp == 2 && lane == 0) {
      long long* o = p.timing + (size_t)blockIdx.x * 14;
      o[12] = e_wait; o[13] = e_busy;
    }
  }

  // ---------------------------------------------------------------------- teardown
  fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u)
                 : "memory");
  }
}

// ------------------------------------------------------------------------------ host side

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CUDA_OK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
    EGN_CHECK(p != nullptr && qres == cudaDriverEntryPointSuccess,
              "cuTensorMapEncodeTiled not available from the driver");
    fn = (PFN_encodeTiled)p;
  }
  return fn;
}

// 4-D map over an NHWC bf16 plane: dims (C, W, H, N), box (32, bw, box_rows, 1), 64-B swizzle.
// promo64: the layer reads a channel window of this buffer that does not start/end on 128-byte
// boundaries, so 128-byte L2 promotion would pull the neighbouring (unread) channels from HBM
// (measured on enc.down_block1.conv21: 39 MB/frame of DRAM reads for 19.7 MB of operands).
static void make_act_map(CUtensorMap* map, const bf16* ptr, int N, int H, int W, int C, int box_w, int box_rows,
                         bool promo64 = false) {
  EGN_CHECK(C % 8 == 0, "activation channels must be a multiple of 8");
  EGN_CHECK(box_rows >= 1 && box_rows <= 256, "activation box rows out of range");
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {EGN_KC, (cuuint32_t)box_w, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  // tuning knob: L2 promotion of the activation boxes (0 none, 1 64 B, 2 128 B, 3 256 B)
  static const int promo_env = getenv("EGN_TC_L2PROMO") ? atoi(getenv("EGN_TC_L2PROMO")) : -1;
  const int promo_sel = promo_env >= 0 ? promo_env : (promo64 ? 1 : 2);
  const CUtensorMapL2promotion promo = promo_sel == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE
                                       : promo_sel == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                                       : promo_sel == 3 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B
                                                        : CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
  CUresult r = get_encode_fn()(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)ptr, dims, strides,
                               box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                               promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  EGN_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(activation) failed: " + std::to_string((int)r));
}

// 5-D map over the d x d polyphase lattice of an NHWC bf16 plane: pixel (y, x) = (yq * d + yp, xq * d + xp), so
// address = n*H*W*C + yq*(d*W*C) + yp*(W*C) + xq*(d*C) + (xp*C + c): dims (d*C, W/d, H/d, d, N), box (32, bw, rows, 1, 1).
// A box is then a dense window of ONE phase; out-of-bounds lattice coordinates zero-fill like the padding of the
// dilated convolution.  Needs H % d == 0 and W % d == 0.
static void make_act_map_phase(CUtensorMap* map, const bf16* ptr, int N, int H, int W, int C, int d, int box_w, int box_rows,
                               bool promo64 = false) {
  EGN_CHECK(H % d == 0 && W % d == 0, "phase lattice needs H and W divisible by the phase");
  cuuint64_t dims[5] = {(cuuint64_t)d * C, (cuuint64_t)(W / d), (cuuint64_t)(H / d), (cuuint64_t)d, (cuuint64_t)N};
  cuuint64_t strides[4] = {(cuuint64_t)d * C * 2, (cuuint64_t)d * W * C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[5] = {EGN_KC, (cuuint32_t)box_w, (cuuint32_t)box_rows, 1, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = get_encode_fn()(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, (void*)ptr, dims, strides, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                               promo64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  EGN_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(phase lattice) failed: " + std::to_string((int)r));
}

// 2-D map over the packed weights [rows = ntaps*cout_pad][kpad], box (32, n_tile).
static void make_w_map(CUtensorMap* map, const bf16* ptr, int rows, int kpad, int n_tile) {
  cuuint64_t dims[2] = {(cuuint64_t)kpad, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)kpad * 2};
  cuuint32_t box[2] = {EGN_KC, (cuuint32_t)n_tile};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = get_encode_fn()(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)ptr, dims, strides,
                               box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  EGN_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(weights) failed: " + std::to_string((int)r));
}

// 3-D map over per-frame packed weights [frame][rows = ntaps*cout_pad][kpad], box (32, n_tile, 1).
static void make_w_map_frames(CUtensorMap* map, const bf16* ptr, int frames, int rows, int kpad, int n_tile) {
  cuuint64_t dims[3] = {(cuuint64_t)kpad, (cuuint64_t)rows, (cuuint64_t)frames};
  cuuint64_t strides[2] = {(cuuint64_t)kpad * 2, (cuuint64_t)rows * kpad * 2};
  cuuint32_t box[3] = {EGN_KC, (cuuint32_t)n_tile, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = get_encode_fn()(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (void*)ptr, dims, strides,
                               box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  EGN_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(per-frame weights) failed: " + std::to_string((int)r));
}

static size_t tc_smem_bytes(const TcParams& p) {
  const int nplanes = p.nsplit == 1 ? 1 : 2;
  return 1024 + (size_t)p.na * nplanes * p.a_plane_bytes + (size_t)p.nw * p.w_slot_taps * nplanes * p.w_plane_bytes +
         8 * (2 * p.na + 2 * p.nw + 4) + 16 + (2 * 2 * 4 * 32 + 768 + 512 + 512) * sizeof(float) + TC_SCHED_BYTES +
         (p.e.up_hi ? 2 * (size_t)p.up_patch_bytes + 16 : 0);
}

// Fills the tiling fields of `p` from the geometry (taps must be sorted by dx) and sizes the rings.
static void tc_configure(TcParams& p, int cout_pad, int nsplit) {
  const ConvGeom& g = p.g;
  p.nsplit = nsplit;
  p.n_blocks = ceil_div(cout_pad, 256);
  // wide layers run as 256-pixel x 128-channel CTA tiles (two sub-tiles, both TMEM buffers): the
  // weight stream per tile halves against 128 x 256 and a run of three taps fits one weight slot
  if (cout_pad >= 256 && cout_pad % 128 == 0 && !getenv("EGN_TC_N256")) p.n_blocks = cout_pad / 128;
  // An MMA into the accumulator the PREVIOUS MMA wrote costs 1.5-1.7x (tools/mma_align_probe.cu: 61 / 77 / 109 cycles at
  // N = 32 / 64 / 128 with one accumulator, 49 / 60 / 81 with two alternating, the 40 / 48 / 64-cycle bound from three
  // on).  EXPERIMENT, off by default (EGN_TC_ACC3=1): layers with a lot of MMA work per tile (K chunks x taps >= 36)
  // take THREE OR FOUR interleaved sub-tile accumulators in ONE TMEM buffer instead of one or two with a second
  // buffer (288 = 256 + 32 merged channels split 3 x 96 for that).  Measured, same box: every affected layer got
  // SLOWER (conv2_2+msblock 18.6 -> 20.0 us/frame, conv3_1 6.8 -> 7.8, conv3_x+msblock 15.3 -> 17.5, dec.up_block3
  // 3.3 -> 3.7): the epilogue of a 384-512-pixel x 96-160-channel tile (TMEM reads at 64 B/cycle plus its stores)
  // and the drain / refill around it cost more than the shorter MMAs save.
  static const bool acc3 = getenv("EGN_TC_ACC3") && atoi(getenv("EGN_TC_ACC3")) != 0;
  const bool heavy = acc3 && g.groups == 1 && !g.phase && g.nchunks * g.ntaps >= 36;
  if (heavy && cout_pad == 288) p.n_blocks = 3;
  EGN_CHECK(cout_pad % (16 * p.n_blocks) == 0, "cout_pad must split into equal 16-aligned N tiles");
  p.n_tile = cout_pad / p.n_blocks;
  EGN_CHECK(p.n_tile % 16 == 0 && p.n_tile <= 256, "bad n_tile");
  // shared activation box (3x3 layers, dilation <= 2): one (8 + 2*dxmax)-pixel-wide box per K chunk
  // serves every tap; needs 16-row x 8-pixel sub-tiles so that the 8-pixel groups are uniformly strided
  int dxmax = 0, dymax = 0, ndx = 0;
  {
    int seen[64]; int ns = 0;
    for (int t = 0; t < g.ntaps; ++t) {
      dxmax = std::max(dxmax, std::abs((int)g.tap_dx[t])); dymax = std::max(dymax, std::abs((int)g.tap_dy[t]));
      bool f = false;
      for (int k = 0; k < ns; ++k) f |= seen[k] == g.tap_dx[t];
      if (!f && ns < 64) seen[ns++] = g.tap_dx[t];
    }
    ndx = ns;
  }
  p.xshare = (ndx > 1 && dxmax <= 2 && dymax <= 2 && g.groups == 1 && !getenv("EGN_TC_NO_XSHARE")) ? 1 : 0;
  // MSBlock tail on the phase lattice: dilations 4/8/12 are 1/2/3 there and ONE (8 + 6) x (16 + 6) box serves all 27 taps
  if (g.phase) { EGN_CHECK(dxmax <= 3 && dymax <= 3, "phase lattice: tap offsets beyond 3"); p.xshare = 1; }
  p.dxmax = dxmax;
  // sub-tile shape: 8 rows x 16 px or 16 rows x 8 px, whichever covers the frame with less padding
  auto padded = [&](int bw) { const int sr = 128 / bw; return (long long)round_up(g.W, bw) * round_up(g.H, sr); };
  if (padded(8) > padded(16) && !g.phase) p.xshare = 0;      // 120x160: 16-row sub-tiles would waste 6.7 % of the MMAs
  const int bw = p.xshare ? 8 : (padded(8) < padded(16) ? 8 : 16);
  p.bw_log2 = bw == 8 ? 3 : 4;
  p.sr = 128 / bw;
  // latency mode (streaming micro-batches of a few frames, evaluate.py's per-image path): when the
  // planned batch cannot fill the SMs even with single sub-tiles, split N further so that more CTAs
  // share the layer (each re-reads the activation boxes, which is free while SMs would idle)
  const int sm_target = 148;
  auto tiles_at = [&](int S, int n_blocks) {
    return (long long)ceil_div(g.W, bw) * ceil_div(g.H, S * p.sr) * g.batch * n_blocks;
  };
  const bool latency_mode = !getenv("EGN_TC_NO_LATENCY_MODE");
  while (latency_mode && tiles_at(1, p.n_blocks) * 2 <= sm_target && p.n_tile >= 64 && p.n_tile % 32 == 0) {
    p.n_blocks *= 2;
    p.n_tile /= 2;
  }
  // narrow single-group layers are bound by the shared-memory read of the activation tile (an
  // M=128, N<=64 MMA takes (128+N)/4 cycles, tools/mma_probe.cu): fold hi*hi and hi*lo into one MMA
  static const int wide_max = getenv("EGN_TC_WIDE_MAX") ? atoi(getenv("EGN_TC_WIDE_MAX")) : 64;   // tuning knob
  p.wide_b = (nsplit == 3 && g.groups == 1 && p.n_tile <= wide_max && 2 * p.n_tile <= 256) ? 1 : 0;
  if (const char* e = getenv("EGN_TC_WIDE")) p.wide_b = atoi(e) ? p.wide_b : 0;
  if (p.prod_mode == 2) p.wide_b = 0;            // the wide MMA cannot leave out hi*lo
  if (g.phase) { EGN_CHECK(nsplit == 3 && g.groups == 3 && p.n_tile == 32, "phase lattice is the MSBlock tail's layout"); p.wide_b = 1; }
  p.tap_triples = 0;
  if (g.phase && g.ntaps % 3 == 0 && !getenv("EGN_NO_TRIPLES") && p.prod_mode == 0) {
    p.tap_triples = 1;
    for (int t = 0; t < g.ntaps; ++t) if (g.tap_grp[t] != t % 3) p.tap_triples = 0;
  }
  const int cols = g.groups * p.n_tile * (p.wide_b ? 2 : 1);
  EGN_CHECK(cols <= 512, "accumulator groups exceed TMEM");
  // more sub-tiles per CTA tile amortise the weight stream and give the MMA pipe independent
  // accumulators to interleave; prefer shapes that keep two TMEM buffers
  // (measured: a second sub-tile that costs the double buffer loses ~15 % on N = 256 layers)
  p.S = 1;
  const int rows_avail = ceil_div(g.H, p.sr);
  if (rows_avail >= 2 && 2 * cols <= TC_ACC_STRIDE) p.S = 2;
  if (rows_avail >= 4 && 4 * cols <= TC_ACC_STRIDE) p.S = 4;
  if (getenv("EGN_TC_S4SINGLE") && p.wide_b && rows_avail >= 4 && 4 * cols <= 512) p.S = 4;           // tuning knob: one TMEM buffer
  if (heavy && !p.wide_b && p.S < 3 && 3 * cols > TC_ACC_STRIDE) {
    // one TMEM buffer, three or four accumulators (only where two buffers cannot hold three of them)
    const int s1 = std::min(std::min(4, 512 / cols), rows_avail);
    const int rows_last = rows_avail % s1;            // sub-tiles of the last (partial) tile row
    if (s1 >= 3 && (rows_last == 0 || rows_last >= 3 || rows_avail / s1 >= 4)) p.S = s1;
  }
  if (const char* e = getenv("EGN_TC_SMAX")) p.S = std::min(p.S, std::max(1, atoi(e)));   // tuning knob
  while (latency_mode && p.S > 1 && tiles_at(p.S, p.n_blocks) < sm_target) p.S /= 2;      // fewer sub-tiles, more CTAs
  p.tr = p.S * p.sr;
  p.acc_bufs = p.S * cols <= TC_ACC_STRIDE ? 2 : 1;
  p.tiles_x = ceil_div(g.W, bw);
  p.tiles_y = ceil_div(g.H, p.tr);
  p.total_tiles = p.tiles_x * p.tiles_y * g.batch * p.n_blocks;
  // A boxes: one per distinct dx, covering every dy (or a single wide one, see xshare)
  p.dmax = 0;
  p.n_aloads = 0;
  for (int t = 0; t < g.ntaps; ++t) {
    p.dmax = std::max(p.dmax, std::abs((int)g.tap_dy[t]));
    if (p.xshare) {
      if (p.n_aloads == 0) { p.aload_dx[0] = (int8_t)(-dxmax); p.aload_tap0[0] = 0; p.aload_ntaps[0] = 0; p.n_aloads = 1; }
      ++p.aload_ntaps[0];
      continue;
    }
    if (p.n_aloads == 0 || p.aload_dx[p.n_aloads - 1] != g.tap_dx[t]) {
      for (int l = 0; l < p.n_aloads; ++l) EGN_CHECK(p.aload_dx[l] != g.tap_dx[t], "taps must be grouped by dx");
      EGN_CHECK(p.n_aloads < TC_MAX_ALOADS, "too many distinct horizontal tap offsets");
      p.aload_dx[p.n_aloads] = g.tap_dx[t];
      p.aload_tap0[p.n_aloads] = (uint8_t)t;
      p.aload_ntaps[p.n_aloads] = 0;
      ++p.n_aloads;
    }
    ++p.aload_ntaps[p.n_aloads - 1];
  }
  p.box_rows = p.tr + 2 * p.dmax;
  p.l2_prefetch = p.n_blocks == 1 ? 1 : 0;
  p.l2_prefetch = 0;
  if (const char* e = getenv("EGN_TC_PREFETCH")) p.l2_prefetch = atoi(e);
  p.dbg = 0;
  if (const char* e = getenv("EGN_TC_DBG")) p.dbg = atoi(e);
  p.box_w = p.xshare ? bw + 2 * dxmax : bw;
  p.a_box_bytes = (uint32_t)p.box_rows * p.box_w * 64u;
  p.a_plane_bytes = (p.a_box_bytes + 1023u) & ~1023u;
  p.w_plane_bytes = (uint32_t)p.n_tile * 64u;
  const int nplanes = nsplit == 1 ? 1 : 2;
  // upsample-add layers stage two half-resolution patches next to the operand rings (pixel stride padded by
  // 16 bytes so that the 16-byte reads of neighbouring pixels fall into different banks)
  p.up_rows = p.up_cols = 0; p.up_pix_stride = p.up_plane_stride = p.up_patch_bytes = 0;
  if (p.e.up_hi) {
    p.up_rows = p.tr / 2 + 2; p.up_cols = (1 << p.bw_log2) / 2 + 2;
    p.up_pix_stride = (uint32_t)p.n_tile * 4u + 16u;           // 8 channels = hi 16 B | lo 16 B, converted in place to 8 fp32
    p.up_plane_stride = 0;
    p.up_patch_bytes = (uint32_t)(p.up_rows * p.up_cols) * p.up_pix_stride;
    EGN_CHECK(p.up_rows * p.up_cols * (p.n_tile / 8) <= TC_UP_PAIRS * TC_EPI_WARPS * 32, "upsample patch exceeds the staging slots");
  }
  const size_t budget = 227 * 1024 - 1024 - 512 - 2048 - 7168 - TC_SCHED_BYTES - (p.e.up_hi ? 2 * (size_t)p.up_patch_bytes + 16 : 0);
  const size_t a_slot = (size_t)nplanes * p.a_plane_bytes, w_slot = (size_t)nplanes * p.w_plane_bytes;
  // small N: an MMA is short (40-48 cycles) and the issue queue holds only ~8 of them, so the
  // per-tap barrier round trips of the weight ring would starve the pipe; load every tap of a box
  // into one slot instead (measured with tools/mma_queue_probe.cu)
  int max_box_taps = 1;
  for (int l = 0; l < p.n_aloads; ++l) max_box_taps = std::max(max_box_taps, (int)p.aload_ntaps[l]);
  p.w_box = 1;
  if (const char* e = getenv("EGN_TC_WBOX")) p.w_box = atoi(e) ? p.w_box : 0;
  p.w_slot_taps = p.w_box ? std::min(max_box_taps, 3) : 1;
  if (budget < 2 * a_slot + 2 * w_slot * p.w_slot_taps) { p.w_box = 0; p.w_slot_taps = 1; }   // 256-wide tiles: one tap per slot
  // small layers (32 -> 32 3x3: 36 KB, the full-resolution 1x1 layers: 8-20 KB): keep every weight tile
  // resident instead of re-streaming it per pixel tile - frees the weight barriers and ~30 % of the
  // TMA writes into the shared memory the MMAs read their operands from
  p.w_res = 0;
  {
    const size_t all_w = w_slot * (size_t)g.ntaps * g.nchunks;
    if (p.n_blocks == 1 && !p.w_frames && (all_w <= 48 * 1024 || g.phase) && budget >= 2 * a_slot + all_w && !getenv("EGN_TC_NO_WRES")) {
      p.w_res = 1; p.w_box = 1; p.w_slot_taps = g.ntaps * g.nchunks;
    }
  }
  const size_t w_slot_all = w_slot * p.w_slot_taps;
  p.na = 2;
  EGN_CHECK(budget >= p.na * a_slot + (p.w_res ? 1 : 2) * w_slot_all, "conv_tc: tile does not fit in shared memory");
  p.nw = p.w_res ? 1 : (int)std::min<size_t>(p.w_box ? 3 : 8, (budget - p.na * a_slot) / w_slot_all);
  while (p.na < 4 && budget >= (p.na + 1) * a_slot + (size_t)p.nw * w_slot_all) ++p.na;
}

// The dynamic shared-memory opt-in is a per-DEVICE function attribute: egn_create calls this with the
// context's device current, so every device that owns a context has it (a process-wide flag would leave
// the second GPU of a single process without it).
static void tc_prepare_device() {
  CUDA_OK(cudaFuncSetAttribute(conv_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  CUDA_OK(cudaFuncSetAttribute(conv_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
}

static void tc_launch(const TcParams& p, int num_sms, cudaStream_t stream) {
  const size_t smem = tc_smem_bytes(p);
  EGN_CHECK(smem <= 227 * 1024, "conv_tc smem budget exceeded");
  int grid = p.total_tiles < num_sms ? p.total_tiles : num_sms;
  static const bool timing = getenv("EGN_TC_TIMING") != nullptr;
  if (timing) {
    // debugging aid: per-role cycle counters of one launch, averaged over CTAs, to stderr
    static long long* buf = nullptr;
    if (!buf) CUDA_OK(cudaMalloc(&buf, 148 * 14 * sizeof(long long)));
    CUDA_OK(cudaMemsetAsync(buf, 0, 148 * 14 * sizeof(long long), stream));
    TcParams q = p;
    q.timing = buf;
    if (q.e.up_hi) conv_tc_kernel<true><<<grid, TC_THREADS, smem, stream>>>(q);
    else conv_tc_kernel<false><<<grid, TC_THREADS, smem, stream>>>(q);
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaStreamSynchronize(stream));
    long long h[148 * 14];
    CUDA_OK(cudaMemcpy(h, buf, sizeof(h), cudaMemcpyDeviceToHost));
    double a[14] = {0};
    for (int b = 0; b < grid; ++b) for (int k = 0; k < 14; ++k) a[k] += (double)h[b * 14 + k] / grid;
    fprintf(stderr, "tc_timing H=%d W=%d chunks=%d taps=%d ntile=%d S=%d na=%d nw=%d batch=%d | mma: total %.0f tempty %.0f a_full %.0f w_full %.0f issue %.0f commit %.0f tfullcommit %.0f head %.0f tiles %.1f | prod: a_empty %.0f w_empty %.0f total %.0f | epi(warp 2): wait %.0f busy %.0f\n",
            p.g.H, p.g.W, p.g.nchunks, p.g.ntaps, p.n_tile, p.S, p.na, p.nw, p.g.batch, a[0], a[1], a[2], a[3], a[4], a[9], a[10], a[11], a[5], a[6], a[7], a[8], a[12], a[13]);
    return;
  }
  if (p.e.up_hi) conv_tc_kernel<true><<<grid, TC_THREADS, smem, stream>>>(p);
  else conv_tc_kernel<false><<<grid, TC_THREADS, smem, stream>>>(p);
  CUDA_OK(cudaGetLastError());
}
