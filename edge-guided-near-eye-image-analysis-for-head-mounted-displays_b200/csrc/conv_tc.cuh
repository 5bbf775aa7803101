// Implicit-GEMM convolution on tcgen05 tensor cores (sm_100a).
//
//   out[n, y, x, :] = epilogue( sum_taps sum_k  A[n, y+dy, x+dx, k] * Wt[tap][:, k] )
//
// * GEMM view: M = 128 output pixels (an 8-row x 16-column spatial tile of one frame),
//   N = n_tile output channels (<= 256), K = 64-channel chunks x taps.
// * A tiles come straight from the NHWC activation planes with one 4-D TMA box
//   {64 ch, 16 px, 8 rows, 1 frame} per (tap, chunk), shifted by the tap offset; TMA out-of-bounds
//   zero fill implements the convolution padding for every dilation (1/2/4/8/12), and a chunk table
//   (source buffer, channel offset, frame offset) implements channel concatenation without copies.
// * Weight tiles {64 k, n_tile rows} come from a 2-D map over the repacked [tap][cout_pad][kpad]
//   matrix.  Both land in shared memory in the 128-byte swizzled K-major layout UMMA consumes.
// * Precision: nsplit==3 issues hi*hi + lo*hi + hi*lo bf16 MMAs into one fp32 TMEM accumulator.
// * Warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2-5 = epilogue
//   (tcgen05.ld -> bias/activation/affine -> split-bf16 NHWC stores, or the fused MSBlock tail).
//   smem ring (full/empty mbarriers) between producer and MMA; two TMEM accumulator buffers
//   (tmem_full/tmem_empty) between MMA and epilogue; persistent CTAs stride over the tile list.
#pragma once
#include "common.cuh"

struct TcParams {
  CUtensorMap a_map[2][EGN_MAX_SRC];  // [plane hi/lo][source]
  CUtensorMap w_map[2];               // [plane hi/lo]
  ConvGeom g;
  ConvEpi e;
  int nsplit;                          // 1: hi*hi only, 3: split product
  int n_tile, n_blocks, tiles_x, tiles_y, total_tiles, stages;
  int* err_flag;
};

#define TC_THREADS 192
#define TC_A_BYTES 16384
#define TC_ACC_STRIDE 256   // TMEM columns between the two accumulator buffers

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
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int* err_flag, int code) {
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

__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint32_t bar, uint32_t dst,
                                            int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}

__device__ __forceinline__ void prefetch_map(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// K-major, 128-byte swizzle: rows of 128 B, 8-row atoms 1024 B apart (SBO), version 1 (sm_100).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;                 // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;       // stride byte offset
  d |= (uint64_t)1 << 46;                 // descriptor version
  d |= (uint64_t)2 << 61;                 // SWIZZLE_128B
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

}  // namespace tc

__global__ void __launch_bounds__(TC_THREADS, 1) conv_tc_kernel(const __grid_constant__ TcParams p) {
  using namespace tc;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;      // SWIZZLE_128B tiles need 1024-B alignment
  uint8_t* smem = smem_raw + (base - raw_addr);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nplanes = p.nsplit == 1 ? 1 : 2;
  const uint32_t w_bytes = (uint32_t)p.n_tile * 128u;
  const uint32_t stage_bytes = nplanes * (TC_A_BYTES + w_bytes);
  const int stages = p.stages;

  const uint32_t bar_base = base + stages * stage_bytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (stages + s); };
  auto tfull_bar = [&](int b) { return bar_base + 8u * (2 * stages + b); };
  auto tempty_bar = [&](int b) { return bar_base + 8u * (2 * stages + 2 + b); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + stages * stage_bytes + 8 * (2 * stages + 4));

  if (warp == 0 && lane == 0) {
    prefetch_map(&p.w_map[0]);
    prefetch_map(&p.a_map[0][0]);
    for (int s = 0; s < stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull_bar(b), 1);
      mbar_init(tempty_bar(b), 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int nb = tile % p.n_blocks;
        int rest = tile / p.n_blocks;
        const int tx = rest % p.tiles_x;
        rest /= p.tiles_x;
        const int ty = rest % p.tiles_y;
        const int n = rest / p.tiles_y;
        for (int t = 0; t < p.g.ntaps; ++t) {
          const int x = tx * 16 + p.g.tap_dx[t];
          const int y = ty * 8 + p.g.tap_dy[t];
          const int wrow = t * p.g.cout_pad + nb * p.n_tile;
          for (int c = 0; c < p.g.nchunks; ++c) {
            mbar_wait(empty_bar(stage), phase ^ 1u, p.err_flag, 1);
            mbar_expect_tx(full_bar(stage), stage_bytes);
            const uint32_t sA = base + stage * stage_bytes;
            const uint32_t sW = sA + nplanes * TC_A_BYTES;
            const int src = p.g.chunk_src[c];
            const int c0 = p.g.chunk_c0[c];
            const int nn = n + p.g.chunk_noff[c];
            tma_load_4d(&p.a_map[0][src], full_bar(stage), sA, c0, x, y, nn);
            tma_load_2d(&p.w_map[0], full_bar(stage), sW, c * EGN_KC, wrow);
            if (nplanes == 2) {
              tma_load_4d(&p.a_map[1][src], full_bar(stage), sA + TC_A_BYTES, c0, x, y, nn);
              tma_load_2d(&p.w_map[1], full_bar(stage), sW + w_bytes, c * EGN_KC, wrow);
            }
            if (++stage == stages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) |
                             ((uint32_t)(p.n_tile >> 3) << 17) | ((128u >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0;
      int ab = 0;
      uint32_t aphase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        mbar_wait(tempty_bar(ab), aphase ^ 1u, p.err_flag, 2);
        fence_after();
        uint32_t started = 0;
        for (int t = 0; t < p.g.ntaps; ++t) {
          const int grp = p.g.tap_grp[t];
          const uint32_t d_tmem = tmem_base + ab * TC_ACC_STRIDE + grp * p.n_tile;
          for (int c = 0; c < p.g.nchunks; ++c) {
            mbar_wait(full_bar(stage), phase, p.err_flag, 3);
            fence_after();
            const uint32_t sA = base + stage * stage_bytes;
            const uint32_t sW = sA + nplanes * TC_A_BYTES;
            const uint64_t dA_hi = make_desc(sA);
            const uint64_t dW_hi = make_desc(sW);
            const uint64_t dA_lo = make_desc(sA + TC_A_BYTES);
            const uint64_t dW_lo = make_desc(sW + w_bytes);
#pragma unroll
            for (int k = 0; k < EGN_KC / 16; ++k) {
              const uint64_t koff = (uint64_t)((k * 32) >> 4);   // 16 bf16 = 32 bytes along K
              const uint32_t acc = ((started >> grp) & 1u) | (k > 0 ? 1u : 0u);
              mma_bf16(d_tmem, dA_hi + koff, dW_hi + koff, idesc, acc);
              if (nplanes == 2) {
                mma_bf16(d_tmem, dA_lo + koff, dW_hi + koff, idesc, 1u);
                mma_bf16(d_tmem, dA_hi + koff, dW_lo + koff, idesc, 1u);
              }
            }
            started |= 1u << grp;
            mma_commit(empty_bar(stage));          // frees the smem slot when these MMAs retire
            if (++stage == stages) { stage = 0; phase ^= 1u; }
          }
        }
        mma_commit(tfull_bar(ab));                 // accumulator complete -> epilogue
        ab ^= 1;
        if (ab == 0) aphase ^= 1u;
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..5)
    const int quarter = warp & 3;                  // TMEM lane quarter this warp may read
    const int m = quarter * 32 + lane;             // accumulator row = pixel within the tile
    int ab = 0;
    uint32_t aphase = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int nb = tile % p.n_blocks;
      int rest = tile / p.n_blocks;
      const int tx = rest % p.tiles_x;
      rest /= p.tiles_x;
      const int ty = rest % p.tiles_y;
      const int n = rest / p.tiles_y;
      const int py = ty * 8 + (m >> 4);
      const int px = tx * 16 + (m & 15);
      const bool valid = (py < p.g.H) && (px < p.g.W);
      const size_t pix = ((size_t)n * p.g.H + py) * p.g.W + px;

      mbar_wait(tfull_bar(ab), aphase, p.err_flag, 4);
      fence_after();
      const uint32_t tbase = tmem_base + ((uint32_t)(quarter * 32) << 16) + ab * TC_ACC_STRIDE;

      if (p.e.mode == CONV_STORE) {
        for (int c0 = 0; c0 < p.n_tile; c0 += 16) {
          uint32_t r[16];
          tmem_ld16(tbase + c0, r);
          tmem_ld_wait();
          const int cb = nb * p.n_tile + c0;      // first output channel of this 16-column group
          if (valid) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int ch = cb + h * 8;
              if (ch + 8 <= p.e.cout_store) {
                float v[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  float a = __uint_as_float(r[h * 8 + i]) + __ldg(p.e.bias + ch + i);
                  a = apply_act(a, p.e.act);
                  if (p.e.post_scale) a = a * __ldg(p.e.post_scale + ch + i) + __ldg(p.e.post_shift + ch + i);
                  v[i] = a;
                }
                store8(p.e.out_hi, p.e.out_lo, pix * p.e.out_C + p.e.out_coff + ch, v);
              }
            }
          }
        }
      } else {
        // fused MSBlock tail (bdcn_new.py:49-55 + conv*_down/score_dsn* collapsed)
        float s0 = 0.f, s1 = 0.f;
        for (int c0 = 0; c0 < 32; c0 += 16) {
          uint32_t r0[16], r1[16], r2[16];
          tmem_ld16(tbase + c0, r0);
          tmem_ld16(tbase + 32 + c0, r1);
          tmem_ld16(tbase + 64 + c0, r2);
          tmem_ld_wait();
          if (valid) {
            float o[16];
            load8(p.e.o_hi, p.e.o_lo, pix * 32 + c0, o);
            load8(p.e.o_hi, p.e.o_lo, pix * 32 + c0 + 8, o + 8);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int ch = c0 + i;
              float v = o[i];
              v += fmaxf(__uint_as_float(r0[i]) + __ldg(p.e.bias + ch), 0.f);
              v += fmaxf(__uint_as_float(r1[i]) + __ldg(p.e.bias + p.g.cout_pad + ch), 0.f);
              v += fmaxf(__uint_as_float(r2[i]) + __ldg(p.e.bias + 2 * p.g.cout_pad + ch), 0.f);
              s0 += v * __ldg(p.e.score_w + ch);
              s1 += v * __ldg(p.e.score_w + 32 + ch);
            }
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
      fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(ab));
      ab ^= 1;
      if (ab == 0) aphase ^= 1u;
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

// 4-D map over an NHWC bf16 plane: dims (C, W, H, N), box (64, 16, 8, 1), 128-B swizzle.
static void make_act_map(CUtensorMap* map, const bf16* ptr, int N, int H, int W, int C) {
  EGN_CHECK(C % 8 == 0, "activation channels must be a multiple of 8");
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {EGN_KC, 16, 8, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = get_encode_fn()(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)ptr, dims, strides,
                               box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  EGN_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(activation) failed: " + std::to_string((int)r));
}

// 2-D map over the packed weights [rows = ntaps*cout_pad][kpad], box (64, n_tile).
static void make_w_map(CUtensorMap* map, const bf16* ptr, int rows, int kpad, int n_tile) {
  cuuint64_t dims[2] = {(cuuint64_t)kpad, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)kpad * 2};
  cuuint32_t box[2] = {EGN_KC, (cuuint32_t)n_tile};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = get_encode_fn()(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)ptr, dims, strides,
                               box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  EGN_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(weights) failed: " + std::to_string((int)r));
}

static size_t tc_smem_bytes(const TcParams& p) {
  const int nplanes = p.nsplit == 1 ? 1 : 2;
  const size_t stage = (size_t)nplanes * (TC_A_BYTES + (size_t)p.n_tile * 128);
  return 1024 + p.stages * stage + 8 * (2 * p.stages + 4) + 16;
}

// Fills the tiling fields of `p` from geometry + cout and picks the pipeline depth.
static void tc_configure(TcParams& p, int cout_pad, int nsplit) {
  p.nsplit = nsplit;
  p.n_blocks = ceil_div(cout_pad, 256);
  EGN_CHECK(cout_pad % (16 * p.n_blocks) == 0, "cout_pad must split into equal 16-aligned N tiles");
  p.n_tile = cout_pad / p.n_blocks;
  EGN_CHECK(p.n_tile % 16 == 0 && p.n_tile <= 256, "bad n_tile");
  EGN_CHECK(p.g.groups * p.n_tile <= TC_ACC_STRIDE, "accumulator groups exceed a TMEM buffer");
  p.tiles_x = ceil_div(p.g.W, 16);
  p.tiles_y = ceil_div(p.g.H, 8);
  p.total_tiles = p.tiles_x * p.tiles_y * p.g.batch * p.n_blocks;
  const int nplanes = nsplit == 1 ? 1 : 2;
  const size_t stage = (size_t)nplanes * (TC_A_BYTES + (size_t)p.n_tile * 128);
  int st = (int)((200 * 1024) / stage);
  if (st > 8) st = 8;
  EGN_CHECK(st >= 2, "pipeline needs at least two stages");
  p.stages = st;
}

static void tc_launch(const TcParams& p, int num_sms, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    CUDA_OK(cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 227 * 1024));
    attr_set = true;
  }
  const size_t smem = tc_smem_bytes(p);
  EGN_CHECK(smem <= 227 * 1024, "conv_tc smem budget exceeded");
  int grid = p.total_tiles < num_sms ? p.total_tiles : num_sms;
  conv_tc_kernel<<<grid, TC_THREADS, smem, stream>>>(p);
  CUDA_OK(cudaGetLastError());
}
