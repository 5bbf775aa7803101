// CUDA-core companion of the tensor-core convolution.
//
// conv_simt_kernel consumes exactly the same ConvGeom / ConvEpi / packed split-bf16 weights as
// conv_tc_kernel but evaluates the products with fp32 FMAs from shared-memory tiles
// (64 pixels x 32 output channels per block).  It is the on-device cross-check for the tcgen05
// path (EGN_CONV=simt selects it for every layer; egn_conv_selfcheck compares the two on identical
// inputs) and is not a performance path.
#pragma once
#include "common.cuh"

struct SimtParams {
  ConvSrc src[EGN_MAX_SRC];
  const bf16* w_hi;   // [ntaps][cout_pad][kpad]
  const bf16* w_lo;
  ConvGeom g;
  ConvEpi e;
  int nsplit;
  int prod_mode;   // same meaning as TcParams::prod_mode
};

#define ST_PX 64
#define ST_CO 32
#define ST_THREADS 256

__global__ void __launch_bounds__(ST_THREADS) conv_simt_kernel(const SimtParams p) {
  // hi and lo parts are kept apart and multiplied like the tensor-core kernel does (hi*hi + lo*hi + hi*lo, the
  // lo*lo term is dropped): the cross-check then holds for ANY buffer contents, also for stale planes of a
  // shared workspace region whose "lo" is not small against its "hi"
  __shared__ float sA[ST_PX][EGN_KC + 1], sAl[ST_PX][EGN_KC + 1];
  __shared__ float sW[ST_CO][EGN_KC + 1], sWl[ST_CO][EGN_KC + 1];
  __shared__ float sRed[ST_PX][4][2];
  const int HW = p.g.H * p.g.W;
  const int tiles_per_frame = (HW + ST_PX - 1) / ST_PX;
  const int n = blockIdx.x / tiles_per_frame;
  const int p0 = (blockIdx.x % tiles_per_frame) * ST_PX;
  const int cb = blockIdx.y * ST_CO;
  const int t = threadIdx.x;
  const int px = t % ST_PX, cog = t / ST_PX;

  float acc[3][8];
#pragma unroll
  for (int gi = 0; gi < 3; ++gi)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[gi][i] = 0.f;

  for (int tap = 0; tap < p.g.ntaps; ++tap) {
    float cur[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) cur[i] = 0.f;
    for (int c = 0; c < p.g.nchunks; ++c) {
      const ConvSrc s = p.src[p.g.chunk_src[c]];
      const int c0 = p.g.chunk_c0[c];
      __syncthreads();
      for (int e = t; e < ST_PX * EGN_KC; e += ST_THREADS) {
        const int q = e / EGN_KC, k = e % EGN_KC;
        const int lin = p0 + q;
        float v = 0.f, vl = 0.f;
        if (lin < HW) {
          const int y = lin / p.g.W + p.g.tap_dy[tap], x = lin % p.g.W + p.g.tap_dx[tap];
          const int cc = c0 + k;
          if (y >= 0 && y < p.g.H && x >= 0 && x < p.g.W && cc < s.C) {   // TMA zero-fills the rest
            const size_t a = (((size_t)(n + p.g.chunk_noff[c]) * p.g.H + y) * p.g.W + x) * s.C + cc;
            v = __bfloat162float(s.hi[a]);
            if (p.nsplit != 1) vl = __bfloat162float(s.lo[a]);
          }
        }
        sA[q][k] = v; sAl[q][k] = vl;
      }
      for (int e = t; e < ST_CO * EGN_KC; e += ST_THREADS) {
        const int co = e / EGN_KC, k = e % EGN_KC;
        float v = 0.f, vl = 0.f;
        if (cb + co < p.g.cout_pad) {
          const size_t w = ((size_t)tap * p.g.cout_pad + cb + co) * p.g.kpad + c * EGN_KC + k;
          v = __bfloat162float(p.w_hi[w]);
          if (p.nsplit != 1) vl = __bfloat162float(p.w_lo[w]);
        }
        sW[co][k] = v; sWl[co][k] = vl;
      }
      __syncthreads();
#pragma unroll 8
      for (int k = 0; k < EGN_KC; ++k) {
        const float a = sA[px][k], al = sAl[px][k];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float wh = sW[cog * 8 + i][k];
          const float lh = p.prod_mode == 1 ? 0.f : al, hl = p.prod_mode == 2 ? 0.f : sWl[cog * 8 + i][k];
          cur[i] = fmaf(a, wh, fmaf(lh, wh, fmaf(a, hl, cur[i])));
        }
      }
    }
    const int grp = p.g.tap_grp[tap];
#pragma unroll
    for (int gi = 0; gi < 3; ++gi)
      if (gi == grp)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[gi][i] += cur[i];
  }

  const int lin = p0 + px;
  const bool valid = lin < HW;
  const size_t pix = (size_t)n * HW + lin;
  if (p.e.mode == CONV_STORE || p.e.mode == CONV_LOGITS) {
    const int ch = cb + cog * 8;
    if (valid && ch + 8 <= p.e.cout_store) {
      float v[8];
      float up[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (p.e.up_hi) upsample_add<8>(p.e, n, lin / p.g.W, lin % p.g.W, p.g.H, p.g.W, ch, up);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float a = acc[0][i] + p.e.bias[ch + i] + up[i];
        a = apply_act(a, p.e.act);
        if (p.e.post_scale) a = a * p.e.post_scale[ch + i] + p.e.post_shift[ch + i];
        v[i] = a;
      }
      if (p.e.mode == CONV_LOGITS) {
        if (ch == 0)
          for (int i = 0; i < p.e.logits_c && i < 8; ++i)
            p.e.logits[((size_t)n * p.e.logits_c + i) * HW + lin] = v[i];
      } else
      store8(p.e.out_hi, p.e.out_lo, pix * p.e.out_C + p.e.out_coff + ch, v);
      if (p.e.stats) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          double* st = p.e.stats + ((size_t)n * p.e.stats_C + p.e.stats_coff + ch + i) * 2;
          atomicAdd(st, (double)v[i]);
          atomicAdd(st + 1, (double)v[i] * (double)v[i]);
        }
      }
    }
  } else {
    float s0 = 0.f, s1 = 0.f;
    if (valid) {
      float o[8];
      load8(p.e.o_hi, p.e.o_lo, pix * p.e.o_C + p.e.o_coff + cog * 8, o);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int ch = cog * 8 + i;
        float v = o[i];
#pragma unroll
        for (int gi = 0; gi < 3; ++gi) v += fmaxf(acc[gi][i] + p.e.bias[gi * p.g.cout_pad + ch], 0.f);
        s0 += v * p.e.score_w[ch];
        s1 += v * p.e.score_w[32 + ch];
      }
    }
    sRed[px][cog][0] = s0;
    sRed[px][cog][1] = s1;
    __syncthreads();
    if (valid && cog == 0) {
      float2* dst = reinterpret_cast<float2*>(p.e.score) + pix;
      float2 cur2 = p.e.score_accum ? *dst : make_float2(0.f, 0.f);
      cur2.x += sRed[px][0][0] + sRed[px][1][0] + sRed[px][2][0] + sRed[px][3][0];
      cur2.y += sRed[px][0][1] + sRed[px][1][1] + sRed[px][2][1] + sRed[px][3][1];
      *dst = cur2;
    }
  }
}

static void simt_launch(const SimtParams& p, cudaStream_t stream) {
  const int HW = p.g.H * p.g.W;
  const int tiles = (HW + ST_PX - 1) / ST_PX;
  const int cper = p.e.mode == CONV_MSBLOCK ? 32 : p.e.cout_store;
  dim3 grid((unsigned)(tiles * p.g.batch), (unsigned)((cper + ST_CO - 1) / ST_CO));
  conv_simt_kernel<<<grid, ST_THREADS, 0, stream>>>(p);
  CUDA_OK(cudaGetLastError());
}
