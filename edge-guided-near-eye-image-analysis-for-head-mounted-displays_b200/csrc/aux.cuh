// Bandwidth-bound helper kernels of the edge+ESF-Net graph (all split-bf16 NHWC unless noted).
#pragma once
#include "common.cuh"

// A channel window of a split buffer.
struct View {
  bf16* hi;
  bf16* lo;
  int C;      // channels of the underlying buffer
  int coff;   // first channel of the window
  int n_off;  // first frame of the window
};

static inline View make_view(const Act& a, int coff, int n_off = 0) {
  View v;
  v.hi = a.hi; v.lo = a.lo; v.C = a.C; v.coff = coff; v.n_off = n_off;
  return v;
}

// ------------------------------------------------------------------------------------------
// First layer: 3x3, pad 1, from up to two fp32 single-channel planes [B][H][W]
// (vgg16_c.py:11 conv1_1 on cat(img,img,img) with the three input channels pre-summed;
//  utils.py:1042 convBlock.conv1 of the ESF-Net head, input_concat => two planes).
struct FirstConvParams {
  const float* in[3];   // input planes (frame 0); cin of them are used
  long long fstride[3]; // floats between consecutive frames of each plane
  int cin;
  const float* w;       // [cin][9][cout]
  const float* bias;    // [cout]
  View dst;
  int B, H, W, cout, act;
};

#define FC_TW 32   // tile width (pixels)
#define FC_TH 8    // tile height
#define FC_YT 1    // vertically consecutive tiles per block (the weights are staged once per block)

// Block = 8 x 32 output pixels.  The fp32 input window (10 x 34 per plane) and the weights are
// staged in shared memory; a thread produces 8 output channels of 4 horizontally adjacent pixels,
// so each weight read serves four pixels and each input value is read once per 8 channels.
template <int CIN, int COUT>
__global__ void __launch_bounds__(256) first_conv_kernel(const FirstConvParams p) {
  constexpr int G = COUT / 8;                 // 8-channel groups
  __shared__ __align__(16) float sw[CIN * 9 * COUT];           // [cin][tap][half][group][4]
  __shared__ float tile[CIN][FC_TH + 2][FC_TW + 2];
  const int n = blockIdx.z, x0 = blockIdx.x * FC_TW;
  for (int i = threadIdx.x; i < CIN * 9 * COUT; i += 256) {
    const int co = i % COUT, ct = i / COUT;                    // p.w is [cin][tap][cout]
    const int g = co >> 3, h = (co >> 2) & 1, k = co & 3;
    sw[((ct * 2 + h) * G + g) * 4 + k] = p.w[i];
  }
  for (int yt = 0; yt < FC_YT; ++yt) {
  const int y0 = (blockIdx.y * FC_YT + yt) * FC_TH;
  if (y0 >= p.H) break;
  if (yt) __syncthreads();                                     // everybody is done with the previous tile
#pragma unroll
  for (int ci = 0; ci < CIN; ++ci) {
    const float* in = p.in[ci] + (size_t)n * p.fstride[ci];
    for (int i = threadIdx.x; i < (FC_TH + 2) * (FC_TW + 2); i += 256) {
      const int ty = i / (FC_TW + 2), tx = i % (FC_TW + 2);
      const int yy = y0 + ty - 1, xx = x0 + tx - 1;
      tile[ci][ty][tx] = (yy >= 0 && yy < p.H && xx >= 0 && xx < p.W) ? __ldg(in + (size_t)yy * p.W + xx) : 0.f;
    }
  }
  __syncthreads();
  constexpr int RUNS = FC_TH * FC_TW / 4;      // runs of 4 pixels in the tile
  for (int item = threadIdx.x; item < RUNS * G; item += 256) {
    const int g = item % G, run = item / G;
    const int ty = run / (FC_TW / 4), tx = (run % (FC_TW / 4)) * 4;
    float acc[4][8];
    {
      const float4 b0 = *reinterpret_cast<const float4*>(p.bias + g * 8);
      const float4 b1 = *reinterpret_cast<const float4*>(p.bias + g * 8 + 4);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        acc[q][0] = b0.x; acc[q][1] = b0.y; acc[q][2] = b0.z; acc[q][3] = b0.w;
        acc[q][4] = b1.x; acc[q][5] = b1.y; acc[q][6] = b1.z; acc[q][7] = b1.w;
      }
    }
#pragma unroll
    for (int ci = 0; ci < CIN; ++ci) {
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        float a[6];
#pragma unroll
        for (int j = 0; j < 6; ++j) a[j] = tile[ci][ty + r][tx + j];
#pragma unroll
        for (int s2 = 0; s2 < 3; ++s2) {
          const int ct = ci * 9 + r * 3 + s2;
          const float4 wa = *reinterpret_cast<const float4*>(sw + ((ct * 2 + 0) * G + g) * 4);
          const float4 wb = *reinterpret_cast<const float4*>(sw + ((ct * 2 + 1) * G + g) * 4);
          const float w[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
          for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[q][i] = fmaf(a[s2 + q], w[i], acc[q][i]);
        }
      }
    }
    const int y = y0 + ty, x = x0 + tx;
    if (y < p.H) {
      const size_t o = ((size_t)(n + p.dst.n_off) * p.H * p.W + (size_t)y * p.W + x) * p.dst.C + p.dst.coff + g * 8;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (x + q < p.W) {
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[q][i] = apply_act(acc[q][i], p.act);
          store8(p.dst.hi, p.dst.lo, o + (size_t)q * p.dst.C, acc[q]);
        }
      }
    }
  }
  }
}

// ------------------------------------------------------------------------------------------
// MaxPool2d(2, stride s, ceil_mode=True) (vgg16_c.py:15,20,27,34).  Copies the winning hi/lo pair.
struct PoolParams {
  View src, dst;
  int B, Hi, Wi, Ho, Wo, stride, Cv;   // Cv: channels pooled (multiple of 8)
};

__global__ void maxpool_kernel(const PoolParams p) {
  const int groups = p.Cv / 8;
  const long long total = (long long)p.B * p.Ho * p.Wo * groups;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int gidx = (int)(idx % groups);
  const long long pix = idx / groups;
  const int x = (int)(pix % p.Wo);
  const int y = (int)((pix / p.Wo) % p.Ho);
  const int n = (int)(pix / ((long long)p.Wo * p.Ho));
  BF8 bh, bl;
  float best[8];
  bool first = true;
  for (int r = 0; r < 2; ++r) {
    const int yy = y * p.stride + r;
    if (yy >= p.Hi) continue;
    for (int s = 0; s < 2; ++s) {
      const int xx = x * p.stride + s;
      if (xx >= p.Wi) continue;
      const size_t i = ((size_t)(n + p.src.n_off) * p.Hi * p.Wi + (size_t)yy * p.Wi + xx) * p.src.C + p.src.coff + gidx * 8;
      const BF8 h = *reinterpret_cast<const BF8*>(p.src.hi + i);
      const BF8 l = *reinterpret_cast<const BF8*>(p.src.lo + i);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float v = join_bf16(h.v[k], l.v[k]);
        if (first || v > best[k]) { best[k] = v; bh.v[k] = h.v[k]; bl.v[k] = l.v[k]; }
      }
      first = false;
    }
  }
  const size_t o = ((size_t)(n + p.dst.n_off) * p.Ho * p.Wo + (size_t)y * p.Wo + x) * p.dst.C + p.dst.coff + gidx * 8;
  *reinterpret_cast<BF8*>(p.dst.hi + o) = bh;
  *reinterpret_cast<BF8*>(p.dst.lo + o) = bl;
}

// ------------------------------------------------------------------------------------------
// InstanceNorm2d(affine=False, eps=1e-5, biased variance) (RITnet_v2.py:37,56):
// y = act((x - mean) * rstd) with the per-(frame, channel) sum / sum of squares accumulated by the
// producing convolutions' epilogues, optionally followed by AvgPool2d(2) (Transition_down,
// RITnet_v2.py:40-44; the 1x1 conv that the reference applies before the pool is applied after it
// by the caller - both are linear, so conv(pool(z)) == pool(conv(z))).
struct NormApplyParams {
  View src, dst;
  const double* sums;   // [B][sums_C][2]; the window's channel c sits at sums_coff + c
  int sums_C, sums_coff;
  int B, H, W, Cv, act, pool;
};

__global__ void __launch_bounds__(256) instnorm_apply_kernel(const NormApplyParams p) {
  // grid (pixel slabs, frames): mean / rstd of the frame's channels are derived once per block
  extern __shared__ float2 na_mr[];       // [Cv] (mean, rstd)
  const int n = blockIdx.y;
  const double inv = 1.0 / ((double)p.H * p.W);
  for (int c = threadIdx.x; c < p.Cv; c += blockDim.x) {
    const double s = p.sums[((size_t)n * p.sums_C + p.sums_coff + c) * 2];
    const double q = p.sums[((size_t)n * p.sums_C + p.sums_coff + c) * 2 + 1];
    const double m = s * inv;
    double var = q * inv - m * m;
    if (var < 0.0) var = 0.0;
    na_mr[c] = make_float2((float)m, (float)(1.0 / sqrt(var + 1e-5)));
  }
  __syncthreads();
  const int groups = p.Cv / 8;
  const int Ho = p.pool ? p.H / 2 : p.H, Wo = p.pool ? p.W / 2 : p.W;
  const int total = Ho * Wo * groups;
  const int per = (total + gridDim.x - 1) / gridDim.x;
  const int i0 = blockIdx.x * per, i1 = min(total, i0 + per);
  for (int idx = i0 + threadIdx.x; idx < i1; idx += blockDim.x) {
    const int gidx = idx % groups;
    const int pix = idx / groups;
    const int x = pix % Wo, y = pix / Wo;
    float out[8];
    if (p.pool) {
#pragma unroll
      for (int i = 0; i < 8; ++i) out[i] = 0.f;
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int s2 = 0; s2 < 2; ++s2) {
          float v[8];
          load8(p.src.hi, p.src.lo,
                ((size_t)(n + p.src.n_off) * p.H * p.W + (size_t)(2 * y + r) * p.W + 2 * x + s2) * p.src.C + p.src.coff + gidx * 8, v);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float2 mr = na_mr[gidx * 8 + i];
            out[i] += apply_act((v[i] - mr.x) * mr.y, p.act);
          }
        }
#pragma unroll
      for (int i = 0; i < 8; ++i) out[i] *= 0.25f;
    } else {
      float v[8];
      load8(p.src.hi, p.src.lo, ((size_t)(n + p.src.n_off) * p.H * p.W + (size_t)y * p.W + x) * p.src.C + p.src.coff + gidx * 8, v);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float2 mr = na_mr[gidx * 8 + i];
        out[i] = apply_act((v[i] - mr.x) * mr.y, p.act);
      }
    }
    store8(p.dst.hi, p.dst.lo, ((size_t)(n + p.dst.n_off) * Ho * Wo + (size_t)y * Wo + x) * p.dst.C + p.dst.coff + gidx * 8, out);
  }
}


// ------------------------------------------------------------------------------------------
// InstanceNorm folded into the 3x3 convolution that follows it (RITnet_v2.py:59-60: conv1(norm(x))):
//   conv(W, (x - m) * r) = conv(W * r, x) - sum_{taps inside the frame} sum_k W[tap][:, k] * r[k] * m[k]
// (zero padding applies to the NORMALISED map, so taps that fall outside the frame contribute nothing - the
// mean term therefore depends on the pixel's border class: {first, inner, last} row x {first, inner, last} column).
// One block per frame writes that frame's scaled split-bf16 weights and its 9 bias vectors; the convolution then
// reads the raw map and the normalisation pass (one read + one write of the whole map) disappears.
struct InFoldParams {
  const float* w32;     // [ntaps][cout_pad][kpad] fp32, packed like the layer's bf16 weights (zero rows / columns for padding)
  const float* bias;    // [cout_pad]
  const double* sums;   // InstanceNorm statistics [N][sums_C][2]; K index k sits at channel sums_coff + k
  int sums_C, sums_coff;
  int ntaps, cout_pad, kpad, kreal, H, W;
  int8_t tap_dy[9], tap_dx[9];
  bf16* w_hi;           // [N][ntaps][cout_pad][kpad]
  bf16* w_lo;
  float* bias_fc;       // [N][9][cout_pad]
};

__global__ void __launch_bounds__(256) in_fold_kernel(const InFoldParams p) {
  extern __shared__ float inf_sm[];
  float* rstd = inf_sm;                 // [kpad]
  float* mr = rstd + p.kpad;            // [kpad] mean * rstd
  float* S = mr + p.kpad;               // [ntaps][cout_pad]
  const int n = blockIdx.x;
  const double inv = 1.0 / ((double)p.H * p.W);
  for (int k = threadIdx.x; k < p.kpad; k += blockDim.x) {
    float r = 0.f, m = 0.f;
    if (k < p.kreal) {
      const double s = p.sums[((size_t)n * p.sums_C + p.sums_coff + k) * 2];
      const double q = p.sums[((size_t)n * p.sums_C + p.sums_coff + k) * 2 + 1];
      const double mean = s * inv;
      double var = q * inv - mean * mean;
      if (var < 0.0) var = 0.0;
      r = (float)(1.0 / sqrt(var + 1e-5));
      m = (float)mean;
    }
    rstd[k] = r; mr[k] = m * r;
  }
  __syncthreads();
  const int rows = p.ntaps * p.cout_pad;
  const size_t fbase = (size_t)n * rows * p.kpad;
  for (int i = threadIdx.x; i < rows * p.kpad; i += blockDim.x) {
    const int k = i % p.kpad;
    bf16 h, l;
    split_bf16(p.w32[i] * rstd[k], h, l);
    p.w_hi[fbase + i] = h; p.w_lo[fbase + i] = l;
  }
  for (int r = threadIdx.x; r < rows; r += blockDim.x) {
    float s = 0.f;
    for (int k = 0; k < p.kreal; ++k) s = fmaf(p.w32[(size_t)r * p.kpad + k], mr[k], s);
    S[r] = s;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 9 * p.cout_pad; i += blockDim.x) {
    const int co = i % p.cout_pad, cls = i / p.cout_pad;
    const int ry = cls / 3, rx = cls % 3;
    float b = p.bias[co];
    for (int t = 0; t < p.ntaps; ++t) {
      const bool out = (ry == 0 && p.tap_dy[t] < 0) || (ry == 2 && p.tap_dy[t] > 0) || (rx == 0 && p.tap_dx[t] < 0) || (rx == 2 && p.tap_dx[t] > 0);
      if (!out) b -= S[t * p.cout_pad + co];
    }
    p.bias_fc[((size_t)n * 9 + cls) * p.cout_pad + co] = b;
  }
}

// ------------------------------------------------------------------------------------------
// Spatial mean of a channel window -> fp32 [B][Cs]  (latent, RITnet_v2.py:282).
// Block per frame; blockDim.x = 4 * 160: four pixel quarters per channel, combined in shared memory.
__global__ void spatial_mean_kernel(View src, float* out, int B, int HW, int Cs) {
  __shared__ float part[4][160];
  const int n = blockIdx.x;
  const int c = threadIdx.x % 160, qtr = threadIdx.x / 160;
  float s = 0.f;
  if (c < Cs) {
    const int per = (HW + 3) / 4;
    const int p1 = min(HW, (qtr + 1) * per);
    for (int px = qtr * per; px < p1; ++px) {
      const size_t i = ((size_t)(n + src.n_off) * HW + px) * src.C + src.coff + c;
      s += join_bf16(src.hi[i], src.lo[i]);
    }
  }
  part[qtr][c] = s;
  __syncthreads();
  if (qtr == 0 && c < Cs) out[(size_t)n * Cs + c] = (part[0][c] + part[1][c] + part[2][c] + part[3][c]) / (float)HW;
}

// ------------------------------------------------------------------------------------------
// BDCN tail (bdcn_new.py:118-191 collapsed): per full-resolution pixel
//   fuse = b + sum_k alpha_k * U_k(sA_k + cA_k) + beta_k * U_k(sB_k + cB_k),  edge = sigmoid(fuse)
// U_1 = identity, U_2..5 = the learnable ConvTranspose2d kernels followed by the crop.
struct BdcnTailParams {
  const float* score[5];   // [N][h][w][2]  (A = score_dsn k, B = score_dsn k_1), without biases
  int h[5], w[5];
  const float* kern[5];    // [K][K] fp32 (kern[0] unused)
  int K[5], stride[5], crop[5];
  int shift[5];            // log2(stride): the strides of bdcn_new.py:101-108 are 2 / 4 / 8 / 8
  float alpha[5], beta[5], cA[5], cB[5], fuse_bias;
  float* out;              // [N][H][W]
  // optional: the ten per-scale sigmoids BDCN.forward returns before the fused map (bdcn_new.py:163-191),
  // side + k * side_stride -> [N][H][W] for k = p1_1..p5_1, p1_2..p5_2; nullptr skips them (only [-1] is consumed)
  float* side;
  long long side_stride;
  int N, H, W;
};

__global__ void bdcn_tail_kernel(const BdcnTailParams p) {
  const long long total = (long long)p.N * p.H * p.W;
  const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= total) return;
  const int x = (int)(pix % p.W);
  const int y = (int)((pix / p.W) % p.H);
  const int n = (int)(pix / ((long long)p.W * p.H));
  const float2 s1 = reinterpret_cast<const float2*>(p.score[0])[pix];
  float acc = p.fuse_bias + p.alpha[0] * (s1.x + p.cA[0]) + p.beta[0] * (s1.y + p.cB[0]);
  float u[5], v[5];          // upsampled + cropped s_k / s_k1 at this pixel (used by the side outputs only)
  u[0] = s1.x + p.cA[0]; v[0] = s1.y + p.cB[0];
#pragma unroll
  for (int k = 1; k < 5; ++k) {
    const int st = p.stride[k], K = p.K[k];
    const int yy = y + p.crop[k], xx = x + p.crop[k];      // coordinate in the uncropped output
    // out[yy] = sum_i in[i] * kern[yy - i*st],  0 <= yy - i*st < K
    const int sh = p.shift[k];
    int i0 = (yy - K + st) >> sh;  if (yy - K + 1 <= 0) i0 = 0;
    int j0 = (xx - K + st) >> sh;  if (xx - K + 1 <= 0) j0 = 0;
    const int i1 = min(yy >> sh, p.h[k] - 1), j1 = min(xx >> sh, p.w[k] - 1);
    float ua = 0.f, ub = 0.f;
    for (int i = i0; i <= i1; ++i) {
      const int ky = yy - i * st;
      if (ky < 0 || ky >= K) continue;
      for (int j = j0; j <= j1; ++j) {
        const int kx = xx - j * st;
        if (kx < 0 || kx >= K) continue;
        const float kv = __ldg(p.kern[k] + ky * K + kx);
        const float2 s = reinterpret_cast<const float2*>(p.score[k])[((size_t)n * p.h[k] + i) * p.w[k] + j];
        ua = fmaf(kv, s.x + p.cA[k], ua);
        ub = fmaf(kv, s.y + p.cB[k], ub);
      }
    }
    acc += p.alpha[k] * ua + p.beta[k] * ub;
    u[k] = ua; v[k] = ub;
  }
  p.out[pix] = 1.f / (1.f + __expf(-acc));
  if (p.side) {
    // p_k_1 = s_k + o_{k-1} + ... + o_1 ; p_k_2 = s_k1 + o_{k+1,1} + ... + o_51   (bdcn_new.py:165-174)
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      a += u[k];
      b += v[4 - k];
      p.side[(long long)k * p.side_stride + pix] = 1.f / (1.f + __expf(-a));
      p.side[(long long)(9 - k) * p.side_stride + pix] = 1.f / (1.f + __expf(-b));
    }
  }
}

// ------------------------------------------------------------------------------------------
// AdaIN on the bottleneck before the regression head (RITnet_v2.py:297-308, 251-259: per
// (frame, channel) mean / UNBIASED variance + 1e-5 over the 300 bottleneck pixels, then
// gamma * x_hat + beta with the MLP's parameters).  Reads the image (and edge) bottleneck windows
// and writes the transformed channels into the split-bf16 buffer the head's first convolution reads;
// source s's channel c lands at dst channel s * Cslot + c.
struct AdainApplyParams {
  View src[2];
  int nsrc, Cs, Cslot;
  const float* adain;      // [B][2][Cf] (gamma, beta), Cf = nsrc * Cs
  View dst;
  int B, HW;
};

__global__ void adain_apply_kernel(const AdainApplyParams p) {
  const int Cf = p.nsrc * p.Cs;
  const int n = blockIdx.x;
  for (int c = threadIdx.x; c < Cf; c += blockDim.x) {
    const View& s = p.src[c / p.Cs];
    const int cc = c % p.Cs;
    const size_t base = (size_t)(n + s.n_off) * p.HW * s.C + s.coff + cc;
    float sum = 0.f;
    for (int px = 0; px < p.HW; ++px) sum += join_bf16(s.hi[base + (size_t)px * s.C], s.lo[base + (size_t)px * s.C]);
    const float mean = sum / p.HW;
    float q = 0.f;
    for (int px = 0; px < p.HW; ++px) {
      const float d = join_bf16(s.hi[base + (size_t)px * s.C], s.lo[base + (size_t)px * s.C]) - mean;
      q = fmaf(d, d, q);
    }
    const float istd = 1.f / sqrtf(q / (p.HW - 1) + 1e-5f);
    const float g = p.adain[((size_t)n * 2 + 0) * Cf + c], b = p.adain[((size_t)n * 2 + 1) * Cf + c];
    const size_t ob = (size_t)(n + p.dst.n_off) * p.HW * p.dst.C + p.dst.coff + (c / p.Cs) * p.Cslot + cc;
    for (int px = 0; px < p.HW; ++px) {
      const float v = join_bf16(s.hi[base + (size_t)px * s.C], s.lo[base + (size_t)px * s.C]);
      split_bf16((v - mean) * istd * g + b, p.dst.hi[ob + (size_t)px * p.dst.C], p.dst.lo[ob + (size_t)px * p.dst.C]);
    }
  }
}

// ------------------------------------------------------------------------------------------
// Regression head after its first convolution (utils.py:1013-1037): AvgPool2d(2) of the 14x18
// valid region of c1's output -> c2 3x3 valid (128->128, bias, lrelu) -> c3 3x3 valid (128->32,
// no bias, lrelu) -> flatten -> l1 (480->256, SELU) -> l2 (256->10) -> tanh / sigmoid split.
// One CTA per frame keeps every intermediate in shared memory.
struct HeadTailParams {
  View c1;              // [B][15][20][128] lrelu(c1(x)), valid region 14 x 18
  const float* w_c2;    // [3][3][128][128]  (kh, kw, ci, co)
  const float* b_c2;    // [128]
  const float* w_c3;    // [3][3][128][32]
  const float* w_l1t;   // [480][256], input index = (y*5 + x)*32 + c
  const float* b_l1;    // [256]
  const float* w_l2;    // [10][256]
  const float* b_l2;    // [10]
  float* el_out;        // [B][10]
  int B;
};

#define HEAD_TAIL_THREADS 256
#define HEAD_TAIL_SMEM ((63 * 128 + 35 * 128) * sizeof(float))

__global__ void __launch_bounds__(HEAD_TAIL_THREADS) head_tail_kernel(const HeadTailParams p) {
  extern __shared__ float ht_sm[];
  float* p1 = ht_sm;                  // [7*9][128] pooled c1
  float* c2o = ht_sm + 63 * 128;      // [5*7][128]
  float* c3o = p1;                    // [3*5][32]   (p1 is dead once c2 is done)
  float* l1o = p1 + 512;              // [256]
  const int n = blockIdx.x, t = threadIdx.x;
  // ---- AvgPool2d(2) (utils.py:990) of the valid 14x18 region
  for (int i = t; i < 63 * 16; i += HEAD_TAIL_THREADS) {
    const int g8 = i & 15, pix = i >> 4;
    const int py = pix / 9, px = pix % 9;
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        float v[8];
        load8(p.c1.hi, p.c1.lo, ((size_t)(n + p.c1.n_off) * 300 + (size_t)(2 * py + r) * 20 + 2 * px + s) * p.c1.C + p.c1.coff + g8 * 8, v);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] += v[k];
      }
#pragma unroll
    for (int k = 0; k < 8; ++k) p1[pix * 128 + g8 * 8 + k] = 0.25f * acc[k];
  }
  __syncthreads();
  // ---- c2: thread = (output channel, pixel parity); 18 pixels per thread
  {
    const int co = t & 127, half = t >> 7;
    float acc[18];
    int off[18];
    const float bias = p.b_c2[co];
#pragma unroll
    for (int i = 0; i < 18; ++i) {
      const int q = min(half + 2 * i, 34);
      acc[i] = bias;
      off[i] = ((q / 7) * 9 + q % 7) * 128;
    }
    for (int tap = 0; tap < 9; ++tap) {
      const int toff = ((tap / 3) * 9 + tap % 3) * 128;
      const float* w = p.w_c2 + (size_t)tap * 128 * 128 + co;
      for (int ci = 0; ci < 128; ci += 4) {
        const float w0 = __ldg(w + (size_t)ci * 128), w1 = __ldg(w + (size_t)(ci + 1) * 128);
        const float w2 = __ldg(w + (size_t)(ci + 2) * 128), w3 = __ldg(w + (size_t)(ci + 3) * 128);
#pragma unroll
        for (int i = 0; i < 18; ++i) {
          const float4 a = *reinterpret_cast<const float4*>(p1 + off[i] + toff + ci);
          acc[i] = fmaf(a.x, w0, fmaf(a.y, w1, fmaf(a.z, w2, fmaf(a.w, w3, acc[i]))));
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 18; ++i) {
      const int q = half + 2 * i;
      if (q < 35) c2o[q * 128 + co] = apply_act(acc[i], ACT_LRELU);
    }
  }
  __syncthreads();
  // ---- c3: thread = (output channel, pixel group); pixels pg and pg + 8 of the 3x5 map
  {
    const int co = t & 31, pg = t >> 5;
    const int q0 = pg, q1 = min(pg + 8, 14);
    const int o0 = ((q0 / 5) * 7 + q0 % 5) * 128, o1 = ((q1 / 5) * 7 + q1 % 5) * 128;
    float a0 = 0.f, a1 = 0.f;
    for (int tap = 0; tap < 9; ++tap) {
      const int toff = ((tap / 3) * 7 + tap % 3) * 128;
      const float* w = p.w_c3 + (size_t)tap * 128 * 32 + co;
#pragma unroll 4
      for (int ci = 0; ci < 128; ci += 4) {
        const float w0 = __ldg(w + ci * 32), w1 = __ldg(w + (ci + 1) * 32), w2 = __ldg(w + (ci + 2) * 32), w3 = __ldg(w + (ci + 3) * 32);
        const float4 x0 = *reinterpret_cast<const float4*>(c2o + o0 + toff + ci);
        const float4 x1 = *reinterpret_cast<const float4*>(c2o + o1 + toff + ci);
        a0 = fmaf(x0.x, w0, fmaf(x0.y, w1, fmaf(x0.z, w2, fmaf(x0.w, w3, a0))));
        a1 = fmaf(x1.x, w0, fmaf(x1.y, w1, fmaf(x1.z, w2, fmaf(x1.w, w3, a1))));
      }
    }
    c3o[q0 * 32 + co] = apply_act(a0, ACT_LRELU);
    if (pg + 8 < 15) c3o[(pg + 8) * 32 + co] = apply_act(a1, ACT_LRELU);
  }
  __syncthreads();
  // ---- l1 + SELU (utils.py:1021)
  {
    float acc = p.b_l1[t];
#pragma unroll 8
    for (int i = 0; i < 480; ++i) acc = fmaf(c3o[i], __ldg(p.w_l1t + (size_t)i * 256 + t), acc);
    const float alpha = 1.6732632423543772848170429916717f, scale = 1.0507009873554804934193349852946f;
    l1o[t] = scale * (acc > 0.f ? acc : alpha * (expf(acc) - 1.f));
  }
  __syncthreads();
  // ---- l2 + the ellipse split (utils.py:1022-1036): one warp per output
  for (int o = t >> 5; o < 10; o += HEAD_TAIL_THREADS / 32) {
    const int lane = t & 31;
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc = fmaf(l1o[lane + 32 * i], __ldg(p.w_l2 + o * 256 + lane + 32 * i), acc);
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
    if (lane == 0) {
      acc += p.b_l2[o];
      const int k = o % 5;
      if (k < 2) acc = tanhf(acc);
      else if (k < 4) acc = 1.f / (1.f + expf(-acc));
      p.el_out[(size_t)n * 10 + o] = acc;
    }
  }
}

// ------------------------------------------------------------------------------------------
// AdaIN style encoder staging (RITnet_v2.py:91-106, Conv2dBlock utils.py:1093-1149, reflect padding).
// Layer 0 (7x7, 3 -> 64, reflect 3): the 7 horizontal neighbours x 3 classes of the softmaxed
// segmentation are folded into 21 (padded to 32) channels of a [B][H+6][W][32] tensor whose rows
// already carry the vertical reflection, so the layer runs as a 7x1 convolution on the tensor cores.
struct StyleFoldParams {
  const float* sm;    // softmax, fp32 NHWC [B][H*W][3]
  View dst;           // [B][H+6][W][32]
  int B, H, W;
};

__device__ __forceinline__ int reflect_idx(int v, int n) {
  if (v < 0) v = -v;
  if (v >= n) v = 2 * n - 2 - v;
  return v;
}

__global__ void __launch_bounds__(256) style_fold_kernel(const StyleFoldParams p) {
  const int Hp = p.H + 6;
  const long long total = (long long)p.B * Hp * p.W * 4;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int g = (int)(idx & 3);
  const long long pix = idx >> 2;
  const int x = (int)(pix % p.W);
  const int r = (int)((pix / p.W) % Hp);
  const int n = (int)(pix / ((long long)p.W * Hp));
  const int y = reflect_idx(r - 3, p.H);
  float v[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int ch = g * 8 + k;
    float val = 0.f;
    if (ch < 21) {
      const int dx = ch / 3, c = ch % 3;
      val = p.sm[((size_t)n * p.H * p.W + (size_t)y * p.W + reflect_idx(x + dx - 3, p.W)) * 3 + c];
    }
    v[k] = val;
  }
  store8(p.dst.hi, p.dst.lo, ((size_t)(n + p.dst.n_off) * Hp * p.W + (size_t)r * p.W + x) * p.dst.C + p.dst.coff + g * 8, v);
}

// Layers 1-4 (4x4, stride 2, reflect 1): reflect-pad by one pixel and space-to-depth by 2, so the
// layer becomes a 2x2 stride-1 convolution over 4C channels on the [(H+2)/2][(W+2)/2] grid (its last
// row / column of outputs is never read).  Copies the hi/lo planes verbatim.
struct S2dParams {
  View src, dst;
  int B, Hg, Wg;      // source grid (rows/cols allocated)
  int Hs, Ws, C;      // valid source region and channels
};

__global__ void __launch_bounds__(256) s2d_reflect_kernel(const S2dParams p) {
  const int Ho = (p.Hs + 2) / 2, Wo = (p.Ws + 2) / 2, groups = p.C / 8;
  const long long total = (long long)p.B * Ho * Wo * 4 * groups;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int g = (int)(idx % groups);
  long long rest = idx / groups;
  const int par = (int)(rest & 3);
  rest >>= 2;
  const int ox = (int)(rest % Wo);
  const int oy = (int)((rest / Wo) % Ho);
  const int n = (int)(rest / ((long long)Wo * Ho));
  const int sy = reflect_idx(2 * oy + (par >> 1) - 1, p.Hs), sx = reflect_idx(2 * ox + (par & 1) - 1, p.Ws);
  const size_t si = ((size_t)(n + p.src.n_off) * p.Hg * p.Wg + (size_t)sy * p.Wg + sx) * p.src.C + p.src.coff + g * 8;
  const size_t di = ((size_t)(n + p.dst.n_off) * Ho * Wo + (size_t)oy * Wo + ox) * p.dst.C + p.dst.coff + par * p.C + g * 8;
  *reinterpret_cast<BF8*>(p.dst.hi + di) = *reinterpret_cast<const BF8*>(p.src.hi + si);
  *reinterpret_cast<BF8*>(p.dst.lo + di) = *reinterpret_cast<const BF8*>(p.src.lo + si);
}

// AdaptiveAvgPool2d(1) over the valid Hv x Wv region of a split-bf16 grid -> fp32 [B][C]
__global__ void gap_act_kernel(View src, float* out, int Hg, int Wg, int Hv, int Wv, int C) {
  const int n = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (int y = 0; y < Hv; ++y)
      for (int x = 0; x < Wv; ++x) {
        const size_t i = ((size_t)(n + src.n_off) * Hg * Wg + (size_t)y * Wg + x) * src.C + src.coff + c;
        s += join_bf16(src.hi[i], src.lo[i]);
      }
    out[(size_t)n * C + c] = s / (float)(Hv * Wv);
  }
}

// Linear layer fp32: out[b][o] = act(in[b] . w[o] + bias[o]);  act: 0 none, 1 relu, 3 selu,
// 4 = ellipse head split (tanh on 0:2,5:7; sigmoid on 2:4,7:9; identity on 4,9; utils.py:1023-1036).
__global__ void linear_kernel(const float* in, const float* w, const float* bias, float* out, int B,
                              int I, int O, int act) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * O) return;
  const int o = idx % O, b = idx / O;
  float acc = bias[o];
  const float* a = in + (size_t)b * I;
  const float* ww = w + (size_t)o * I;
  for (int i = 0; i < I; ++i) acc = fmaf(a[i], ww[i], acc);
  if (act == 1) acc = fmaxf(acc, 0.f);
  else if (act == 3) {
    const float alpha = 1.6732632423543772848170429916717f, scale = 1.0507009873554804934193349852946f;
    acc = scale * (acc > 0.f ? acc : alpha * (expf(acc) - 1.f));
  } else if (act == 4) {
    const int k = o % 5;
    if (k < 2) acc = tanhf(acc);
    else if (k < 4) acc = 1.f / (1.f + expf(-acc));
  }
  out[idx] = acc;
}

// softmax over the 3 classes of fp32 NCHW logits -> fp32 NHWC [B][H][W][3] (RITnet_v2.py:290-294)
__global__ void softmax3_nhwc_kernel(const float* logits, float* out, int B, int HW) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * HW) return;
  const int n = (int)(idx / HW);
  const int px = (int)(idx % HW);
  const float* l = logits + (size_t)n * 3 * HW + px;
  const float a = l[0], b = l[HW], c = l[2 * (size_t)HW];
  const float m = fmaxf(a, fmaxf(b, c));
  const float ea = expf(a - m), eb = expf(b - m), ec = expf(c - m);
  const float inv = 1.f / (ea + eb + ec);
  out[idx * 3 + 0] = ea * inv;
  out[idx * 3 + 1] = eb * inv;
  out[idx * 3 + 2] = ec * inv;
}

template <typename K, typename P>
static void launch_1d(K kernel, const P& p, long long total, cudaStream_t stream, int threads = 256) {
  const long long blocks = (total + threads - 1) / threads;
  kernel<<<(unsigned)blocks, threads, 0, stream>>>(p);
  CUDA_OK(cudaGetLastError());
}
