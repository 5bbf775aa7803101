// Shared types for the egn sm_100a engine.
//
// Activation storage ("split-bf16 NHWC"): every feature map is held as two bf16 planes
// hi = bf16(x), lo = bf16(x - hi) laid out [N][H][W][C] with C a multiple of 8.  hi+lo carries
// ~16 mantissa bits, which is what the parity bar of BASELINE.json needs;
// the tensor-core convolution multiplies (a_hi + a_lo) * (w_hi + w_lo) as three bf16 MMAs
// (hi*hi + lo*hi + hi*lo) with fp32 accumulation in TMEM.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <stdexcept>

typedef __nv_bfloat16 bf16;

#define EGN_H 240
#define EGN_W 320

struct EgnError : public std::runtime_error {
  explicit EgnError(const std::string& s) : std::runtime_error(s) {}
};

#define EGN_CHECK(cond, msg)                                                         \
  do {                                                                               \
    if (!(cond)) {                                                                   \
      throw EgnError(std::string(__FILE__) + ":" + std::to_string(__LINE__) + ": " + \
                     (msg));                                                         \
    }                                                                                \
  } while (0)

#define CUDA_OK(expr)                                                                   \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess) {                                                            \
      throw EgnError(std::string(__FILE__) + ":" + std::to_string(__LINE__) + ": " #expr \
                     " -> " + cudaGetErrorString(_e));                                  \
    }                                                                                   \
  } while (0)

// A split-bf16 NHWC buffer (device memory).
struct Act {
  bf16* hi = nullptr;
  bf16* lo = nullptr;
  int N = 0, H = 0, W = 0, C = 0;
  size_t plane_elems() const { return (size_t)N * H * W * C; }
};

enum ActKind { ACT_NONE = 0, ACT_RELU = 1, ACT_LRELU = 2 };

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == ACT_RELU) return fmaxf(v, 0.f);
  if (act == ACT_LRELU) return v > 0.f ? v : 0.01f * v;   // F.leaky_relu default slope
  return v;
}

__device__ __forceinline__ void split_bf16(float v, bf16& h, bf16& l) {
  h = __float2bfloat16_rn(v);
  l = __float2bfloat16_rn(v - __bfloat162float(h));
}

__device__ __forceinline__ float join_bf16(bf16 h, bf16 l) {
  return __bfloat162float(h) + __bfloat162float(l);
}

// 8 channels (16 bytes) of one plane
struct __align__(16) BF8 {
  bf16 v[8];
};

__device__ __forceinline__ void load8(const bf16* hi, const bf16* lo, size_t idx, float out[8]) {
  BF8 a = *reinterpret_cast<const BF8*>(hi + idx);
  BF8 b = *reinterpret_cast<const BF8*>(lo + idx);
#pragma unroll
  for (int i = 0; i < 8; ++i) out[i] = join_bf16(a.v[i], b.v[i]);
}

__device__ __forceinline__ void store8(bf16* hi, bf16* lo, size_t idx, const float in[8]) {
  BF8 a, b;
#pragma unroll
  for (int i = 0; i < 8; ++i) split_bf16(in[i], a.v[i], b.v[i]);
  *reinterpret_cast<BF8*>(hi + idx) = a;
  *reinterpret_cast<BF8*>(lo + idx) = b;
}

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline int round_up(int a, int b) { return ceil_div(a, b) * b; }

// ---- convolution description shared by the tensor-core kernel and its SIMT reference ----
#define EGN_MAX_TAPS 27
#define EGN_MAX_CHUNKS 32
#define EGN_MAX_SRC 4
#define EGN_KC 32  // channels per K chunk (64-byte swizzled rows)

enum ConvMode { CONV_STORE = 0, CONV_MSBLOCK = 1, CONV_LOGITS = 2 };

struct ConvGeom {
  int H, W, batch;            // output == input spatial size (stride 1, "same" padding)
  int ntaps, nchunks, groups; // groups: independent accumulators (1; 3 for the fused MSBlock tail)
  int cout_pad;               // weight rows per tap (multiple of 16)
  int kpad;                   // nchunks * EGN_KC
  int phase;                  // tensor-core kernel only: d > 1 = the layer runs on the d x d polyphase lattice of its
                              // input (H, W are the lattice sizes, batch counts frames x d*d phases, tap offsets are
                              // in lattice units); a dilation-(k*d) 3x3 becomes a dilation-k 3x3 there
  int8_t tap_dy[EGN_MAX_TAPS], tap_dx[EGN_MAX_TAPS], tap_grp[EGN_MAX_TAPS];
  uint8_t chunk_src[EGN_MAX_CHUNKS];
  int16_t chunk_c0[EGN_MAX_CHUNKS];
  int chunk_noff[EGN_MAX_CHUNKS];
};

struct ConvSrc {
  const bf16* hi;
  const bf16* lo;
  int C;   // channels of the source buffer
  int N;   // frames in the source buffer
};

struct ConvEpi {
  int mode, act;
  int cout_store;          // channels written (multiple of 8, <= cout_pad)
  const float* bias;       // [groups][cout_pad]
  const float* bias_fc;    // optional per-frame, per-border-class bias [N][9][cout_pad] (InstanceNorm folded into the layer; tensor-core kernel only)
  const float* post_scale; // optional affine after the activation (eval BatchNorm), [cout_pad]
  const float* post_shift;
  bf16* out_hi;
  bf16* out_lo;
  int out_C, out_coff;
  // optional InstanceNorm statistics of the stored values: stats[(n*stats_C + stats_coff + ch)*2 + {0: sum, 1: sum of squares}]
  double* stats;
  int stats_C, stats_coff;
  // optional fused MaxPool2d(2, 2) (vgg16_c.py:15,20,27; even H and W): channels [0, pool_ch) of the stored map
  // are also written, pooled, to a half-resolution buffer of pool_C channels (tensor-core kernel only)
  bf16* pool_hi;
  bf16* pool_lo;
  int pool_C, pool_ch;
  // CONV_MSBLOCK: v = o + sum_g relu(acc_g + b_g); score[p][j] (+)= v . score_w[j]
  const bf16* o_hi;
  const bf16* o_lo;
  int o_C, o_coff;         // channels of the buffer holding o and its first channel there
  const float* score_w;    // [2][32]
  float* score;            // [N][H][W][2]
  int score_accum;
  // CONV_LOGITS: the first logits_c channels go to an fp32 NCHW tensor [N][logits_c][H][W] (the
  // segmentation logits the reference returns, models/RITnet_v2.py:288) instead of a split-bf16 buffer
  float* logits;
  int logits_c;
  // optional "+ bilinear x2 upsample of a half-resolution tensor" before the activation: the decoder's
  // 1x1 convolutions over cat[upsample(x), skip] are evaluated as upsample(W_a x) + W_b skip + b
  // (both are linear and the interpolation weights sum to one), so the upsampled x never exists
  const bf16* up_hi;
  const bf16* up_lo;
  int up_C, up_coff;       // channels of the half-resolution buffer, first channel of the window
};

// v[0..NCH) += bilinear x2 upsample (F.interpolate, align_corners=False, models/RITnet_v2.py:82) of the
// half-resolution split-bf16 tensor at output pixel (py, px) of frame n, channels [ch, ch + NCH).
// For o = 2i: 0.25 * in[i-1] + 0.75 * in[i]; for o = 2i+1: 0.75 * in[i] + 0.25 * in[i+1], indices clamped.
template <int NCH>
__device__ __forceinline__ void upsample_add(const ConvEpi& e, int n, int py, int px, int H, int W, int ch, float* v) {
  const int Hi = H >> 1, Wi = W >> 1;
  const int i = py >> 1, j = px >> 1;
  const int ya = (py & 1) ? i : max(i - 1, 0), yb = (py & 1) ? min(i + 1, Hi - 1) : i;
  const int xa = (px & 1) ? j : max(j - 1, 0), xb = (px & 1) ? min(j + 1, Wi - 1) : j;
  const float wya = (py & 1) ? 0.75f : 0.25f, wyb = 1.f - wya;
  const float wxa = (px & 1) ? 0.75f : 0.25f, wxb = 1.f - wxa;
  const size_t fb = (size_t)n * Hi * Wi;
  const size_t c0 = (size_t)e.up_coff + ch;
#pragma unroll
  for (int g = 0; g < NCH; g += 8) {
    float t00[8], t01[8], t10[8], t11[8];
    load8(e.up_hi, e.up_lo, (fb + (size_t)ya * Wi + xa) * e.up_C + c0 + g, t00);
    load8(e.up_hi, e.up_lo, (fb + (size_t)ya * Wi + xb) * e.up_C + c0 + g, t01);
    load8(e.up_hi, e.up_lo, (fb + (size_t)yb * Wi + xa) * e.up_C + c0 + g, t10);
    load8(e.up_hi, e.up_lo, (fb + (size_t)yb * Wi + xb) * e.up_C + c0 + g, t11);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float top = wxa * t00[k] + wxb * t01[k];
      const float bot = wxa * t10[k] + wxb * t11[k];
      v[g + k] += wya * top + wyb * bot;
    }
  }
}
