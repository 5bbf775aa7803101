// Graph executor for the BDCN edge net + ESF-Net forward (host side of libegn.so).
//
// Reference graph: bdcn_new.py:116-191, vgg16_c.py:65-88, models/RITnet_v2.py:261-354,
// utils.py:1013-1037.  The engine owns every activation buffer (split-bf16 NHWC, see common.cuh),
// repacks the reference state_dict tensors into tensor-core operands once, and runs the forward
// in micro-batches so the workspace stays bounded for any caller batch.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "aux.cuh"
#include "common.cuh"
#include "conv_simt.cuh"
#include "conv_tc.cuh"
#include "post.cuh"

struct HostTensor {
  std::vector<int64_t> shape;
  std::vector<float> data;
  int64_t numel() const {
    int64_t n = 1;
    for (auto s : shape) n *= s;
    return n;
  }
};
typedef std::map<std::string, HostTensor> StateDict;

// Blob layout (little endian), produced by egn_b200.pack.pack_state_dict:
//   u32 magic 'EGNW', u32 count, then per tensor: u32 name_len, name bytes, u32 ndim,
//   i64 dims[ndim], f32 data[numel]
static StateDict parse_blob(const void* blob, size_t bytes) {
  StateDict sd;
  const uint8_t* p = (const uint8_t*)blob;
  const uint8_t* end = p + bytes;
  auto need = [&](size_t n) { EGN_CHECK((size_t)(end - p) >= n, "weight blob truncated"); };
  need(8);
  uint32_t magic, count;
  memcpy(&magic, p, 4); memcpy(&count, p + 4, 4); p += 8;
  EGN_CHECK(magic == 0x574e4745u, "weight blob: bad magic");
  for (uint32_t i = 0; i < count; ++i) {
    need(4);
    uint32_t nl; memcpy(&nl, p, 4); p += 4;
    need(nl + 4);
    std::string name((const char*)p, nl); p += nl;
    uint32_t nd; memcpy(&nd, p, 4); p += 4;
    HostTensor t;
    need(8 * (size_t)nd);
    for (uint32_t d = 0; d < nd; ++d) { int64_t v; memcpy(&v, p, 8); p += 8; t.shape.push_back(v); }
    const int64_t n = t.numel();
    need(4 * (size_t)n);
    t.data.resize(n);
    memcpy(t.data.data(), p, 4 * (size_t)n); p += 4 * (size_t)n;
    sd[name] = std::move(t);
  }
  return sd;
}

static const HostTensor& sd_get(const StateDict& sd, const std::string& k) {
  auto it = sd.find(k);
  EGN_CHECK(it != sd.end(), "missing tensor in state_dict: " + k);
  return it->second;
}

// ------------------------------------------------------------------------------------------
struct DevMem {
  std::vector<void*> ptrs;
  size_t total = 0;
  void* alloc(size_t bytes, bool zero = true) {
    void* p = nullptr;
    if (bytes == 0) bytes = 16;
    CUDA_OK(cudaMalloc(&p, bytes));
    if (zero) CUDA_OK(cudaMemset(p, 0, bytes));
    ptrs.push_back(p);
    total += bytes;
    return p;
  }
  template <typename T>
  T* upload(const std::vector<T>& v) {
    T* d = (T*)alloc(v.size() * sizeof(T), false);
    if (!v.empty()) CUDA_OK(cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return d;
  }
  void release() {
    for (void* p : ptrs) cudaFree(p);
    ptrs.clear();
    total = 0;
  }
};

static inline bf16 host_bf16(float v) { return __float2bfloat16_rn(v); }

// Opt-in (egn_share_workspace / EGN_SHARE_WORKSPACE=1): ONE activation arena per device for every context of the
// process.  The BDCN context's buffers are dead once its edge map is out, so an evaluator that drives both modules
// in stream order (bench.py, calc_acc) can lend them to the ESF-Net: max(19, 37) instead of 19 + 37 GB at micro-batch
// 256.  Off by default: two independent nn.Modules may legitimately run on different streams.
struct SharedPool { void* base = nullptr; size_t bytes = 0; int refs = 0; };
static std::map<int, SharedPool>& shared_pools() { static std::map<int, SharedPool> m; return m; }
static bool& share_workspace_flag() {
  static bool f = getenv("EGN_SHARE_WORKSPACE") && atoi(getenv("EGN_SHARE_WORKSPACE")) != 0;
  return f;
}

struct Piece {       // a run of reference input channels living in a buffer
  const Act* buf;
  int buf_c, len, n_off;
};

struct ConvLayer {
  std::string name;
  ConvGeom g{};
  ConvEpi e{};
  ConvSrc src[EGN_MAX_SRC]{};
  int nsrc = 0;
  const Act* src_act[EGN_MAX_SRC]{};
  bf16* w_hi = nullptr;
  bf16* w_lo = nullptr;
  int cout = 0;
  TcParams tc{};
  SimtParams simt{};
  int w_frames = 0;      // > 0: per-frame weights (InstanceNorm folded into the layer): fw_hi / fw_lo [w_frames][ntaps][cout_pad][kpad]
  bf16* fw_hi = nullptr;
  bf16* fw_lo = nullptr;
  InFoldParams fold{};
  int prod_mode = 0;     // precision probe: products left out for this layer (TcParams::prod_mode)
  int phase = 0;         // tensor-core path runs on the phase x phase polyphase lattice of its input (MSBlock tails, stages 1-2)
  double flops = 0;      // algorithmic 2*MAC at unpadded sizes, per frame
  double prof_ms = 0;    // accumulated kernel time (profiling mode)
  double prof_frames = 0;
  long long prof_n = 0;
};

struct EgnConfig {
  int add_edge = 0, add_seg = 0, input_concat = 0, only_edge = 0, style_dim = 8, seg_detach = 0;
};

struct Engine {
  int device = 0;
  int num_sms = 148;
  EgnConfig cfg;
  int mb = 0;                 // micro-batch (frames)
  bool use_tc = true;
  int nsplit = 3;
  StateDict sd_bdcn, sd_esf;
  bool has_bdcn = false, has_esf = false, built_bdcn = false, built_esf = false;
  DevMem mem_bdcn, mem_esf, mem_misc;
  int* err_flag = nullptr;
  std::map<std::string, std::pair<const Act*, std::pair<int, int>>> debug_acts;  // name -> (buf,(coff,C))
  std::map<std::string, std::pair<const float*, std::vector<int>>> debug_f32;
  std::vector<std::unique_ptr<Act>> acts;
  std::map<std::string, ConvLayer*> conv_index;
  std::map<std::string, ConvLayer*> conv_alias;     // reference layer names of merged launches (lookup only)

  // ---- BDCN
  struct {
    Act *f[13], *pool[4], *o[5];
    float* score[5];
    int h[5], w[5];
    FirstConvParams first;
    const float *first_w_gray, *first_w_rgb;
    ConvLayer vgg[13];           // index 0 unused (first_conv)
    ConvLayer ms_in[13], ms_tail[13];
    bool pool_fused[4];          // pool k (2x2 / stride 2) is written by the epilogue of the convolution that produces its input
    bool merged[13];             // merged[i]: msblock i's `conv` rides as 32 extra output channels of vgg[i + 1] (same input map)
    int f_c[13];                 // VGG channels of f[i] (a merged f[i] buffer holds 32 more: the MSBlock's `o`)
    BdcnTailParams tail;
  } bd;

  // ---- ESF
  struct Block {
    Act *buf, *xn, *t, *tdin;
    int in_c, in_pad, inter, op_c, off_out, off_x, off_x1, off_x22, H, W;
    bool fold;          // InstanceNorm(x) folded into conv1 through per-frame weights (no xn buffer, no normalisation pass)
    double* stats;      // [E][inter + in_pad][2]: sum / sum of squares per buffer channel of [out | x]
    int stats_C;
    ConvLayer conv1, conv21, conv22, conv31, conv32, td;
  };
  struct UpBlock {
    Act *buf, *t, *out, *y;     // buf: x1; y: half-resolution [W11_a x | W21_a x] (see build_esf)
    int in_c, out_c, out_pad, skip_c, H, W;
    ConvLayer pre, c11, c12, c21, c22;
  };
  struct {
    int E = 0;                    // encoder frames per micro-batch
    Act *h1, *bt, *fin;
    FirstConvParams first;
    ConvLayer head2, final1, final2;
    Block blk[5];
    UpBlock up[4];
    // regression head + AdaIN (fp32)
    int Cf = 0;
    Act *hin, *c1o;               // AdaIN-transformed head input (add_seg only); lrelu(c1(x))
    ConvLayer head_c1;
    HeadTailParams head;
    float *sm, *gap, *sty, *m1, *m2, *adain;
    Act *st_in[5], *st_out[5];    // style encoder: staged (folded / space-to-depth) inputs and outputs
    ConvLayer st_conv[5];
    std::vector<float> st_w[5];   // host copies of the remapped style-encoder weights (build time only)
    float *w_se6, *b_se6, *w_m[3], *b_m[3];
    float* logits_tmp;
    double* stats_all;
    size_t stats_bytes;
  } es;

  bool uses_pool = false;
  ~Engine() {
    mem_bdcn.release(); mem_esf.release(); mem_misc.release();
    if (uses_pool) {
      SharedPool& p = shared_pools()[device];
      if (--p.refs == 0) { cudaFree(p.base); p = SharedPool(); }
    }
    for (cudaEvent_t e : side_events) cudaEventDestroy(e);
    if (side_stream) cudaStreamDestroy(side_stream);
  }

  // ---- side stream (streaming micro-batches only).  With a handful of frames most launches fill a fraction of the
  // SMs, and the graph has independent branches: every MSBlock (its `conv` and its tail only feed the score maps,
  // bdcn_new.py:120-160) next to the VGG trunk, and the regression head (utils.py:1013-1037, reads the bottleneck
  // only) next to the decoder.  In that regime the branches run on an internal second stream, forked and joined
  // with events relative to the caller's stream (legal inside a caller's graph capture too).  Buffers are then NOT
  // shared by liveness (the schedule ticks no longer describe the execution order; the workspace is small there).
  cudaStream_t side_stream = nullptr;
  std::vector<cudaEvent_t> side_events;
  size_t side_used = 0;
  bool concurrent = false;            // decided by egn_plan: micro-batch <= EGN_SIDE_STREAM_MAX_MB (default 16)

  void decide_concurrency() {
    const int max_mb = getenv("EGN_SIDE_STREAM_MAX_MB") ? atoi(getenv("EGN_SIDE_STREAM_MAX_MB")) : 16;
    concurrent = use_tc && mb <= max_mb;
  }
  cudaEvent_t next_event() {
    if (side_used == side_events.size()) {
      cudaEvent_t e;
      CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      side_events.push_back(e);
    }
    return side_events[side_used++];
  }
  // work enqueued on `to` after this call waits for everything enqueued on `from` so far
  void stream_edge(cudaStream_t from, cudaStream_t to) {
    cudaEvent_t e = next_event();
    CUDA_OK(cudaEventRecord(e, from));
    CUDA_OK(cudaStreamWaitEvent(to, e, 0));
  }
  cudaStream_t side_of(cudaStream_t st) {
    if (!concurrent || profiling) return st;
    if (!side_stream) CUDA_OK(cudaStreamCreateWithFlags(&side_stream, cudaStreamNonBlocking));
    return side_stream;
  }

  // per-device function attributes (dynamic shared memory above 48 KB); called by egn_create with
  // this engine's device current
  static void prepare_device() {
    tc_prepare_device();
    CUDA_OK(cudaFuncSetAttribute(head_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HEAD_TAIL_SMEM));
  }

  // ---- liveness-planned workspace.  Every split-bf16 activation declares the interval [t0, t1] of the forward
  // schedule (ticks, see build_bdcn / build_esf) in which it is live; commit_acts packs the planes of a net into ONE
  // allocation so that buffers with disjoint intervals share memory (largest first, lowest offset that is free
  // for the whole interval).  Sharing is safe because (1) every plane holds bf16 values written by our own
  // kernels or zeros, so a K chunk that over-reads into a not-yet-written window meets finite numbers times
  // zero weights, and (2) a context runs its forwards in stream order.  Buffers that rely on channels nobody
  // ever writes staying zero are declared with t0 = 0, t1 = INT_MAX (never shared).  EGN_NO_ARENA=1 gives every
  // buffer its own allocation (tools/gpu_debug.py reads intermediate maps after the forward).
  struct ActReq { Act* act; int t0, t1; size_t off_hi, off_lo; };
  std::vector<ActReq> pending_acts;
  size_t arena_bytes = 0, arena_naive_bytes = 0;

  Act* new_act(DevMem& mem, int N, int H, int W, int C, int t0 = 0, int t1 = 0x7fffffff) {
    EGN_CHECK(C % 8 == 0, "Act channels must be a multiple of 8");
    std::unique_ptr<Act> a(new Act());
    a->N = N; a->H = H; a->W = W; a->C = C;
    static const bool no_arena = getenv("EGN_NO_ARENA") != nullptr;
    if (no_arena || concurrent) {
      const size_t bytes = a->plane_elems() * sizeof(bf16);
      a->hi = (bf16*)mem.alloc(bytes);
      a->lo = (bf16*)mem.alloc(bytes);
    } else {
      pending_acts.push_back({a.get(), t0, t1, 0, 0});
    }
    acts.push_back(std::move(a));
    return acts.back().get();
  }

  void commit_acts(DevMem& mem) {
    if (pending_acts.empty()) return;
    const bool share = share_workspace_flag();
    if (share) {
      // buffers that rely on never-written channels staying zero cannot live in memory another context writes
      std::vector<ActReq> rest;
      for (auto& r : pending_acts) {
        if (r.t1 == 0x7fffffff) {
          const size_t bytes = r.act->plane_elems() * sizeof(bf16);
          r.act->hi = (bf16*)mem.alloc(bytes); r.act->lo = (bf16*)mem.alloc(bytes);
          arena_naive_bytes += 2 * bytes;
        } else {
          rest.push_back(r);
        }
      }
      pending_acts.swap(rest);
      if (pending_acts.empty()) return;
    }
    struct Item { size_t bytes; int t0, t1; size_t off; bool placed; };
    std::vector<Item> items;
    for (auto& r : pending_acts) {
      const size_t bytes = (r.act->plane_elems() * sizeof(bf16) + 1023) & ~(size_t)1023;
      items.push_back({bytes, r.t0, r.t1, 0, false});    // hi plane
      items.push_back({bytes, r.t0, r.t1, 0, false});    // lo plane
      arena_naive_bytes += 2 * bytes;
    }
    std::vector<int> order(items.size());
    for (size_t i = 0; i < order.size(); ++i) order[i] = (int)i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return items[a].bytes > items[b].bytes; });
    size_t total = 0;
    for (int idx : order) {
      Item& it = items[idx];
      std::vector<std::pair<size_t, size_t>> busy;       // [begin, end) of placed items alive at the same time
      for (const Item& o : items)
        if (o.placed && !(o.t1 < it.t0 || it.t1 < o.t0)) busy.push_back({o.off, o.off + o.bytes});
      std::sort(busy.begin(), busy.end());
      size_t off = 0;
      for (auto& b : busy) {
        if (off + it.bytes <= b.first) break;
        off = std::max(off, b.second);
      }
      it.off = off; it.placed = true;
      total = std::max(total, off + it.bytes);
    }
    uint8_t* base = nullptr;
    if (share) {
      SharedPool& p = shared_pools()[device];
      if (!p.base) {
        // sized for the larger net of this micro-batch up front (the ESF-Net with the edge branch: ~145 MB per frame)
        p.bytes = std::max(total, (size_t)mb * 150u * 1000u * 1000u);
        CUDA_OK(cudaMalloc(&p.base, p.bytes));
        CUDA_OK(cudaMemset(p.base, 0, p.bytes));
      }
      if (p.bytes >= total) {
        base = (uint8_t*)p.base;
        if (!uses_pool) { ++p.refs; uses_pool = true; }
      }
    }
    if (!base) base = (uint8_t*)mem.alloc(total);        // zero-filled
    for (size_t i = 0; i < pending_acts.size(); ++i) {
      pending_acts[i].act->hi = (bf16*)(base + items[2 * i].off);
      pending_acts[i].act->lo = (bf16*)(base + items[2 * i + 1].off);
    }
    arena_bytes += total;
    pending_acts.clear();
    (void)share;
  }

  // ---------------------------------------------------------------------------------------
  // Packs a reference conv weight [cout][cin][kh][kw] for the (pieces -> K chunk) layout.
  // extra groups (MSBlock tail) are packed by passing several weights with their own dilation.
  struct PackSpec {
    const float* w; const float* bias; int dil, pad;
  };

  void build_conv(ConvLayer& L, DevMem& mem, const std::string& name, const std::vector<Piece>& pieces,
                  const std::vector<PackSpec>& specs, int cout, int cin, int kh, int kw, int H, int W,
                  int batch, bool round_robin_groups = false, std::vector<float>* w32_out = nullptr) {
    L.name = name;
    L.cout = cout;
    ConvGeom& g = L.g;
    g.H = H; g.W = W; g.batch = batch;
    g.groups = (int)specs.size();
    g.ntaps = kh * kw * g.groups;
    EGN_CHECK(g.ntaps <= EGN_MAX_TAPS, name + ": too many taps");
    g.cout_pad = round_up(cout, 16);
    // ---- views / chunks
    struct ViewSpan { const Act* buf; int n_off, c_lo, c_hi, first_chunk, src; };
    std::vector<ViewSpan> views;
    std::vector<int> piece_view(pieces.size());
    int ref_total = 0;
    for (size_t i = 0; i < pieces.size(); ++i) {
      const Piece& pc = pieces[i];
      ref_total += pc.len;
      if (!views.empty() && views.back().buf == pc.buf && views.back().n_off == pc.n_off &&
          pc.buf_c >= views.back().c_lo) {
        views.back().c_hi = std::max(views.back().c_hi, pc.buf_c + pc.len);
      } else {
        views.push_back({pc.buf, pc.n_off, pc.buf_c, pc.buf_c + pc.len, 0, 0});
      }
      piece_view[i] = (int)views.size() - 1;
    }
    EGN_CHECK(ref_total == cin, name + ": pieces do not cover cin");
    L.nsrc = 0;
    int nch = 0;
    for (auto& v : views) {
      int s = -1;
      for (int k = 0; k < L.nsrc; ++k) if (L.src_act[k] == v.buf) s = k;
      if (s < 0) {
        EGN_CHECK(L.nsrc < EGN_MAX_SRC, name + ": too many source buffers");
        s = L.nsrc++;
        L.src_act[s] = v.buf;
        L.src[s].hi = v.buf->hi; L.src[s].lo = v.buf->lo; L.src[s].C = v.buf->C; L.src[s].N = v.buf->N;
      }
      v.src = s;
      v.first_chunk = nch;
      const int n = ceil_div(v.c_hi - v.c_lo, EGN_KC);
      for (int k = 0; k < n; ++k) {
        EGN_CHECK(nch < EGN_MAX_CHUNKS, name + ": too many K chunks");
        g.chunk_src[nch] = (uint8_t)s;
        g.chunk_c0[nch] = (int16_t)(v.c_lo + k * EGN_KC);
        g.chunk_noff[nch] = v.n_off;
        ++nch;
      }
    }
    g.nchunks = nch;
    g.kpad = nch * EGN_KC;
    // ---- taps, grouped by horizontal offset: the tensor-core kernel loads one activation box per
    // distinct dx and serves every dy of it from that box
    struct TapRef { int gi, r, s, dy, dx; };
    std::vector<TapRef> taps;
    for (int gi = 0; gi < g.groups; ++gi)
      for (int r = 0; r < kh; ++r)
        for (int s = 0; s < kw; ++s)
          taps.push_back({gi, r, s, r * specs[gi].dil - specs[gi].pad, s * specs[gi].dil - specs[gi].pad});
    std::stable_sort(taps.begin(), taps.end(), [](const TapRef& a, const TapRef& b) { return a.dx < b.dx; });
    if (round_robin_groups) {
      // (phase-lattice MSBlock tail: one shared box serves every tap, so the order is free) group 0's k-th tap, group 1's
      // k-th tap, group 2's k-th tap, ...: consecutive taps accumulate into different TMEM columns
      std::vector<std::vector<TapRef>> per(g.groups);
      for (const TapRef& t : taps) per[t.gi].push_back(t);
      taps.clear();
      for (int k = 0; k < kh * kw; ++k)
        for (int gi = 0; gi < g.groups; ++gi) taps.push_back(per[gi][k]);
    }
    for (int t = 0; t < g.ntaps; ++t) {
      g.tap_dy[t] = (int8_t)taps[t].dy;
      g.tap_dx[t] = (int8_t)taps[t].dx;
      g.tap_grp[t] = (int8_t)taps[t].gi;
    }
    // ---- weights
    const size_t wn = (size_t)g.ntaps * g.cout_pad * g.kpad;
    std::vector<bf16> whi(wn, host_bf16(0.f)), wlo(wn, host_bf16(0.f));
    std::vector<float> bias((size_t)g.groups * g.cout_pad, 0.f);
    if (w32_out) w32_out->assign(wn, 0.f);
    {
      int ref_c = 0;
      for (size_t i = 0; i < pieces.size(); ++i) {
        const Piece& pc = pieces[i];
        const ViewSpan& v = views[piece_view[i]];
        for (int j = 0; j < pc.len; ++j, ++ref_c) {
          const int rel = pc.buf_c - v.c_lo + j;
          const int kidx = (v.first_chunk + rel / EGN_KC) * EGN_KC + rel % EGN_KC;
          for (int t = 0; t < g.ntaps; ++t) {
            const TapRef& tr = taps[t];
            for (int co = 0; co < cout; ++co) {
              const float wv = specs[tr.gi].w[(((size_t)co * cin + ref_c) * kh + tr.r) * kw + tr.s];
              const size_t o = ((size_t)t * g.cout_pad + co) * g.kpad + kidx;
              const bf16 h = host_bf16(wv);
              whi[o] = h;
              wlo[o] = host_bf16(wv - __bfloat162float(h));
              if (w32_out) (*w32_out)[o] = wv;
            }
          }
        }
      }
    }
    for (int gi = 0; gi < g.groups; ++gi)
      if (specs[gi].bias)
        for (int co = 0; co < cout; ++co) bias[(size_t)gi * g.cout_pad + co] = specs[gi].bias[co];
    L.w_hi = mem.upload(whi);
    L.w_lo = mem.upload(wlo);
    L.e.bias = mem.upload(bias);
    L.flops = 2.0 * cout * cin * kh * kw * g.groups * H * W;
  }

  void set_store_epilogue(ConvLayer& L, const Act* dst, int coff, int act, const float* post_scale = nullptr,
                          const float* post_shift = nullptr) {
    L.e.mode = CONV_STORE;
    L.e.act = act;
    L.e.cout_store = L.g.cout_pad;           // whole 16-channel groups: 32-byte stores in the epilogue
    L.e.out_hi = dst->hi; L.e.out_lo = dst->lo; L.e.out_C = dst->C; L.e.out_coff = coff;
    L.e.post_scale = post_scale; L.e.post_shift = post_shift;
    EGN_CHECK(dst->H == L.g.H && dst->W == L.g.W, L.name + ": destination size mismatch");
    EGN_CHECK(coff % 16 == 0 && dst->C % 16 == 0 && coff + L.e.cout_store <= dst->C, L.name + ": destination channel window");
  }

  void set_stats(ConvLayer& L, double* stats, int stats_C, int stats_coff) {
    L.e.stats = stats; L.e.stats_C = stats_C; L.e.stats_coff = stats_coff;
  }

  // EGN_PRODUCTS="prefix=mode,prefix=mode,...": layers whose name starts with `prefix` run with mode 1 (hi*hi + hi*lo:
  // activations act as bf16) or 2 (hi*hi + lo*hi: weights act as bf16).  Precision probe only (tools/gpu_precision_probe.py);
  // egn_info reports how many layers are lowered and bench.py refuses to time such a context as the parity engine.
  static int product_policy(const std::string& name) {
    const char* e = getenv("EGN_PRODUCTS");
    if (!e) return 0;
    std::string s(e);
    size_t pos = 0;
    while (pos < s.size()) {
      size_t end = s.find(',', pos);
      if (end == std::string::npos) end = s.size();
      const std::string item = s.substr(pos, end - pos);
      const size_t eq = item.find('=');
      if (eq != std::string::npos && name.compare(0, eq, item, 0, eq) == 0) return atoi(item.c_str() + eq + 1);
      pos = end + 1;
    }
    return 0;
  }

  void finalize_conv(ConvLayer& L) {
    conv_index[L.name] = &L;
    L.prod_mode = nsplit == 3 ? product_policy(L.name) : 0;
    EGN_CHECK(!(L.phase && L.prod_mode == 2), L.name + ": the phase-lattice tail is wide-only (EGN_PRODUCTS mode 2 unsupported there)");
    L.simt.prod_mode = L.prod_mode; L.tc.prod_mode = L.prod_mode;
    // SIMT companion
    L.simt.g = L.g; L.simt.e = L.e; L.simt.nsplit = nsplit;
    L.simt.w_hi = L.w_hi; L.simt.w_lo = L.w_lo;
    for (int s = 0; s < EGN_MAX_SRC; ++s) L.simt.src[s] = L.src[s];
    // tensor-core parameters
    L.tc.g = L.g; L.tc.e = L.e; L.tc.err_flag = err_flag; L.tc.timing = nullptr;
    L.tc.w_frames = L.w_frames;
    if (use_tc) {
      if (L.phase) {
        // same taps and weights on the polyphase lattice: H/d x W/d "frames", d*d of them per image
        const int d = L.phase;
        ConvGeom& tg = L.tc.g;
        EGN_CHECK(tg.H % d == 0 && tg.W % d == 0, L.name + ": phase lattice needs divisible sizes");
        tg.phase = d; tg.H /= d; tg.W /= d; tg.batch *= d * d;
        for (int t = 0; t < tg.ntaps; ++t) {
          EGN_CHECK(tg.tap_dy[t] % d == 0 && tg.tap_dx[t] % d == 0, L.name + ": tap offsets must be multiples of the phase");
          tg.tap_dy[t] = (int8_t)(tg.tap_dy[t] / d); tg.tap_dx[t] = (int8_t)(tg.tap_dx[t] / d);
        }
      }
      tc_configure(L.tc, L.g.cout_pad, nsplit);
      for (int s = 0; s < L.nsrc; ++s) L.tc.a_C[s] = L.src_act[s]->C;
      for (int s = 0; s < L.nsrc; ++s) {
        const Act* a = L.src_act[s];
        if (L.phase) {
          const bool p64 = a->C != 32;      // o as a 32-channel window of a wider [f | o] buffer: 64-byte promotion
          make_act_map_phase(&L.tc.a_map[0][s], a->hi, a->N, a->H, a->W, a->C, L.phase, L.tc.box_w, L.tc.box_rows, p64);
          make_act_map_phase(&L.tc.a_map[1][s], a->lo, a->N, a->H, a->W, a->C, L.phase, L.tc.box_w, L.tc.box_rows, p64);
          continue;
        }
        // channel window of this source the layer reads: [lo, hi) in 64-byte chunks
        int lo = a->C, hi = 0;
        for (int c = 0; c < L.g.nchunks; ++c)
          if (L.g.chunk_src[c] == s) { lo = std::min(lo, (int)L.g.chunk_c0[c]); hi = std::max(hi, (int)L.g.chunk_c0[c] + EGN_KC); }
        hi = std::min(hi, a->C);
        // (only where the operands stream from HBM: at 60x80 and below the buffers are L2-resident and
        //  the wider promotion is the faster one - measured on enc.down_block3.conv21/conv31)
        // (a pixel of 96 or 160 channels is 192 / 320 bytes: 128-byte lines straddle pixels there, so a window that is
        //  narrower than the pixel would drag the neighbouring channels in as well - measured on the merged [f | o]
        //  buffers: msblock1_2.conv read 29.6 MB/frame for 19.7 MB of operands)
        const bool partial = !(lo == 0 && hi == a->C);
        const bool promo64 = a->H * a->W >= 120 * 160 &&
                             ((lo * 2) % 128 != 0 || ((hi * 2) % 128 != 0 && hi != a->C) || (partial && (a->C * 2) % 128 != 0));
        make_act_map(&L.tc.a_map[0][s], a->hi, a->N, a->H, a->W, a->C, L.tc.box_w, L.tc.box_rows, promo64);
        make_act_map(&L.tc.a_map[1][s], a->lo, a->N, a->H, a->W, a->C, L.tc.box_w, L.tc.box_rows, promo64);
      }
      for (int s = L.nsrc; s < EGN_MAX_SRC; ++s) {
        L.tc.a_map[0][s] = L.tc.a_map[0][0];
        L.tc.a_map[1][s] = L.tc.a_map[1][0];
      }
      if (L.w_frames) {
        make_w_map_frames(&L.tc.w_map[0], L.fw_hi, L.w_frames, L.g.ntaps * L.g.cout_pad, L.g.kpad, L.tc.n_tile);
        make_w_map_frames(&L.tc.w_map[1], L.fw_lo, L.w_frames, L.g.ntaps * L.g.cout_pad, L.g.kpad, L.tc.n_tile);
      } else {
        make_w_map(&L.tc.w_map[0], L.w_hi, L.g.ntaps * L.g.cout_pad, L.g.kpad, L.tc.n_tile);
        make_w_map(&L.tc.w_map[1], L.w_lo, L.g.ntaps * L.g.cout_pad, L.g.kpad, L.tc.n_tile);
      }
    }
  }

  void run_conv(ConvLayer& L, int batch, cudaStream_t st, int noff = -1) {
    if (profiling) {
      if (prof_used >= 8192) profile_resolve();
      profile_begin(st);
    }
    run_conv_impl(L, batch, st, noff);
    if (profiling) profile_end(st, L.flops * batch, &L, nullptr, batch);
  }

  // noff >= 0 re-bases every K chunk that reads a frame-offset window (the edge frames of the
  // shared encoder sit at [nb, 2nb) of its buffers, and nb shrinks on a ragged last micro-batch)
  void run_conv_impl(ConvLayer& L, int batch, cudaStream_t st, int noff = -1) {
    if (use_tc) {
      TcParams p = L.tc;
      p.g.batch = batch * (L.phase ? L.phase * L.phase : 1);
      if (noff >= 0)
        for (int c = 0; c < p.g.nchunks; ++c) if (p.g.chunk_noff[c]) p.g.chunk_noff[c] = noff;
      p.total_tiles = p.tiles_x * p.tiles_y * p.g.batch * p.n_blocks;
      tc_launch(p, num_sms, st);
      ++launches;
    } else {
      SimtParams p = L.simt;
      p.g.batch = batch;
      if (noff >= 0)
        for (int c = 0; c < p.g.nchunks; ++c) if (p.g.chunk_noff[c]) p.g.chunk_noff[c] = noff;
      simt_launch(p, st);
      ++launches;
    }
  }

  long long launches = 0;

  // ---- optional per-launch timing of the convolution kernel (bench.py roofline): CUDA event pairs
  // around every conv launch on the launching stream, resolved lazily by profile_read().
  bool profiling = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events;
  struct ProfTag { ConvLayer* layer; const char* aux; int frames; };
  std::vector<ProfTag> prof_tags;
  std::map<std::string, std::pair<double, long long>> prof_aux;   // name -> (ms, launches)
  size_t prof_used = 0;
  double prof_flops = 0, prof_ms = 0;
  long long prof_launches = 0;

  void profile_begin(cudaStream_t st) {
    if (prof_used == prof_events.size()) {
      cudaEvent_t a, b;
      CUDA_OK(cudaEventCreate(&a)); CUDA_OK(cudaEventCreate(&b));
      prof_events.push_back({a, b});
    }
    CUDA_OK(cudaEventRecord(prof_events[prof_used].first, st));
  }
  void profile_end(cudaStream_t st, double flops, ConvLayer* layer, const char* aux, int frames) {
    CUDA_OK(cudaEventRecord(prof_events[prof_used].second, st));
    if (prof_tags.size() <= prof_used) prof_tags.resize(prof_used + 1);
    prof_tags[prof_used] = {layer, aux, frames};
    ++prof_used;
    if (layer) { prof_flops += flops; ++prof_launches; }
  }
  // brackets a non-conv launch sequence
  template <typename F>
  void aux(const char* name, cudaStream_t st, F&& fn) {
    if (profiling) {
      if (prof_used >= 8192) profile_resolve();
      profile_begin(st);
    }
    fn();
    if (profiling) profile_end(st, 0, nullptr, name, 0);
  }
  void profile_resolve() {
    for (size_t i = 0; i < prof_used; ++i) {
      CUDA_OK(cudaEventSynchronize(prof_events[i].second));
      float ms = 0;
      CUDA_OK(cudaEventElapsedTime(&ms, prof_events[i].first, prof_events[i].second));
      const ProfTag& t = prof_tags[i];
      if (t.layer) {
        prof_ms += ms;
        t.layer->prof_ms += ms; t.layer->prof_frames += t.frames; ++t.layer->prof_n;
      } else {
        auto& a = prof_aux[t.aux];
        a.first += ms; ++a.second;
      }
    }
    prof_used = 0;
  }

  // ======================================================================================= BDCN
  void build_bdcn() {
    EGN_CHECK(has_bdcn, "BDCN weights not set");
    EGN_CHECK(mb > 0, "egn_plan must be called before the first forward");
    const StateDict& sd = sd_bdcn;
    DevMem& mem = mem_bdcn;
    static const char* names[13] = {"conv1_1", "conv1_2", "conv2_1", "conv2_2", "conv3_1", "conv3_2", "conv3_3",
                                    "conv4_1", "conv4_2", "conv4_3", "conv5_1", "conv5_2", "conv5_3"};
    static const int cin[13] = {3, 64, 64, 128, 128, 256, 256, 256, 512, 512, 512, 512, 512};
    static const int cout[13] = {64, 64, 128, 128, 256, 256, 256, 512, 512, 512, 512, 512, 512};
    static const int stage_of[13] = {0, 0, 1, 1, 2, 2, 2, 3, 3, 3, 4, 4, 4};
    const int sh[5] = {240, 120, 60, 30, 29}, sw_[5] = {320, 160, 80, 40, 39};
    // msblock{s}_{j}.conv and features.conv{s}_{j+1} are both 3x3 / pad 1 / ReLU convolutions of the SAME map
    // (bdcn_new.py:120-160, vgg16_c.py:65-88): where the dilations agree (stages 1-4) they run as ONE launch with
    // the weight rows stacked, N = Cout_vgg + 32, writing [f | o] into one buffer whose channel windows the
    // consumers read - the narrow N = 32 launch and one full read of the shared map disappear.
    // EGN_MS_MERGE: bit s enables stage s + 1 (default: stages 1-3; 512 + 32 does not split into equal 16-aligned N tiles).
    {
      const int mask = getenv("EGN_MS_MERGE") ? atoi(getenv("EGN_MS_MERGE")) : 0x7;
      for (int i = 0; i < 13; ++i) {
        const bool next_same_stage = i + 1 < 13 && stage_of[i + 1] == stage_of[i];
        bd.merged[i] = use_tc && next_same_stage && stage_of[i] < 3 && ((mask >> stage_of[i]) & 1);
        bd.f_c[i] = cout[i];
      }
    }
    // schedule ticks (bdcn_forward): iteration i of the layer loop owns [10 i, 10 i + 9]; f[i] is written at 10 i (or
    // during iteration i - 1 when it rides with the merged launch) and last read by iteration i + 1's convolution
    for (int i = 0; i < 13; ++i) {
      const int s = stage_of[i];
      bd.f[i] = new_act(mem, mb, sh[s], sw_[s], cout[i] + ((i > 0 && bd.merged[i - 1]) ? 32 : 0), 10 * (i - 1), 10 * (i + 1) + 9);
      debug_acts[std::string("features.") + names[i]] = {bd.f[i], {0, cout[i]}};
    }
    const int pc[4] = {64, 128, 256, 512};
    static const int pool_src[4] = {1, 3, 6, 9};
    for (int i = 0; i < 4; ++i)
      bd.pool[i] = new_act(mem, mb, sh[i + 1], sw_[i + 1], pc[i], 10 * (pool_src[i] - 1), 10 * (pool_src[i] + 1) + 9);
    static const int stage_first[5] = {0, 2, 4, 7, 10}, stage_last[5] = {1, 3, 6, 9, 12};
    for (int s = 0; s < 5; ++s) {
      bd.o[s] = new_act(mem, mb, sh[s], sw_[s], 32, 10 * stage_first[s], 10 * stage_last[s] + 9);
      bd.h[s] = sh[s]; bd.w[s] = sw_[s];
      bd.score[s] = (float*)mem.alloc((size_t)mb * sh[s] * sw_[s] * 2 * sizeof(float));
      debug_f32["score" + std::to_string(s + 1)] = {bd.score[s], {sh[s], sw_[s], 2}};
    }
    commit_acts(mem);
    // conv1_1 on cat(img,img,img): sum the three input-channel slices (utils.py:649)
    {
      const HostTensor& w = sd_get(sd, "features.conv1_1.weight");
      const HostTensor& b = sd_get(sd, "features.conv1_1.bias");
      std::vector<float> wp(9 * 64);
      for (int co = 0; co < 64; ++co)
        for (int t = 0; t < 9; ++t) {
          double s = 0;
          for (int ci = 0; ci < 3; ++ci) s += w.data[((size_t)co * 3 + ci) * 9 + t];
          wp[(size_t)t * 64 + co] = (float)s;
        }
      std::vector<float> w3(3 * 9 * 64);
      for (int co = 0; co < 64; ++co)
        for (int ci = 0; ci < 3; ++ci)
          for (int t = 0; t < 9; ++t) w3[((size_t)ci * 9 + t) * 64 + co] = w.data[((size_t)co * 3 + ci) * 9 + t];
      bd.first_w_gray = mem.upload(wp);
      bd.first_w_rgb = mem.upload(w3);
      bd.first.w = bd.first_w_gray;
      bd.first.bias = mem.upload(b.data);
      bd.first.dst = make_view(*bd.f[0], 0);
      bd.first.H = 240; bd.first.W = 320; bd.first.cout = 64; bd.first.act = ACT_RELU;
    }
    for (int i = 1; i < 13; ++i) {
      const int s = stage_of[i];
      const bool after_pool = (i == 2 || i == 4 || i == 7 || i == 10);
      const Act* in = after_pool ? bd.pool[s - 1] : bd.f[i - 1];
      const int dil = s == 4 ? 2 : 1;
      const HostTensor& w = sd_get(sd, std::string("features.") + names[i] + ".weight");
      const HostTensor& b = sd_get(sd, std::string("features.") + names[i] + ".bias");
      if (bd.merged[i - 1]) {
        // rows [0, cout) = features.conv, rows [cout, cout + 32) = msblock.conv of the map both read
        static const int blk_of[13] = {1, 2, 1, 2, 1, 2, 3, 1, 2, 3, 1, 2, 3};
        const std::string mp = "msblock" + std::to_string(stage_of[i - 1] + 1) + "_" + std::to_string(blk_of[i - 1]);
        const HostTensor& w0 = sd_get(sd, mp + ".conv.weight");
        const HostTensor& b0 = sd_get(sd, mp + ".conv.bias");
        EGN_CHECK(w0.shape[1] == cin[i] && w0.shape[0] == 32, mp + ".conv: unexpected weight shape");
        std::vector<float> wc(w.data), bc(b.data);
        wc.insert(wc.end(), w0.data.begin(), w0.data.end());
        bc.insert(bc.end(), b0.data.begin(), b0.data.end());
        build_conv(bd.vgg[i], mem, std::string("features.") + names[i] + "+" + mp + ".conv", {{in, 0, cin[i], 0}},
                   {{wc.data(), bc.data(), dil, dil}}, cout[i] + 32, cin[i], 3, 3, sh[s], sw_[s], mb);
        set_store_epilogue(bd.vgg[i], bd.f[i], 0, ACT_RELU);
        finalize_conv(bd.vgg[i]);
        conv_alias[std::string("features.") + names[i]] = &bd.vgg[i];
        conv_alias[mp + ".conv"] = &bd.vgg[i];
        continue;
      }
      build_conv(bd.vgg[i], mem, std::string("features.") + names[i], {{in, 0, cin[i], 0}},
                 {{w.data.data(), b.data.data(), dil, dil}}, cout[i], cin[i], 3, 3, sh[s], sw_[s], mb);
      set_store_epilogue(bd.vgg[i], bd.f[i], 0, ACT_RELU);
      finalize_conv(bd.vgg[i]);
    }
    // pools 1-3 (2x2 / stride 2 on even sizes) ride in the producing convolution's epilogue; pool4 (stride 1) stays a kernel
    {
      static const int src_of_pool[3] = {1, 3, 6};
      const bool fuse = use_tc && !getenv("EGN_NO_POOL_FUSE");
      for (int k = 0; k < 4; ++k) bd.pool_fused[k] = false;
      for (int k = 0; k < 3 && fuse; ++k) {
        ConvLayer& L = bd.vgg[src_of_pool[k]];
        L.tc.e.pool_hi = bd.pool[k]->hi; L.tc.e.pool_lo = bd.pool[k]->lo; L.tc.e.pool_C = bd.pool[k]->C; L.tc.e.pool_ch = cout[src_of_pool[k]];
        bd.pool_fused[k] = true;
      }
    }
    // MSBlocks + collapsed side chain
    static const int nblk[5] = {2, 2, 3, 3, 3};
    const HostTensor& fw = sd_get(sd, "fuse.weight");
    const HostTensor& fb = sd_get(sd, "fuse.bias");
    int fi = 0;
    for (int s = 0; s < 5; ++s) {
      const std::string S = std::to_string(s + 1);
      const HostTensor& wsA = sd_get(sd, "score_dsn" + S + ".weight");
      const HostTensor& bsA = sd_get(sd, "score_dsn" + S + ".bias");
      const HostTensor& wsB = sd_get(sd, "score_dsn" + S + "_1.weight");
      const HostTensor& bsB = sd_get(sd, "score_dsn" + S + "_1.bias");
      double cA = bsA.data[0], cB = bsB.data[0];
      for (int j = 0; j < nblk[s]; ++j, ++fi) {
        const std::string J = std::to_string(j + 1);
        const std::string mp = "msblock" + S + "_" + J;
        const HostTensor& w0 = sd_get(sd, mp + ".conv.weight");
        const HostTensor& b0 = sd_get(sd, mp + ".conv.bias");
        ConvLayer& Lin = bd.ms_in[fi];
        // `o` = relu(conv(f)): its own launch into the stage's o buffer, or (merged) channels [cout, cout + 32) of f[fi + 1]
        const Act* o_buf = bd.merged[fi] ? bd.f[fi + 1] : bd.o[s];
        const int o_off = bd.merged[fi] ? cout[fi + 1] : 0;
        if (!bd.merged[fi]) {
          build_conv(Lin, mem, mp + ".conv", {{bd.f[fi], 0, cout[fi], 0}}, {{w0.data.data(), b0.data.data(), 1, 1}},
                     32, cout[fi], 3, 3, sh[s], sw_[s], mb);
          set_store_epilogue(Lin, bd.o[s], 0, ACT_RELU);
          finalize_conv(Lin);
        }
        ConvLayer& Lt = bd.ms_tail[fi];
        std::vector<PackSpec> specs;
        for (int d = 1; d <= 3; ++d) {
          const HostTensor& wd = sd_get(sd, mp + ".conv" + std::to_string(d) + ".weight");
          const HostTensor& bdv = sd_get(sd, mp + ".conv" + std::to_string(d) + ".bias");
          specs.push_back({wd.data.data(), bdv.data.data(), 4 * d, 4 * d});
        }
        static const int tail_phase = getenv("EGN_TAIL_PHASE") ? atoi(getenv("EGN_TAIL_PHASE")) : 4;
        const int ph = (use_tc && nsplit == 3 && s <= 1 && tail_phase > 1) ? tail_phase : 0;
        build_conv(Lt, mem, mp + ".tail", {{o_buf, o_off, 32, 0}}, specs, 32, 32, 3, 3, sh[s], sw_[s], mb, ph != 0);
        // collapsed conv{s}_{j}_down -> score_dsn{s}, score_dsn{s}_1 (both linear)
        const HostTensor& wd = sd_get(sd, "conv" + S + "_" + J + "_down.weight");   // [21][32]
        const HostTensor& bdn = sd_get(sd, "conv" + S + "_" + J + "_down.bias");
        std::vector<float> swv(64);
        for (int c = 0; c < 32; ++c) {
          double a = 0, b = 0;
          for (int m = 0; m < 21; ++m) {
            a += (double)wsA.data[m] * wd.data[(size_t)m * 32 + c];
            b += (double)wsB.data[m] * wd.data[(size_t)m * 32 + c];
          }
          swv[c] = (float)a; swv[32 + c] = (float)b;
        }
        for (int m = 0; m < 21; ++m) { cA += (double)wsA.data[m] * bdn.data[m]; cB += (double)wsB.data[m] * bdn.data[m]; }
        Lt.e.mode = CONV_MSBLOCK;
        Lt.e.act = ACT_RELU;
        Lt.e.cout_store = 32;
        Lt.e.o_hi = o_buf->hi; Lt.e.o_lo = o_buf->lo; Lt.e.o_C = o_buf->C; Lt.e.o_coff = o_off;
        Lt.e.score_w = mem.upload(swv);
        Lt.e.score = bd.score[s];
        Lt.e.score_accum = j > 0;
        Lt.flops += 2.0 * 32 * 21 * sh[s] * sw_[s];      // conv_down (the reference's 1x1)
        // stages 1-2 (240x320, 120x160): run on the 4x4 polyphase lattice, where the three dilations are 1/2/3, one
        // shared activation box serves all 27 taps and the weights stay resident (conv_tc.cuh)
        Lt.phase = ph;
        finalize_conv(Lt);
      }
      bd.tail.cA[s] = (float)cA; bd.tail.cB[s] = (float)cB;
      double al = 0, be = 0;
      for (int j = s; j < 5; ++j) al += fw.data[j];          // s_k appears in p_j_1 for j >= k
      for (int j = 0; j <= s; ++j) be += fw.data[5 + j];     // s_k1 appears in p_j_2 for j <= k
      bd.tail.alpha[s] = (float)al; bd.tail.beta[s] = (float)be;
      bd.tail.score[s] = bd.score[s];
      bd.tail.h[s] = sh[s]; bd.tail.w[s] = sw_[s];
    }
    static const char* upn[5] = {"", "upsample_2", "upsample_4", "upsample_8", "upsample_8_5"};
    static const int upk[5] = {0, 4, 8, 16, 16}, ups[5] = {1, 2, 4, 8, 8}, upc[5] = {0, 1, 2, 4, 0};
    bd.tail.kern[0] = nullptr; bd.tail.K[0] = 0; bd.tail.stride[0] = 1; bd.tail.crop[0] = 0; bd.tail.shift[0] = 0;
    for (int s = 1; s < 5; ++s) {
      const HostTensor& k = sd_get(sd, std::string(upn[s]) + ".weight");
      EGN_CHECK(k.numel() == upk[s] * upk[s], "upsample kernel size");
      bd.tail.kern[s] = mem.upload(k.data);
      bd.tail.K[s] = upk[s]; bd.tail.stride[s] = ups[s]; bd.tail.crop[s] = upc[s];
      bd.tail.shift[s] = ups[s] == 2 ? 1 : ups[s] == 4 ? 2 : 3;
      EGN_CHECK((1 << bd.tail.shift[s]) == ups[s], "upsampler strides are powers of two");
    }
    bd.tail.fuse_bias = fb.data[0];
    bd.tail.H = 240; bd.tail.W = 320;
    built_bdcn = true;
  }

  void maxpool(const Act* src, const Act* dst, int stride, int batch, cudaStream_t st) {
    PoolParams p;
    p.src = make_view(*src, 0); p.dst = make_view(*dst, 0);
    p.B = batch; p.Hi = src->H; p.Wi = src->W; p.Ho = dst->H; p.Wo = dst->W; p.stride = stride; p.Cv = dst->C;   // a merged source carries 32 more channels
    launch_1d(maxpool_kernel, p, (long long)batch * p.Ho * p.Wo * (p.Cv / 8), st);
    ++launches;
  }

  // x: device fp32 [B][planes][H][W]; planes == 1 is the grey frame the reference replicates with
  // cat(img,img,img) (utils.py:649; conv1_1 then uses input-channel-summed weights), planes == 3 is
  // the general BDCN.forward input.  edge_out: [B][H][W].
  // side (optional): [10][B][H][W], the per-scale sigmoids of bdcn_new.py:178-191 in the reference's order.
  void bdcn_forward(const float* x, int planes, float* edge_out, int B, cudaStream_t st, float* side = nullptr) {
    if (!built_bdcn) build_bdcn();
    const size_t hw = (size_t)EGN_H * EGN_W;
    for (int b0 = 0; b0 < B; b0 += mb) {
      const int nb = std::min(mb, B - b0);
      FirstConvParams fp = bd.first;
      fp.B = nb; fp.cin = planes;
      fp.w = planes == 1 ? bd.first_w_gray : bd.first_w_rgb;
      for (int c = 0; c < planes; ++c) { fp.in[c] = x + ((size_t)b0 * planes + c) * hw; fp.fstride[c] = (long long)planes * hw; }
      aux("bdcn.first_conv", st, [&] { launch_first(fp, st); });
      static const int pool_after[13] = {-1, 0, -1, 1, -1, -1, 2, -1, -1, 3, -1, -1, -1};
      // the MSBlock branch (standalone `conv` launches and all tails, in block order: the tails of a stage accumulate
      // into one score map) runs on the side stream when there is one; `ss == st` otherwise
      const cudaStream_t ss = side_of(st);
      side_used = 0;
      for (int i = 0; i < 13; ++i) {
        if (i > 0 && !bd.merged[i - 1]) run_conv(bd.vgg[i], nb, st);
        if (bd.merged[i]) run_conv(bd.vgg[i + 1], nb, st);       // features.conv(i+1) + msblock(i).conv in one launch
        if (ss != st) stream_edge(st, ss);                       // f[i] (and a merged o) are complete
        if (!bd.merged[i]) run_conv(bd.ms_in[i], nb, ss);
        run_conv(bd.ms_tail[i], nb, ss);
        if (pool_after[i] >= 0 && !bd.pool_fused[pool_after[i]])
          aux("bdcn.maxpool", st, [&] { maxpool(bd.f[i], bd.pool[pool_after[i]], pool_after[i] == 3 ? 1 : 2, nb, st); });
      }
      if (ss != st) stream_edge(ss, st);                         // the score maps are complete
      BdcnTailParams tp = bd.tail;
      tp.N = nb; tp.out = edge_out + b0 * hw;
      tp.side = side ? side + b0 * hw : nullptr; tp.side_stride = (long long)B * (long long)hw;
      aux("bdcn.tail", st, [&] { launch_1d(bdcn_tail_kernel, tp, (long long)nb * hw, st); ++launches; });
    }
  }

  // ======================================================================================= ESF
  static std::vector<float> bn_scale(const StateDict& sd, const std::string& p, int c, int pad, std::vector<float>& shift) {
    const HostTensor& w = sd_get(sd, p + ".weight");
    const HostTensor& b = sd_get(sd, p + ".bias");
    const HostTensor& m = sd_get(sd, p + ".running_mean");
    const HostTensor& v = sd_get(sd, p + ".running_var");
    std::vector<float> scale(pad, 0.f);
    shift.assign(pad, 0.f);
    for (int i = 0; i < c; ++i) {
      const double s = (double)w.data[i] / std::sqrt((double)v.data[i] + 1e-5);
      scale[i] = (float)s;
      shift[i] = (float)((double)b.data[i] - (double)m.data[i] * s);
    }
    return scale;
  }

  // fp32 conv weights [co][ci][kh][kw] -> [kh][kw][ci][co]
  float* upload_hwio(DevMem& mem, const HostTensor& w) {
    const int co = (int)w.shape[0], ci = (int)w.shape[1], kh = (int)w.shape[2], kw = (int)w.shape[3];
    std::vector<float> o((size_t)co * ci * kh * kw);
    for (int a = 0; a < co; ++a)
      for (int b = 0; b < ci; ++b)
        for (int r = 0; r < kh; ++r)
          for (int s = 0; s < kw; ++s)
            o[(((size_t)r * kw + s) * ci + b) * co + a] = w.data[(((size_t)a * ci + b) * kh + r) * kw + s];
    return mem.upload(o);
  }

  void build_esf() {
    EGN_CHECK(has_esf, "ESF-Net weights not set");
    EGN_CHECK(mb > 0, "egn_plan must be called before the first forward");
    EGN_CHECK(cfg.input_concat + cfg.add_edge < 2, "edge can use only 1 time!");   // RITnet_v2.py:273
    const StateDict& sd = sd_esf;
    DevMem& mem = mem_esf;
    const int E = mb * (cfg.add_edge ? 2 : 1);
    es.E = E;
    const int inter[5] = {32, 64, 96, 128, 128}, in_c[5] = {32, 38, 76, 115, 153}, op_c[5] = {38, 76, 115, 153, 153};
    const int bh[5] = {240, 120, 60, 30, 15}, bw[5] = {320, 160, 80, 40, 20};
    static const char* bname[5] = {"enc.down_block1", "enc.down_block2", "enc.down_block3", "enc.down_block4", "enc.bottleneck"};
    // schedule ticks (esf_forward): head 1000-1001; encoder block i owns [1010 + 10 i, +9] (instnorm_x, conv1, conv21,
    // conv22, conv31, conv32, instnorm_td, TD.conv = +0 .. +7); up block j owns [1100 + 10 j, +9] (pre, conv11, conv12,
    // conv21, conv22 = +0 .. +4); final 1140-1141; style encoder 1150-1159; regression head 1160-1163
    auto TB = [](int i) { return 1010 + 10 * i; };
    auto TD_ = [](int j) { return 1100 + 10 * j; };
    es.h1 = new_act(mem, E, 240, 320, 32, 999, 1002);
    es.bt = new_act(mem, E, 15, 20, 160, TB(4) + 6, 1170);
    debug_acts["bt"] = {es.bt, {0, 153}};
    for (int i = 0; i < 5; ++i) {
      Block& b = es.blk[i];
      b.in_c = in_c[i]; b.in_pad = round_up(in_c[i], 16); b.inter = inter[i]; b.op_c = op_c[i];
      b.H = bh[i]; b.W = bw[i];
      b.off_out = 0; b.off_x = inter[i]; b.off_x1 = b.off_x + b.in_pad; b.off_x22 = b.off_x1 + inter[i];
      // [out | x | x1 | x22]: x arrives with the previous block's TD.conv (head.conv2 for block 1); the skip [out, x] is
      // last read by conv21 of up block 3 - i (the bottleneck's by its own TD)
      b.buf = new_act(mem, E, b.H, b.W, b.off_x22 + inter[i], i == 0 ? 1000 : TB(i - 1) + 6, i < 4 ? TD_(3 - i) + 5 : TB(4) + 8);
      // EXPERIMENT, off by default (EGN_IN_FOLD=<blocks>): fold the InstanceNorm in front of conv1 into per-frame weights
      // (aux.cuh in_fold_kernel) so that the normalisation pass (a full read + write of the map) disappears.  Measured
      // with EGN_IN_FOLD=2: -7.8 us/frame (+1.5 % frames/s), but the convolution then multiplies the UN-centred map
      // and the common mode costs operand bits: the ellipse parameters moved from 3.1e-3 to 1.6e-2 relative on the
      // 48-frame probe set (bar 1e-2, tools/gpu_precision_probe.py) - it does not meet the parity bars and stays off.
      static const int fold_blocks = getenv("EGN_IN_FOLD") ? atoi(getenv("EGN_IN_FOLD")) : 0;
      b.fold = use_tc && nsplit == 3 && i < fold_blocks;
      b.xn = b.fold ? nullptr : new_act(mem, E, b.H, b.W, b.in_pad, TB(i) - 1, TB(i) + 2);
      b.t = new_act(mem, E, b.H, b.W, inter[i], TB(i) + 1, TB(i) + 6);
      const bool pool = i < 4;
      b.tdin = new_act(mem, E, pool ? b.H / 2 : b.H, pool ? b.W / 2 : b.W, inter[i] + b.in_pad, TB(i) + 5, TB(i) + 8);
      b.stats_C = inter[i] + b.in_pad;
      const std::string P = bname[i];
      debug_acts[P + ".x"] = {b.buf, {b.off_x, b.in_c}};
      debug_acts[P + ".x1"] = {b.buf, {b.off_x1, b.inter}};
      debug_acts[P + ".x22"] = {b.buf, {b.off_x22, b.inter}};
      debug_acts[P + ".out"] = {b.buf, {b.off_out, b.inter}};
    }
    // InstanceNorm statistics come from the producing convolutions' epilogues (one array, one memset)
    {
      size_t tot = 0;
      for (int i = 0; i < 5; ++i) tot += (size_t)E * es.blk[i].stats_C * 2;
      es.stats_all = (double*)mem.alloc(tot * sizeof(double));
      es.stats_bytes = tot * sizeof(double);
      size_t off = 0;
      for (int i = 0; i < 5; ++i) { es.blk[i].stats = es.stats_all + off; off += (size_t)E * es.blk[i].stats_C * 2; }
    }
    // ---- the remaining activations (decoder, final block, regression head, style encoder) are declared here too, so
    // that the whole net is packed before any tensor map is encoded
    const int d_in[4] = {cfg.add_edge ? 306 : 153, cfg.add_edge ? 180 : 115, cfg.add_edge ? 100 : 76, cfg.add_edge ? 62 : 38};
    const int d_out[4] = {cfg.add_edge ? 180 : 115, cfg.add_edge ? 100 : 76, cfg.add_edge ? 62 : 38, 32};
    for (int i = 0; i < 4; ++i) {
      UpBlock& u = es.up[i];
      Block& sk = es.blk[3 - i];
      u.in_c = d_in[i]; u.out_c = d_out[i]; u.out_pad = round_up(d_out[i], 16);
      u.skip_c = sk.inter + sk.in_c; u.H = sk.H; u.W = sk.W;
      u.buf = new_act(mem, mb, u.H, u.W, u.out_pad, TD_(i) + 1, TD_(i) + 4);
      u.t = new_act(mem, mb, u.H, u.W, u.out_pad, TD_(i), TD_(i) + 5);
      u.out = new_act(mem, mb, u.H, u.W, u.out_pad, TD_(i) + 3, TD_(i) + 12);
      u.y = new_act(mem, mb, u.H / 2, u.W / 2, 2 * u.out_pad, TD_(i) - 1, TD_(i) + 4);
    }
    es.fin = new_act(mem, mb, 240, 320, 32, 1139, 1142);
    es.c1o = new_act(mem, mb, 15, 20, 128, 1159, 1164);
    // hin: channels [153, 160) (and [313, 320)) are never written and must stay zero -> never shared
    es.hin = cfg.add_seg ? new_act(mem, mb, 15, 20, cfg.add_edge ? 320 : 160) : nullptr;
    int st_gh[5], st_gw[5], st_cin[5];
    if (cfg.add_seg) {
      const int sc[5] = {64, 128, 256, 256, 256};
      int vh = 240, vw = 320;                      // valid region of the current layer's input
      for (int i = 0; i < 5; ++i) {
        if (i == 0) { st_gh[i] = vh + 6; st_gw[i] = vw; st_cin[i] = 32; }
        else { st_gh[i] = (vh + 2) / 2; st_gw[i] = (vw + 2) / 2; st_cin[i] = 4 * sc[i - 1]; }
        es.st_in[i] = new_act(mem, mb, st_gh[i], st_gw[i], st_cin[i], 1149, 1159);
        es.st_out[i] = new_act(mem, mb, st_gh[i], st_gw[i], sc[i], 1149, 1159);
        if (i > 0) { vh /= 2; vw /= 2; }
      }
    }
    commit_acts(mem);
    // head: conv1 (SIMT first layer) -> h1; conv2 + lrelu + BN -> block1.x   (utils.py:1046-1050)
    {
      const HostTensor& w = sd_get(sd, "enc.head.conv1.weight");
      const HostTensor& b = sd_get(sd, "enc.head.conv1.bias");
      const int ci = (int)w.shape[1];
      EGN_CHECK(ci == (cfg.input_concat ? 2 : 1), "enc.head.conv1 input channels do not match the setting");
      std::vector<float> wp((size_t)ci * 9 * 32);
      for (int co = 0; co < 32; ++co)
        for (int c = 0; c < ci; ++c)
          for (int t = 0; t < 9; ++t) wp[((size_t)c * 9 + t) * 32 + co] = w.data[((size_t)co * ci + c) * 9 + t];
      es.first.w = mem.upload(wp);
      es.first.bias = mem.upload(b.data);
      es.first.H = 240; es.first.W = 320; es.first.cout = 32; es.first.act = ACT_LRELU;
      const HostTensor& w2 = sd_get(sd, "enc.head.conv2.weight");
      const HostTensor& b2 = sd_get(sd, "enc.head.conv2.bias");
      build_conv(es.head2, mem, "enc.head.conv2", {{es.h1, 0, 32, 0}}, {{w2.data.data(), b2.data.data(), 1, 1}}, 32, 32,
                 3, 3, 240, 320, E);
      std::vector<float> shift;
      std::vector<float> scale = bn_scale(sd, "enc.head.bn", 32, es.head2.g.cout_pad, shift);
      set_store_epilogue(es.head2, es.blk[0].buf, es.blk[0].off_x, ACT_LRELU, mem.upload(scale), mem.upload(shift));
      set_stats(es.head2, es.blk[0].stats, es.blk[0].stats_C, es.blk[0].off_x);
      finalize_conv(es.head2);
    }
    for (int i = 0; i < 5; ++i) {
      Block& b = es.blk[i];
      const std::string P = bname[i];
      auto W = [&](const std::string& n) -> const HostTensor& { return sd_get(sd, P + "." + n + ".weight"); };
      auto Bv = [&](const std::string& n) -> const HostTensor& { return sd_get(sd, P + "." + n + ".bias"); };
      // conv1 on IN(x)                                          RITnet_v2.py:60
      if (b.fold) {
        std::vector<float> w32;
        build_conv(b.conv1, mem, P + ".conv1", {{b.buf, b.off_x, b.in_c, 0}}, {{W("conv1").data.data(), Bv("conv1").data.data(), 1, 1}},
                   b.inter, b.in_c, 3, 3, b.H, b.W, E, false, &w32);
        set_store_epilogue(b.conv1, b.buf, b.off_x1, ACT_LRELU);
        ConvLayer& L = b.conv1;
        const size_t per_frame = (size_t)L.g.ntaps * L.g.cout_pad * L.g.kpad;
        L.w_frames = E;
        L.fw_hi = (bf16*)mem.alloc(per_frame * E * sizeof(bf16));
        L.fw_lo = (bf16*)mem.alloc(per_frame * E * sizeof(bf16));
        float* bias_fc = (float*)mem.alloc((size_t)E * 9 * L.g.cout_pad * sizeof(float));
        L.e.bias_fc = bias_fc;
        InFoldParams& f = L.fold;
        f.w32 = mem.upload(w32); f.bias = L.e.bias; f.sums = b.stats; f.sums_C = b.stats_C; f.sums_coff = b.off_x;
        f.ntaps = L.g.ntaps; f.cout_pad = L.g.cout_pad; f.kpad = L.g.kpad; f.kreal = b.in_c; f.H = b.H; f.W = b.W;
        EGN_CHECK(L.g.ntaps == 9, P + ".conv1: fold expects a 3x3 layer");
        for (int t = 0; t < 9; ++t) { f.tap_dy[t] = L.g.tap_dy[t]; f.tap_dx[t] = L.g.tap_dx[t]; }
        f.w_hi = L.fw_hi; f.w_lo = L.fw_lo; f.bias_fc = bias_fc;
        finalize_conv(L);
      } else {
        build_conv(b.conv1, mem, P + ".conv1", {{b.xn, 0, b.in_c, 0}}, {{W("conv1").data.data(), Bv("conv1").data.data(), 1, 1}},
                   b.inter, b.in_c, 3, 3, b.H, b.W, E);
        set_store_epilogue(b.conv1, b.buf, b.off_x1, ACT_LRELU);
        finalize_conv(b.conv1);
      }
      // conv21 (1x1) on [x, x1] -> t ; conv22 3x3 -> x22         RITnet_v2.py:61-62
      build_conv(b.conv21, mem, P + ".conv21", {{b.buf, b.off_x, b.in_c, 0}, {b.buf, b.off_x1, b.inter, 0}},
                 {{W("conv21").data.data(), Bv("conv21").data.data(), 1, 0}}, b.inter, b.in_c + b.inter, 1, 1, b.H, b.W, E);
      set_store_epilogue(b.conv21, b.t, 0, ACT_NONE);
      finalize_conv(b.conv21);
      build_conv(b.conv22, mem, P + ".conv22", {{b.t, 0, b.inter, 0}}, {{W("conv22").data.data(), Bv("conv22").data.data(), 1, 1}},
                 b.inter, b.inter, 3, 3, b.H, b.W, E);
      set_store_epilogue(b.conv22, b.buf, b.off_x22, ACT_LRELU);
      finalize_conv(b.conv22);
      // conv31 on [x, x1, x22] -> t ; conv32 -> out             RITnet_v2.py:63-64
      build_conv(b.conv31, mem, P + ".conv31",
                 {{b.buf, b.off_x, b.in_c, 0}, {b.buf, b.off_x1, b.inter, 0}, {b.buf, b.off_x22, b.inter, 0}},
                 {{W("conv31").data.data(), Bv("conv31").data.data(), 1, 0}}, b.inter, b.in_c + 2 * b.inter, 1, 1, b.H, b.W, E);
      set_store_epilogue(b.conv31, b.t, 0, ACT_NONE);
      finalize_conv(b.conv31);
      build_conv(b.conv32, mem, P + ".conv32", {{b.t, 0, b.inter, 0}}, {{W("conv32").data.data(), Bv("conv32").data.data(), 1, 1}},
                 b.inter, b.inter, 3, 3, b.H, b.W, E);
      set_store_epilogue(b.conv32, b.buf, b.off_out, ACT_LRELU);
      set_stats(b.conv32, b.stats, b.stats_C, b.off_out);
      finalize_conv(b.conv32);
      // TD: IN -> lrelu -> (pool) -> 1x1 conv on skip = [out, x]   RITnet_v2.py:40-44,65-66
      const Act* dst = i < 4 ? es.blk[i + 1].buf : es.bt;
      const int dst_off = i < 4 ? es.blk[i + 1].off_x : 0;
      build_conv(b.td, mem, P + ".TD.conv", {{b.tdin, 0, b.inter, 0}, {b.tdin, b.inter, b.in_c, 0}},
                 {{W("TD.conv").data.data(), Bv("TD.conv").data.data(), 1, 0}}, b.op_c, b.inter + b.in_c, 1, 1,
                 b.tdin->H, b.tdin->W, E);
      set_store_epilogue(b.td, dst, dst_off, ACT_NONE);
      if (i < 4) set_stats(b.td, es.blk[i + 1].stats, es.blk[i + 1].stats_C, es.blk[i + 1].off_x);
      finalize_conv(b.td);
    }
    // decoder
    static const char* uname[4] = {"dec.up_block4", "dec.up_block3", "dec.up_block2", "dec.up_block1"};
    // Each up block (RITnet_v2.py:79-88) applies two 1x1 convolutions to cat[upsample2x(x), skip(, x1)].
    // Both the bilinear interpolation and the 1x1 convolution are linear and the interpolation weights
    // sum to one, so  W . cat[up(x), rest] + b  =  up(W_a x) + W_b rest + b :  `pre` evaluates
    // [W11_a ; W21_a] x once at HALF resolution and conv11 / conv21 add its bilinear upsample in their
    // epilogues.  The upsampled x (up to 320 channels at the skip's resolution) is never materialised.
    for (int i = 0; i < 4; ++i) {
      UpBlock& u = es.up[i];
      Block& sk = es.blk[3 - i];
      const std::string P = uname[i];
      debug_acts[P + ".out"] = {u.out, {0, u.out_c}};
      auto W = [&](const std::string& n) -> const HostTensor& { return sd_get(sd, P + "." + n + ".weight"); };
      auto Bv = [&](const std::string& n) -> const HostTensor& { return sd_get(sd, P + "." + n + ".bias"); };
      const HostTensor& w11 = W("conv11");
      const HostTensor& w21 = W("conv21");
      const int k11 = u.in_c + u.skip_c, k21 = k11 + u.out_c;
      EGN_CHECK(w11.shape[0] == u.out_c && w11.shape[1] == k11 && w21.shape[0] == u.out_c && w21.shape[1] == k21,
                P + ": unexpected 1x1 weight shapes");
      // half-resolution part: rows [0, out_c) <- W11[:, :in_c], rows [out_pad, out_pad + out_c) <- W21[:, :in_c]
      std::vector<float> w_pre((size_t)2 * u.out_pad * u.in_c, 0.f), w11s((size_t)u.out_c * u.skip_c), w21s((size_t)u.out_c * (u.skip_c + u.out_c));
      for (int co = 0; co < u.out_c; ++co) {
        for (int c = 0; c < u.in_c; ++c) {
          w_pre[(size_t)co * u.in_c + c] = w11.data[(size_t)co * k11 + c];
          w_pre[(size_t)(u.out_pad + co) * u.in_c + c] = w21.data[(size_t)co * k21 + c];
        }
        for (int c = 0; c < u.skip_c; ++c) w11s[(size_t)co * u.skip_c + c] = w11.data[(size_t)co * k11 + u.in_c + c];
        for (int c = 0; c < u.skip_c + u.out_c; ++c) w21s[(size_t)co * (u.skip_c + u.out_c) + c] = w21.data[(size_t)co * k21 + u.in_c + c];
      }
      std::vector<Piece> xin;
      if (i == 0) {
        xin.push_back({es.bt, 0, 153, 0});
        if (cfg.add_edge) xin.push_back({es.bt, 0, 153, mb});     // x = cat(x, x_edge) (RITnet_v2.py:286); offset patched per launch
      } else {
        xin.push_back({es.up[i - 1].out, 0, u.in_c, 0});
      }
      build_conv(u.pre, mem, P + ".pre", xin, {{w_pre.data(), nullptr, 1, 0}}, 2 * u.out_pad, u.in_c, 1, 1, u.H / 2, u.W / 2, mb);
      set_store_epilogue(u.pre, u.y, 0, ACT_NONE);
      u.pre.flops = 0;                                           // counted with conv11 / conv21 at the reference's sizes
      finalize_conv(u.pre);
      std::vector<Piece> x;
      x.push_back({sk.buf, sk.off_out, sk.inter, 0});
      x.push_back({sk.buf, sk.off_x, sk.in_c, 0});
      build_conv(u.c11, mem, P + ".conv11", x, {{w11s.data(), Bv("conv11").data.data(), 1, 0}}, u.out_c, u.skip_c, 1, 1, u.H, u.W, mb);
      set_store_epilogue(u.c11, u.t, 0, ACT_NONE);
      u.c11.e.up_hi = u.y->hi; u.c11.e.up_lo = u.y->lo; u.c11.e.up_C = u.y->C; u.c11.e.up_coff = 0;
      u.c11.flops = 2.0 * u.out_c * k11 * u.H * u.W;
      finalize_conv(u.c11);
      build_conv(u.c12, mem, P + ".conv12", {{u.t, 0, u.out_c, 0}}, {{W("conv12").data.data(), Bv("conv12").data.data(), 1, 1}},
                 u.out_c, u.out_c, 3, 3, u.H, u.W, mb);
      set_store_epilogue(u.c12, u.buf, 0, ACT_LRELU);
      finalize_conv(u.c12);
      std::vector<Piece> x21 = x;
      x21.push_back({u.buf, 0, u.out_c, 0});
      build_conv(u.c21, mem, P + ".conv21", x21, {{w21s.data(), Bv("conv21").data.data(), 1, 0}}, u.out_c, u.skip_c + u.out_c, 1, 1,
                 u.H, u.W, mb);
      set_store_epilogue(u.c21, u.t, 0, ACT_NONE);
      u.c21.e.up_hi = u.y->hi; u.c21.e.up_lo = u.y->lo; u.c21.e.up_C = u.y->C; u.c21.e.up_coff = u.out_pad;
      u.c21.flops = 2.0 * u.out_c * k21 * u.H * u.W;
      finalize_conv(u.c21);
      build_conv(u.c22, mem, P + ".conv22", {{u.t, 0, u.out_c, 0}}, {{W("conv22").data.data(), Bv("conv22").data.data(), 1, 1}},
                 u.out_c, u.out_c, 3, 3, u.H, u.W, mb);
      set_store_epilogue(u.c22, u.out, 0, ACT_LRELU);
      finalize_conv(u.c22);
    }
    // final convBlock: conv1 -> fin ; conv2 + lrelu + BN -> fp32 NCHW logits (both on the tensor cores)
    debug_acts["dec.final.conv1"] = {es.fin, {0, 32}};
    {
      const HostTensor& w1 = sd_get(sd, "dec.final.conv1.weight");
      const HostTensor& b1 = sd_get(sd, "dec.final.conv1.bias");
      build_conv(es.final1, mem, "dec.final.conv1", {{es.up[3].out, 0, 32, 0}}, {{w1.data.data(), b1.data.data(), 1, 1}}, 32, 32,
                 3, 3, 240, 320, mb);
      set_store_epilogue(es.final1, es.fin, 0, ACT_LRELU);
      finalize_conv(es.final1);
      const HostTensor& w2 = sd_get(sd, "dec.final.conv2.weight");     // [3][32][3][3]
      const HostTensor& b2 = sd_get(sd, "dec.final.conv2.bias");
      {
        // conv2 (32 -> 3) + lrelu + BN on the tensor cores too: N = 16 tile (3 live columns), fp32 NCHW logits
        build_conv(es.final2, mem, "dec.final.conv2", {{es.fin, 0, 32, 0}}, {{w2.data.data(), b2.data.data(), 1, 1}}, 3, 32,
                   3, 3, 240, 320, mb);
        std::vector<float> sh16;
        std::vector<float> sc16 = bn_scale(sd, "dec.final.bn", 3, es.final2.g.cout_pad, sh16);
        ConvEpi& e = es.final2.e;
        e.mode = CONV_LOGITS; e.act = ACT_LRELU; e.cout_store = es.final2.g.cout_pad;
        e.post_scale = mem.upload(sc16); e.post_shift = mem.upload(sh16);
        e.logits = nullptr; e.logits_c = 3;
        finalize_conv(es.final2);
      }
    }
    // regression head (utils.py:983-1037), fp32
    const int Cf = 153 * (cfg.add_edge ? 2 : 1);
    es.Cf = Cf;
    // c1 (2x3, valid) runs on the tensor cores as a "same"-geometry convolution over the 15x20 grid
    // whose outputs beyond the valid 14x18 region are never read; the rest of the head is one
    // fused SIMT kernel per frame (aux.cuh head_tail_kernel)
    EGN_CHECK(sd_get(sd, "elReg.c1.weight").shape[1] == Cf, "elReg.c1 input channels do not match the setting");
    {
      std::vector<Piece> hp;
      if (cfg.add_seg) {
        hp.push_back({es.hin, 0, 153, 0});
        if (cfg.add_edge) hp.push_back({es.hin, 160, 153, 0});
      } else {
        hp.push_back({es.bt, 0, 153, 0});
        if (cfg.add_edge) hp.push_back({es.bt, 0, 153, mb});      // edge frames: offset patched per launch
      }
      const HostTensor& w = sd_get(sd, "elReg.c1.weight");
      const HostTensor& b = sd_get(sd, "elReg.c1.bias");
      build_conv(es.head_c1, mem, "elReg.c1", hp, {{w.data.data(), b.data.data(), 1, 0}}, 128, Cf, 2, 3, 15, 20, mb);
      set_store_epilogue(es.head_c1, es.c1o, 0, ACT_LRELU);
      es.head_c1.flops = 2.0 * 128 * Cf * 6 * 14 * 18;
      finalize_conv(es.head_c1);
    }
    {
      HeadTailParams& h = es.head;
      h.c1 = make_view(*es.c1o, 0);
      h.w_c2 = upload_hwio(mem, sd_get(sd, "elReg.c2.weight")); h.b_c2 = mem.upload(sd_get(sd, "elReg.c2.bias").data);
      h.w_c3 = upload_hwio(mem, sd_get(sd, "elReg.c3.weight"));
      // l1 consumes the NCHW flatten c*15 + y*5 + x (utils.py:1020); c3's output here is NHWC, and the
      // matrix is stored input-major so consecutive threads read consecutive outputs
      const HostTensor& w = sd_get(sd, "elReg.l1.weight");
      std::vector<float> wt(480 * 256);
      for (int o = 0; o < 256; ++o)
        for (int c = 0; c < 32; ++c)
          for (int px = 0; px < 15; ++px) wt[(size_t)(px * 32 + c) * 256 + o] = w.data[(size_t)o * 480 + c * 15 + px];
      h.w_l1t = mem.upload(wt);
      h.b_l1 = mem.upload(sd_get(sd, "elReg.l1.bias").data);
      h.w_l2 = mem.upload(sd_get(sd, "elReg.l2.weight").data);
      h.b_l2 = mem.upload(sd_get(sd, "elReg.l2.bias").data);
    }
    es.adain = nullptr;
    if (cfg.add_seg) {
      // StyleEncoder (RITnet_v2.py:91-106) on the tensor cores: layer 0 as a 7x1 convolution over the
      // horizontally folded softmax (aux.cuh style_fold_kernel), layers 1-4 as 2x2 convolutions over the
      // reflect-padded space-to-depth of their input (s2d_reflect_kernel); MLP in fp32
      es.sm = (float*)mem.alloc((size_t)mb * 76800 * 3 * 4);
      const int sc[5] = {64, 128, 256, 256, 256};
      int vh = 240, vw = 320;                      // valid region of the current layer's input
      for (int i = 0; i < 5; ++i) {
        const std::string Pn = "seg_encoder.model." + std::to_string(i) + ".conv";
        const HostTensor& w = sd_get(sd, Pn + ".weight");
        const HostTensor& bsv = sd_get(sd, Pn + ".bias");
        int gh, gw, cin_t, kh, kw;
        const int cin_ref = i == 0 ? 3 : sc[i - 1], cout = sc[i], kk = i == 0 ? 7 : 4;
        EGN_CHECK(w.shape[0] == cout && w.shape[1] == cin_ref && w.shape[2] == kk && w.shape[3] == kk, Pn + ": unexpected weight shape");
        if (i == 0) {
          gh = vh + 6; gw = vw; cin_t = 32; kh = 7; kw = 1;
          es.st_w[i].assign((size_t)cout * cin_t * 7, 0.f);
          for (int co = 0; co < cout; ++co)
            for (int c = 0; c < 3; ++c)
              for (int dy = 0; dy < 7; ++dy)
                for (int dx = 0; dx < 7; ++dx)
                  es.st_w[i][((size_t)co * cin_t + dx * 3 + c) * 7 + dy] = w.data[(((size_t)co * 3 + c) * 7 + dy) * 7 + dx];
        } else {
          gh = (vh + 2) / 2; gw = (vw + 2) / 2; cin_t = 4 * cin_ref; kh = 2; kw = 2;
          es.st_w[i].assign((size_t)cout * cin_t * 4, 0.f);
          for (int co = 0; co < cout; ++co)
            for (int c = 0; c < cin_ref; ++c)
              for (int r = 0; r < 4; ++r)
                for (int q = 0; q < 4; ++q) {
                  const int par = (r & 1) * 2 + (q & 1), dy = r >> 1, dx = q >> 1;
                  es.st_w[i][(((size_t)co * cin_t + par * cin_ref + c) * 2 + dy) * 2 + dx] = w.data[(((size_t)co * cin_ref + c) * 4 + r) * 4 + q];
                }
        }
        EGN_CHECK(gh == st_gh[i] && gw == st_gw[i] && cin_t == st_cin[i], Pn + ": staged shapes disagree with the declared ones");
        build_conv(es.st_conv[i], mem, Pn, {{es.st_in[i], 0, cin_t, 0}}, {{es.st_w[i].data(), bsv.data.data(), 1, 0}}, cout, cin_t,
                   kh, kw, gh, gw, mb);
        set_store_epilogue(es.st_conv[i], es.st_out[i], 0, ACT_RELU);
        if (i > 0) { vh /= 2; vw /= 2; }
        es.st_conv[i].flops = 2.0 * cout * cin_ref * kk * kk * vh * vw;
        finalize_conv(es.st_conv[i]);
        es.st_w[i].clear(); es.st_w[i].shrink_to_fit();
      }
      es.gap = (float*)mem.alloc((size_t)mb * 256 * 4);
      es.sty = (float*)mem.alloc((size_t)mb * cfg.style_dim * 4);
      es.m1 = (float*)mem.alloc((size_t)mb * 256 * 4);
      es.m2 = (float*)mem.alloc((size_t)mb * 256 * 4);
      es.adain = (float*)mem.alloc((size_t)mb * 2 * Cf * 4);
      es.w_se6 = mem.upload(sd_get(sd, "seg_encoder.model.6.weight").data);   // [sd][256][1][1]
      es.b_se6 = mem.upload(sd_get(sd, "seg_encoder.model.6.bias").data);
      for (int i = 0; i < 3; ++i) {
        es.w_m[i] = mem.upload(sd_get(sd, "mlp.model." + std::to_string(i) + ".fc.weight").data);
        es.b_m[i] = mem.upload(sd_get(sd, "mlp.model." + std::to_string(i) + ".fc.bias").data);
      }
      EGN_CHECK(sd_get(sd, "mlp.model.2.fc.weight").shape[0] == 2 * Cf, "mlp output does not match feature channels");
    }
    es.logits_tmp = nullptr;
    built_esf = true;
  }

  void launch_first(const FirstConvParams& fp, cudaStream_t st) {
    const dim3 grid(ceil_div(fp.W, FC_TW), ceil_div(ceil_div(fp.H, FC_TH), FC_YT), fp.B);
    if (fp.cout == 64 && fp.cin == 1) first_conv_kernel<1, 64><<<grid, 256, 0, st>>>(fp);
    else if (fp.cout == 64 && fp.cin == 3) first_conv_kernel<3, 64><<<grid, 256, 0, st>>>(fp);
    else if (fp.cout == 32 && fp.cin == 1) first_conv_kernel<1, 32><<<grid, 256, 0, st>>>(fp);
    else if (fp.cout == 32 && fp.cin == 2) first_conv_kernel<2, 32><<<grid, 256, 0, st>>>(fp);
    else EGN_CHECK(false, "first_conv: unsupported cin/cout");
    CUDA_OK(cudaGetLastError());
    ++launches;
  }

  // y = act((x - mean) * rstd) [-> 2x2 average pool]; the statistics were accumulated by the producers' epilogues
  void inorm(const Act* src, int coff, int Cv, const double* stats, int stats_C, const Act* dst, int dcoff, int act,
             bool pool, int batch, cudaStream_t st) {
    NormApplyParams ap;
    ap.src = make_view(*src, coff); ap.dst = make_view(*dst, dcoff); ap.sums = stats; ap.sums_C = stats_C; ap.sums_coff = coff;
    ap.B = batch; ap.H = src->H; ap.W = src->W; ap.Cv = Cv; ap.act = act; ap.pool = pool ? 1 : 0;
    const long long per_frame = (long long)(pool ? src->H / 2 : src->H) * (pool ? src->W / 2 : src->W) * (Cv / 8);
    static const int max_blocks = getenv("EGN_IN_BLOCKS") ? atoi(getenv("EGN_IN_BLOCKS")) : 64;   // tuning knob: pixel slabs per frame
    dim3 grid((unsigned)std::max<long long>(1, std::min<long long>(max_blocks, per_frame / 2048)), (unsigned)batch);
    instnorm_apply_kernel<<<grid, 256, (size_t)Cv * sizeof(float2), st>>>(ap);
    CUDA_OK(cudaGetLastError());
    launches += 1;
  }

  // x, edge: device fp32 [B][H][W]; logits fp32 [B][3][H][W]; el_out [B][10]; latent [B][153]
  void esf_forward(const float* x, const float* edge, float* logits, float* el_out, float* latent, int B, cudaStream_t st) {
    if (!built_esf) build_esf();
    const size_t hw = (size_t)EGN_H * EGN_W;
    for (int b0 = 0; b0 < B; b0 += mb) {
      const int nb = std::min(mb, B - b0);
      const int E = nb * (cfg.add_edge ? 2 : 1);
      const float* xi = x + b0 * hw;
      const float* xe = edge ? edge + b0 * hw : nullptr;
      // ---- head conv1
      FirstConvParams fp = es.first;
      fp.B = nb; fp.dst = make_view(*es.h1, 0, 0);
      fp.in[0] = cfg.only_edge ? xe : xi;                     // RITnet_v2.py:276-278
      fp.in[1] = cfg.input_concat ? xe : nullptr;             // RITnet_v2.py:279-280
      fp.cin = cfg.input_concat ? 2 : 1;
      fp.fstride[0] = fp.fstride[1] = fp.fstride[2] = (long long)hw;
      EGN_CHECK(fp.in[0] != nullptr && (!cfg.input_concat || fp.in[1]), "edge input required by this setting");
      aux("esf.first_conv", st, [&] { launch_first(fp, st); });
      if (cfg.add_edge) {                                     // shared encoder on the edge map (F5)
        EGN_CHECK(xe != nullptr, "edge input required by add_edge");
        fp.in[0] = xe; fp.dst = make_view(*es.h1, 0, nb);
        aux("esf.first_conv", st, [&] { launch_first(fp, st); });
      }
      CUDA_OK(cudaMemsetAsync(es.stats_all, 0, es.stats_bytes, st));
      run_conv(es.head2, E, st);
      // ---- encoder blocks
      for (int i = 0; i < 5; ++i) {
        Block& b = es.blk[i];
        if (b.fold) {
          aux("esf.in_fold", st, [&] {
            const InFoldParams& f = b.conv1.fold;
            const size_t sm = (size_t)(2 * f.kpad + f.ntaps * f.cout_pad) * sizeof(float);
            in_fold_kernel<<<E, 256, sm, st>>>(f);
            CUDA_OK(cudaGetLastError()); ++launches;
          });
        } else {
          aux("esf.instnorm_x", st, [&] { inorm(b.buf, b.off_x, b.in_pad, b.stats, b.stats_C, b.xn, 0, ACT_NONE, false, E, st); });
        }
        run_conv(b.conv1, E, st);
        run_conv(b.conv21, E, st);
        run_conv(b.conv22, E, st);
        run_conv(b.conv31, E, st);
        run_conv(b.conv32, E, st);
        aux("esf.instnorm_td", st, [&] { inorm(b.buf, 0, b.inter + b.in_pad, b.stats, b.stats_C, b.tdin, 0, ACT_LRELU, i < 4, E, st); });
        run_conv(b.td, E, st);
      }
      // with add_edge the edge frames sit at [nb, 2nb) of every encoder buffer
      const int eoff = nb;
      spatial_mean_kernel<<<nb, 640, 0, st>>>(make_view(*es.bt, 0, 0), latent + (size_t)b0 * 153, nb, 300, 153);
      CUDA_OK(cudaGetLastError()); ++launches;
      // ---- regression head (utils.py:1013-1037): reads the bottleneck only unless the AdaIN branch feeds it
      auto run_head = [&](cudaStream_t hs) {
        if (cfg.add_seg) {
          AdainApplyParams ap;
          ap.src[0] = make_view(*es.bt, 0, 0); ap.src[1] = make_view(*es.bt, 0, eoff);
          ap.nsrc = cfg.add_edge ? 2 : 1; ap.Cs = 153; ap.Cslot = 160; ap.adain = es.adain;
          ap.dst = make_view(*es.hin, 0); ap.B = nb; ap.HW = 300;
          adain_apply_kernel<<<nb, 320, 0, hs>>>(ap); CUDA_OK(cudaGetLastError()); ++launches;
        }
        run_conv_impl(es.head_c1, nb, hs, cfg.add_seg ? -1 : eoff);
        HeadTailParams h = es.head;
        h.B = nb; h.el_out = el_out + (size_t)b0 * 10;
        head_tail_kernel<<<nb, HEAD_TAIL_THREADS, HEAD_TAIL_SMEM, hs>>>(h); CUDA_OK(cudaGetLastError()); ++launches;
      };
      // streaming micro-batches: the head runs on the side stream next to the decoder
      const cudaStream_t hs = cfg.add_seg ? st : side_of(st);
      side_used = 0;
      if (hs != st) { stream_edge(st, hs); run_head(hs); }
      // ---- decoder
      for (int i = 0; i < 4; ++i) {
        UpBlock& u = es.up[i];
        run_conv(u.pre, nb, st, (i == 0 && cfg.add_edge) ? eoff : -1);
        run_conv(u.c11, nb, st);
        run_conv(u.c12, nb, st);
        run_conv(u.c21, nb, st);
        run_conv(u.c22, nb, st);
      }
      run_conv(es.final1, nb, st);
      float* lg = logits + (size_t)b0 * 3 * hw;
      es.final2.tc.e.logits = lg; es.final2.simt.e.logits = lg;
      run_conv(es.final2, nb, st);
      // ---- AdaIN parameters from the softmaxed segmentation (RITnet_v2.py:289-308)
      if (cfg.add_seg) {
        {
          const long long total = (long long)nb * hw;
          softmax3_nhwc_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(lg, es.sm, nb, (int)hw);
          CUDA_OK(cudaGetLastError()); ++launches;
        }
        {
          StyleFoldParams fp2;
          fp2.sm = es.sm; fp2.dst = make_view(*es.st_in[0], 0); fp2.B = nb; fp2.H = 240; fp2.W = 320;
          launch_1d(style_fold_kernel, fp2, (long long)nb * 246 * 320 * 4, st); ++launches;
        }
        run_conv_impl(es.st_conv[0], nb, st);
        int vh = 240, vw = 320;
        for (int i = 1; i < 5; ++i) {
          S2dParams sp;
          sp.src = make_view(*es.st_out[i - 1], 0); sp.dst = make_view(*es.st_in[i], 0);
          sp.B = nb; sp.Hg = es.st_out[i - 1]->H; sp.Wg = es.st_out[i - 1]->W; sp.Hs = vh; sp.Ws = vw; sp.C = es.st_out[i - 1]->C;
          launch_1d(s2d_reflect_kernel, sp, (long long)nb * es.st_in[i]->H * es.st_in[i]->W * 4 * (sp.C / 8), st); ++launches;
          run_conv_impl(es.st_conv[i], nb, st);
          vh /= 2; vw /= 2;
        }
        gap_act_kernel<<<nb, 256, 0, st>>>(make_view(*es.st_out[4], 0), es.gap, es.st_out[4]->H, es.st_out[4]->W, 15, 20, 256);
        CUDA_OK(cudaGetLastError());
        linear(es.gap, es.w_se6, es.b_se6, es.sty, nb, 256, cfg.style_dim, 0, st);
        linear(es.sty, es.w_m[0], es.b_m[0], es.m1, nb, cfg.style_dim, 256, 1, st);
        linear(es.m1, es.w_m[1], es.b_m[1], es.m2, nb, 256, 256, 1, st);
        linear(es.m2, es.w_m[2], es.b_m[2], es.adain, nb, 256, 2 * es.Cf, 0, st);
        launches += 2;
      }
      // ---- regression head
      if (profiling) { if (prof_used >= 8192) profile_resolve(); profile_begin(st); }
      if (hs == st) run_head(st);
      else stream_edge(hs, st);
      if (profiling) profile_end(st, 0, nullptr, "esf.reg_head", 0);
    }
  }

  void linear(const float* in, const float* w, const float* b, float* out, int B, int I, int O, int act, cudaStream_t st) {
    const int total = B * O;
    linear_kernel<<<(total + 127) / 128, 128, 0, st>>>(in, w, b, out, B, I, O, act);
    CUDA_OK(cudaGetLastError()); ++launches;
  }

  std::string profile_table() {
    profile_resolve();
    char line[512];
    std::string out = "kind,name,H,W,kpad,cout,ntaps,n_tile,S_na_nw,launches,frames,ms_total,us_per_frame,tflops\n";
    for (auto& kv : conv_index) {
      ConvLayer& L = *kv.second;
      if (L.prof_n == 0) continue;
      snprintf(line, sizeof(line), "conv,%s,%d,%d,%d,%d,%d,%d,%d,%lld,%.0f,%.3f,%.2f,%.1f\n", L.name.c_str(), L.g.H, L.g.W,
               L.g.kpad, L.cout, L.g.ntaps, L.tc.n_tile, L.tc.S * 100 + L.tc.na * 10 + L.tc.nw, L.prof_n, L.prof_frames, L.prof_ms,
               1000.0 * L.prof_ms / L.prof_frames, L.flops * L.prof_frames / (L.prof_ms * 1e9));
      out += line;
    }
    for (auto& kv : prof_aux) {
      snprintf(line, sizeof(line), "aux,%s,,,,,,,,%lld,,%.3f,,\n", kv.first.c_str(), kv.second.second, kv.second.first);
      out += line;
    }
    return out;
  }

  double flops_per_frame_esf() const {
    const int reps = cfg.add_edge ? 2 : 1;
    double f = reps * 2.0 * 32 * (cfg.input_concat ? 2 : 1) * 9 * 240 * 320 + reps * es.head2.flops;
    for (int i = 0; i < 5; ++i) {
      const Block& b = es.blk[i];
      // the reference applies TD.conv before the pool (RITnet_v2.py:42-43): count it at full size
      const double td = 2.0 * b.op_c * (b.inter + b.in_c) * b.H * b.W;
      f += reps * (b.conv1.flops + b.conv21.flops + b.conv22.flops + b.conv31.flops + b.conv32.flops + td);
    }
    for (int i = 0; i < 4; ++i) f += es.up[i].c11.flops + es.up[i].c12.flops + es.up[i].c21.flops + es.up[i].c22.flops;
    f += es.final1.flops + 2.0 * 3 * 32 * 9 * 240 * 320;
    f += 2.0 * (128.0 * es.Cf * 6 * 14 * 18 + 128.0 * 128 * 9 * 5 * 7 + 32.0 * 128 * 9 * 3 * 5 + 480 * 256 + 2560);
    if (cfg.add_seg) {
      f += 2.0 * (64.0 * 3 * 49 * 76800 + 128.0 * 64 * 16 * 19200 + 256.0 * 128 * 16 * 4800 + 256.0 * 256 * 16 * 1200 +
                  256.0 * 256 * 16 * 300 + 256 * cfg.style_dim + cfg.style_dim * 256 + 256 * 256 + 256.0 * 2 * es.Cf);
    }
    return f;
  }

  double flops_per_frame_bdcn() const {
    double f = 2.0 * 64 * 3 * 9 * 240 * 320;
    for (int i = 1; i < 13; ++i) f += bd.vgg[i].flops;
    for (int i = 0; i < 13; ++i) f += bd.ms_in[i].flops + bd.ms_tail[i].flops;
    return f;
  }
};
