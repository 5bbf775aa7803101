// extern "C" surface of libegn.so (declared in include/egn.h).
#include "../../include/egn.h"

#include "engine.cuh"

struct egn_ctx {
  Engine eng;
  float* partial = nullptr;   // seg-post slice partials
  int partial_cap = 0;
  int* counts = nullptr;
  int counts_cap = 0;
  double* loss_partial = nullptr;
  int loss_cap = 0;
};

static thread_local std::string g_err;

// Every entry point runs with the context's device current and puts the caller's device back on exit
// (torch's notion of the current device must not change under the caller).
struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) CUDA_OK(cudaSetDevice(dev)); else prev = -1;
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

#define API_BEGIN try {
#define API_END                                   \
  }                                               \
  catch (const std::exception& e) {               \
    g_err = e.what();                             \
    return 1;                                     \
  }                                               \
  catch (...) {                                   \
    g_err = "unknown error";                      \
    return 1;                                     \
  }                                               \
  return 0;

extern "C" {

const char* egn_last_error(void) { return g_err.c_str(); }
int egn_version(void) { return 200; }

int egn_create(int device, const egn_config* cfg, egn_ctx** out) {
  API_BEGIN
  EGN_CHECK(out != nullptr && cfg != nullptr, "null argument");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  EGN_CHECK(e == cudaSuccess && ndev > 0, std::string("no CUDA device: ") + cudaGetErrorString(e));
  EGN_CHECK(device >= 0 && device < ndev, "bad device index");
  DeviceGuard guard(device);
  cudaDeviceProp prop;
  CUDA_OK(cudaGetDeviceProperties(&prop, device));
  EGN_CHECK(prop.major == 10, "libegn.so is built for sm_100a (B200) only; found sm_" +
                                  std::to_string(prop.major) + std::to_string(prop.minor));
  std::unique_ptr<egn_ctx> c(new egn_ctx());
  c->eng.device = device;
  c->eng.num_sms = prop.multiProcessorCount;
  if (const char* ns = getenv("EGN_NUM_SMS")) c->eng.num_sms = std::max(1, std::min(atoi(ns), prop.multiProcessorCount));   // tuning knob: persistent-grid cap
  c->eng.cfg.add_edge = cfg->add_edge; c->eng.cfg.add_seg = cfg->add_seg;
  c->eng.cfg.seg_detach = cfg->seg_detach; c->eng.cfg.input_concat = cfg->input_concat;
  c->eng.cfg.only_edge = cfg->only_edge; c->eng.cfg.style_dim = cfg->style_dim;
  const char* impl = getenv("EGN_CONV");
  c->eng.use_tc = !(impl && std::string(impl) == "simt");
  const char* ns = getenv("EGN_NSPLIT");
  c->eng.nsplit = ns ? atoi(ns) : 3;
  EGN_CHECK(c->eng.nsplit == 1 || c->eng.nsplit == 3, "EGN_NSPLIT must be 1 or 3");
  c->eng.err_flag = (int*)c->eng.mem_misc.alloc(sizeof(int));
  Engine::prepare_device();
  if (const char* fg = getenv("EGN_L2_FETCH")) {      // tuning knob: L2 fetch granularity hint (32 / 64 / 128 bytes)
    CUDA_OK(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(fg)));
  }
  *out = c.release();
  API_END
}

int egn_destroy(egn_ctx* ctx) {
  API_BEGIN
  if (ctx) {
    DeviceGuard guard(ctx->eng.device);
    cudaDeviceSynchronize();
    delete ctx;
  }
  API_END
}

int egn_set_weights(egn_ctx* ctx, int net, const void* blob, size_t bytes) {
  API_BEGIN
  EGN_CHECK(ctx && blob, "null argument");
  DeviceGuard guard(ctx->eng.device);
  if (net == EGN_NET_BDCN) {
    EGN_CHECK(!ctx->eng.built_bdcn, "BDCN weights already bound; create a new context to reload");
    ctx->eng.sd_bdcn = parse_blob(blob, bytes);
    ctx->eng.has_bdcn = true;
  } else if (net == EGN_NET_ESF) {
    EGN_CHECK(!ctx->eng.built_esf, "ESF-Net weights already bound; create a new context to reload");
    ctx->eng.sd_esf = parse_blob(blob, bytes);
    ctx->eng.has_esf = true;
  } else {
    EGN_CHECK(false, "unknown net id");
  }
  API_END
}

int egn_share_workspace(int enable) {
  share_workspace_flag() = enable != 0;
  return 0;
}

int egn_plan(egn_ctx* ctx, int micro_batch) {
  API_BEGIN
  EGN_CHECK(ctx, "null context");
  EGN_CHECK(micro_batch >= 1 && micro_batch <= 1024, "micro_batch must be in [1,1024]");
  EGN_CHECK(!ctx->eng.built_bdcn && !ctx->eng.built_esf, "plan must precede the first forward");
  ctx->eng.mb = micro_batch;
  ctx->eng.decide_concurrency();
  API_END
}

int egn_bdcn_forward(egn_ctx* ctx, const float* x, int planes, float* edge_out, int batch, void* stream) {
  API_BEGIN
  EGN_CHECK(ctx && x && edge_out && batch > 0, "bad argument");
  EGN_CHECK(planes == 1 || planes == 3, "planes must be 1 (grey, replicated by the caller's cat) or 3");
  DeviceGuard guard(ctx->eng.device);
  ctx->eng.bdcn_forward(x, planes, edge_out, batch, (cudaStream_t)stream);
  API_END
}

int egn_bdcn_forward_all(egn_ctx* ctx, const float* x, int planes, float* edge_out, float* side_out, int batch,
                         void* stream) {
  API_BEGIN
  EGN_CHECK(ctx && x && edge_out && side_out && batch > 0, "bad argument");
  EGN_CHECK(planes == 1 || planes == 3, "planes must be 1 (grey, replicated by the caller's cat) or 3");
  DeviceGuard guard(ctx->eng.device);
  ctx->eng.bdcn_forward(x, planes, edge_out, batch, (cudaStream_t)stream, side_out);
  API_END
}

int egn_info(egn_ctx* ctx, egn_info_t* out) {
  API_BEGIN
  EGN_CHECK(ctx && out, "null argument");
  const Engine& e = ctx->eng;
  out->device = e.device; out->num_sms = e.num_sms; out->micro_batch = e.mb;
  out->products_per_mac = e.nsplit; out->tensor_core_path = e.use_tc ? 1 : 0;
  out->workspace_bytes = (long long)(e.mem_bdcn.total + e.mem_esf.total + e.mem_misc.total);
  out->activation_bytes_unshared = (long long)e.arena_naive_bytes;
  out->shared_pool_bytes = 0;
  if (e.uses_pool) out->shared_pool_bytes = (long long)shared_pools()[e.device].bytes;
  out->lowered_layers = 0;
  for (auto& kv : e.conv_index) out->lowered_layers += (kv.second->prod_mode != 0) + (kv.second->w_frames != 0);
  API_END
}

int egn_esf_forward(egn_ctx* ctx, const float* x, const float* edge, float* logits, float* el_out,
                    float* latent, int batch, void* stream) {
  API_BEGIN
  EGN_CHECK(ctx && x && logits && el_out && latent && batch > 0, "bad argument");
  DeviceGuard guard(ctx->eng.device);
  ctx->eng.esf_forward(x, edge, logits, el_out, latent, batch, (cudaStream_t)stream);
  API_END
}

int egn_seg_post(egn_ctx* ctx, const float* logits, const float* el_out, const float* cond,
                 uint8_t* argmax_u8, float* el_pred, int batch, void* stream) {
  API_BEGIN
  EGN_CHECK(ctx && logits && el_out && argmax_u8 && el_pred && batch > 0, "bad argument");
  DeviceGuard guard(ctx->eng.device);
  cudaStream_t st = (cudaStream_t)stream;
  if (ctx->partial_cap < batch) {
    ctx->partial = (float*)ctx->eng.mem_misc.alloc((size_t)batch * POST_SLICES * 8 * sizeof(float));
    ctx->partial_cap = batch;
  }
  dim3 grid(POST_SLICES, batch);
  seg_post_kernel<<<grid, POST_THREADS, 0, st>>>(logits, argmax_u8, ctx->partial, batch);
  CUDA_OK(cudaGetLastError());
  seg_post_finish_kernel<<<(batch + 127) / 128, 128, 0, st>>>(ctx->partial, el_out, cond, el_pred, batch);
  CUDA_OK(cudaGetLastError());
  ctx->eng.launches += 2;
  API_END
}

int egn_metrics_accumulate(egn_ctx* ctx, const uint8_t* argmax_u8, const void* labels, int label_is_i64,
                           const float* cond, const float* pupil_c, const float* iris_c,
                           const float* el_out, const float* el_pred, double* acc,
                           float* iou_by_sample, int batch, void* stream) {
  API_BEGIN
  EGN_CHECK(ctx && argmax_u8 && labels && cond && acc && batch > 0, "bad argument");
  EGN_CHECK((pupil_c == nullptr) == (iris_c == nullptr), "pupil_c and iris_c must be given together");
  EGN_CHECK(!pupil_c || (el_out && el_pred), "centres need el_out and el_pred");
  DeviceGuard guard(ctx->eng.device);
  cudaStream_t st = (cudaStream_t)stream;
  if (ctx->counts_cap < batch) {
    ctx->counts = (int*)ctx->eng.mem_misc.alloc((size_t)batch * 9 * sizeof(int));
    ctx->counts_cap = batch;
  }
  CUDA_OK(cudaMemsetAsync(ctx->counts, 0, (size_t)batch * 9 * sizeof(int), st));
  dim3 grid(10, batch);
  seg_counts_kernel<<<grid, 256, 0, st>>>(argmax_u8, (const uint8_t*)labels, label_is_i64 ? 8 : 1, ctx->counts,
                                          EGN_H * EGN_W);
  CUDA_OK(cudaGetLastError());
  metrics_finish_kernel<<<(batch + 127) / 128, 128, 0, st>>>(ctx->counts, cond, pupil_c, iris_c, el_out, el_pred, acc,
                                                             iou_by_sample, batch);
  CUDA_OK(cudaGetLastError());
  ctx->eng.launches += 2;
  API_END
}

int egn_forward_loss(egn_ctx* ctx, const float* logits, const void* target, int target_is_i64, const float* spat_w,
                     const float* dist_map, const float* cond, const float* pupil_c, const float* el_norm,
                     const float* el_out, const float* el_pred, float alpha, float* loss, int batch, void* stream) {
  API_BEGIN
  EGN_CHECK(ctx && logits && target && spat_w && dist_map && cond && pupil_c && el_norm && el_out && el_pred && loss &&
                batch > 0, "bad argument");
  DeviceGuard guard(ctx->eng.device);
  cudaStream_t st = (cudaStream_t)stream;
  if (ctx->loss_cap < batch) {
    ctx->loss_partial = (double*)ctx->eng.mem_misc.alloc((size_t)batch * LOSS_SLICES * LOSS_TERMS * sizeof(double));
    ctx->loss_cap = batch;
  }
  dim3 grid(LOSS_SLICES, batch);
  seg_loss_kernel<<<grid, 256, 0, st>>>(logits, (const uint8_t*)target, target_is_i64 ? 8 : 1, spat_w, dist_map,
                                        ctx->loss_partial);
  CUDA_OK(cudaGetLastError());
  seg_loss_finish_kernel<<<1, 256, 0, st>>>(ctx->loss_partial, cond, pupil_c, el_norm, el_out, el_pred, alpha, loss, batch);
  CUDA_OK(cudaGetLastError());
  ctx->eng.launches += 2;
  API_END
}

int egn_ellipse_refine(egn_ctx* ctx, const uint8_t* argmax_u8, const float* ell_norm, double* out,
                       int refine, int batch, void* stream) {
  API_BEGIN
  EGN_CHECK(ctx && argmax_u8 && ell_norm && out && batch > 0, "bad argument");
  DeviceGuard guard(ctx->eng.device);
  dim3 grid(REFINE_CLUSTER, 2, batch);             // one 8-CTA cluster per ellipse (static __cluster_dims__)
  // two directions per raster pass (post.cuh): fewer sequential passes; measured faster at every batch from 1 (1.82 -> 1.30 ms)
  // to 128 (8.61 -> 8.21 ms).  EGN_REFINE_SPECULATE=0 selects the reference's one-candidate-per-pass order (same results).
  const char* spec_env = getenv("EGN_REFINE_SPECULATE");
  const int speculate = spec_env ? (atoi(spec_env) != 0) : 1;
  ellipse_refine_kernel<<<grid, REFINE_THREADS, 0, (cudaStream_t)stream>>>(argmax_u8, ell_norm, out, refine, speculate);
  CUDA_OK(cudaGetLastError());
  ctx->eng.launches += 1;
  API_END
}

int egn_preprocess_u8(egn_ctx* ctx, const uint8_t* frames_u8, float* out, int batch, void* stream) {
  API_BEGIN
  EGN_CHECK(ctx && frames_u8 && out && batch > 0, "bad argument");
  DeviceGuard guard(ctx->eng.device);
  preprocess_u8_kernel<<<batch, PRE_THREADS, 0, (cudaStream_t)stream>>>(frames_u8, out);
  CUDA_OK(cudaGetLastError());
  ctx->eng.launches += 1;
  API_END
}

int egn_profile(egn_ctx* ctx, int enable) {
  API_BEGIN
  EGN_CHECK(ctx, "null context");
  ctx->eng.profiling = enable != 0;
  API_END
}

int egn_profile_read(egn_ctx* ctx, double* conv_ms, double* conv_flops, long long* conv_launches, int reset) {
  API_BEGIN
  EGN_CHECK(ctx && conv_ms && conv_flops && conv_launches, "bad argument");
  DeviceGuard guard(ctx->eng.device);
  ctx->eng.profile_resolve();
  *conv_ms = ctx->eng.prof_ms; *conv_flops = ctx->eng.prof_flops; *conv_launches = ctx->eng.prof_launches;
  if (reset) { ctx->eng.prof_ms = 0; ctx->eng.prof_flops = 0; ctx->eng.prof_launches = 0; }
  API_END
}

long long egn_profile_table(egn_ctx* ctx, char* out, long long capacity) {
  try {
    if (!ctx) return -1;
    DeviceGuard guard(ctx->eng.device);
    const std::string t = ctx->eng.profile_table();
    if (out && capacity > 0) {
      const size_t n = std::min((size_t)capacity - 1, t.size());
      memcpy(out, t.data(), n);
      out[n] = 0;
    }
    return (long long)t.size() + 1;
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}

long long egn_launch_count(egn_ctx* ctx) { return ctx ? ctx->eng.launches : -1; }

double egn_flops_per_frame(egn_ctx* ctx, int net) {
  if (!ctx) return -1;
  try {
    return net == EGN_NET_BDCN ? ctx->eng.flops_per_frame_bdcn() : ctx->eng.flops_per_frame_esf();
  } catch (...) { return -1; }
}

long long egn_debug_read(egn_ctx* ctx, const char* name, float* out, long long capacity, int frames,
                         int* dims) {
  try {
    if (!ctx || !name) return -1;
    Engine& e = ctx->eng;
    DeviceGuard guard(e.device);
    cudaDeviceSynchronize();
    auto it = e.debug_acts.find(name);
    if (it != e.debug_acts.end()) {
      const Act* a = it->second.first;
      const int coff = it->second.second.first, C = it->second.second.second;
      const int n = std::min(frames, a->N);
      if (dims) { dims[0] = C; dims[1] = a->H; dims[2] = a->W; }
      const long long total = (long long)n * C * a->H * a->W;
      if (!out) return total;
      if (capacity < total) return -1;
      const size_t elems = (size_t)n * a->H * a->W * a->C;
      std::vector<bf16> hi(elems), lo(elems);
      cudaMemcpy(hi.data(), a->hi, elems * 2, cudaMemcpyDeviceToHost);
      cudaMemcpy(lo.data(), a->lo, elems * 2, cudaMemcpyDeviceToHost);
      for (int f = 0; f < n; ++f)
        for (int y = 0; y < a->H; ++y)
          for (int x = 0; x < a->W; ++x)
            for (int c = 0; c < C; ++c) {
              const size_t i = (((size_t)f * a->H + y) * a->W + x) * a->C + coff + c;
              out[(((size_t)f * C + c) * a->H + y) * a->W + x] = __bfloat162float(hi[i]) + __bfloat162float(lo[i]);
            }
      return total;
    }
    auto jt = e.debug_f32.find(name);
    if (jt != e.debug_f32.end()) {
      const std::vector<int>& d = jt->second.second;     // H, W, C (NHWC fp32)
      if (dims) { dims[0] = d[2]; dims[1] = d[0]; dims[2] = d[1]; }
      const long long total = (long long)frames * d[0] * d[1] * d[2];
      if (!out) return total;
      if (capacity < total) return -1;
      std::vector<float> tmp(total);
      cudaMemcpy(tmp.data(), jt->second.first, total * 4, cudaMemcpyDeviceToHost);
      for (int f = 0; f < frames; ++f)
        for (int y = 0; y < d[0]; ++y)
          for (int x = 0; x < d[1]; ++x)
            for (int c = 0; c < d[2]; ++c)
              out[(((size_t)f * d[2] + c) * d[0] + y) * d[1] + x] = tmp[(((size_t)f * d[0] + y) * d[1] + x) * d[2] + c];
      return total;
    }
    g_err = std::string("unknown debug tensor: ") + name;
    return -1;
  } catch (const std::exception& ex) {
    g_err = ex.what();
    return -1;
  }
}

int egn_conv_selfcheck(egn_ctx* ctx, const char* layer, int frames, double* max_diff, double* max_ref) {
  API_BEGIN
  EGN_CHECK(ctx && layer && max_diff && max_ref, "bad argument");
  Engine& e = ctx->eng;
  DeviceGuard guard(e.device);
  EGN_CHECK(e.use_tc, "selfcheck needs the tensor-core path (unset EGN_CONV=simt)");
  auto it = e.conv_index.find(layer);
  if (it == e.conv_index.end()) {
    it = e.conv_alias.find(layer);
    EGN_CHECK(it != e.conv_alias.end(), std::string("unknown conv layer: ") + layer);
  }
  ConvLayer& L = *it->second;
  EGN_CHECK(!L.w_frames, std::string(layer) + ": InstanceNorm is folded into this layer (per-frame weights); the SIMT companion has no such mode - "
                                              "set EGN_IN_FOLD=0 to cross-check it");
  const int n = std::min(frames, L.g.batch);
  const size_t px = (size_t)n * L.g.H * L.g.W;
  std::vector<float> a, b;
  float* logits_scratch = nullptr;
  auto fetch = [&](std::vector<float>& v) {
    CUDA_OK(cudaDeviceSynchronize());
    if (L.e.mode == CONV_STORE) {
      const size_t elems = px * L.e.out_C;
      std::vector<bf16> hi(elems), lo(elems);
      CUDA_OK(cudaMemcpy(hi.data(), L.e.out_hi, elems * 2, cudaMemcpyDeviceToHost));
      CUDA_OK(cudaMemcpy(lo.data(), L.e.out_lo, elems * 2, cudaMemcpyDeviceToHost));
      v.resize(px * L.e.cout_store);
      for (size_t p = 0; p < px; ++p)
        for (int c = 0; c < L.e.cout_store; ++c) {
          const size_t i = p * L.e.out_C + L.e.out_coff + c;
          v[p * L.e.cout_store + c] = __bfloat162float(hi[i]) + __bfloat162float(lo[i]);
        }
    } else if (L.e.mode == CONV_LOGITS) {
      v.resize(px * L.e.logits_c);
      CUDA_OK(cudaMemcpy(v.data(), logits_scratch, v.size() * sizeof(float), cudaMemcpyDeviceToHost));
    } else {
      v.resize(px * 2);
      CUDA_OK(cudaMemcpy(v.data(), L.e.score, px * 2 * sizeof(float), cudaMemcpyDeviceToHost));
    }
  };
  if (L.e.mode == CONV_LOGITS) {     // the logits tensor belongs to the caller: run both kernels into a scratch copy
    CUDA_OK(cudaMalloc(&logits_scratch, px * L.e.logits_c * sizeof(float)));
    L.tc.e.logits = logits_scratch; L.simt.e.logits = logits_scratch;
  }
  TcParams tp = L.tc;
  tp.g.batch = n * (L.phase ? L.phase * L.phase : 1);
  tp.total_tiles = tp.tiles_x * tp.tiles_y * tp.g.batch * tp.n_blocks; tp.e.score_accum = 0;
  tc_launch(tp, e.num_sms, 0);
  fetch(a);
  SimtParams sp = L.simt;
  sp.g.batch = n; sp.e.score_accum = 0;
  simt_launch(sp, 0);
  fetch(b);
  double md = 0, mr = 0;
  for (size_t i = 0; i < a.size(); ++i) {
    md = std::max(md, (double)fabsf(a[i] - b[i]));
    mr = std::max(mr, (double)fabsf(b[i]));
    if (a[i] != a[i]) md = 1e30;
  }
  *max_diff = md; *max_ref = mr;
  if (logits_scratch) {
    L.tc.e.logits = nullptr; L.simt.e.logits = nullptr;
    cudaFree(logits_scratch);
  }
  API_END
}

}  // extern "C"
