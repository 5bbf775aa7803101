/* egn.h - C ABI of libegn.so, the sm_100a engine for the edge-guided near-eye hot path.
 *
 * The reference has no FFI/plugin interface for this path: its boundary is the Python object
 * protocol of two nn.Modules.  Each entry point below names the reference
 * interface it stands behind (paths relative to the reference repository).  The Python mirrors
 * (egn_b200.bdcn_new.BDCN, egn_b200.ritnet_v2.DenseNet2D) are the only intended callers; the
 * ctypes binding a maintainer would add is shown in INTEGRATION.md.
 *
 * Conventions: every function returns 0 on success and a non-zero code on failure, after which
 * egn_last_error() returns a thread-local message.  All tensor arguments are raw DEVICE pointers
 * owned by the caller (fp32, contiguous, the reference's NCHW layouts) unless a parameter says
 * "host".  `stream` is a cudaStream_t passed as void*.  A context belongs to one device and is not
 * thread-safe; several contexts (one per device) may live in one process, and every call leaves the
 * caller's current device as it found it.  Frames are 240x320 (utils.py:1007 hard-wires the 15x20 bottleneck).
 */
#ifndef EGN_H_
#define EGN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct egn_ctx egn_ctx;

/* yaml `setting` dict of models/RITnet_v2.py:204-238 (configs/<name>.yaml); unread keys omitted. */
typedef struct egn_config {
  int add_edge;      /* RITnet_v2.py:184,227,283 */
  int add_seg;       /* RITnet_v2.py:231,289    */
  int seg_detach;    /* RITnet_v2.py:291 (no effect at inference) */
  int input_concat;  /* RITnet_v2.py:222,279    */
  int only_edge;     /* RITnet_v2.py:276        */
  int style_dim;     /* RITnet_v2.py:233        */
} egn_config;

enum { EGN_NET_BDCN = 0, EGN_NET_ESF = 1 };

const char* egn_last_error(void);
int egn_version(void);

/* Replaces: BDCN() / DenseNet2D(setting) construction + .cuda() (test.py:280-297). */
int egn_create(int device, const egn_config* cfg, egn_ctx** out);
int egn_destroy(egn_ctx* ctx);

/* Replaces: load_state_dict (test.py:283,295).  `blob` is a HOST buffer holding the state_dict
 * tensors under their reference names (format: csrc/engine.cuh parse_blob). */
int egn_set_weights(egn_ctx* ctx, int net, const void* blob, size_t bytes);

/* Opt-in, process-wide, before the first forward of the contexts it should cover: every context of a device draws
 * its activation arena from ONE pool.  Only for callers that run their contexts in stream order on one stream (an
 * evaluator that owns both modules: bench.py, calc_acc) - the BDCN context's buffers are dead once its edge map is
 * out, so the ESF-Net may overwrite them.  Off by default (independent modules may run on different streams). */
int egn_share_workspace(int enable);

/* Sizes the workspace for micro-batches of `micro_batch` frames (any caller batch is processed in
 * such slices).  Must precede the first forward. */
int egn_plan(egn_ctx* ctx, int micro_batch);

/* Replaces: utils.calc_edge -> BDCN.forward(cat(img,img,img))[-1] (utils.py:645-651,
 * bdcn_new.py:116-191).  x: [B,planes,240,320] with planes == 3 (what BDCN.forward receives) or
 * planes == 1 (the grey frame calc_edge replicates; skips the cat).  edge_out: [B,1,240,320]. */
int egn_bdcn_forward(egn_ctx* ctx, const float* x, int planes, float* edge_out, int batch, void* stream);

/* Replaces: the full return value of BDCN.forward (bdcn_new.py:116-191): besides the fused map,
 * side_out receives the ten per-scale sigmoids in the reference's order
 * [p1_1, p2_1, p3_1, p4_1, p5_1, p1_2, p2_2, p3_2, p4_2, p5_2], laid out [10][B][1][240][320].
 * utils.calc_edge and evaluate.py only read [-1] (egn_bdcn_forward); callers that iterate the list
 * get these through the Python mirror's lazy list. */
int egn_bdcn_forward_all(egn_ctx* ctx, const float* x, int planes, float* edge_out, float* side_out, int batch,
                         void* stream);

/* Replaces: DenseNet2D.forward up to elOut (RITnet_v2.py:261-310).  x, edge: [B,1,240,320]
 * (edge may be NULL when the setting does not read it); logits: [B,3,240,320]; el_out: [B,10];
 * latent: [B,153]. */
int egn_esf_forward(egn_ctx* ctx, const float* x, const float* edge, float* logits, float* el_out,
                    float* latent, int batch, void* stream);

/* Replaces: utils.get_predictions (utils.py:65-81) and the two loss.get_seg2ptLoss centres that
 * get_allLoss folds into elPred (loss.py:16-46, RITnet_v2.py:387-412,334-335).
 * argmax_u8: [B,240,320]; el_pred: [B,10]; cond: [B,4] fp32 or NULL - when no sample of the batch
 * has a GT mask (sum(1-cond[:,1]) == 0) the iris centre is taken from el_out[:,5:7]. */
int egn_seg_post(egn_ctx* ctx, const float* logits, const float* el_out, const float* cond,
                 uint8_t* argmax_u8, float* el_pred, int batch, void* stream);

/* Replaces: utils.getSeg_metrics / getPoint_metric per batch (utils.py:120-162, test.py:159-214).
 * labels: device [B,240,320], u8 when label_is_i64 == 0 else int64; cond: [B,4] fp32;
 * pupil_c / iris_c: [B,2] pixel centres (may both be NULL); acc: 16 doubles, accumulated into
 * (layout in csrc/post.cuh); iou_by_sample: [B,3] fp32 or NULL. */
int egn_metrics_accumulate(egn_ctx* ctx, const uint8_t* argmax_u8, const void* labels, int label_is_i64,
                           const float* cond, const float* pupil_c, const float* iris_c,
                           const float* el_out, const float* el_pred, double* acc,
                           float* iou_by_sample, int batch, void* stream);

/* Replaces: the loss slot of DenseNet2D.forward, get_allLoss (models/RITnet_v2.py:312-323,372-440) with
 * loss.get_segLoss / SurfaceLoss / GDiceLoss / wCE / get_ptLoss / get_seg2ptLoss (loss.py:16-137):
 * total = l_seg2pt + 20 * l_seg + 10 * (l_pt + l_ellipse), forward value only (no gradients).
 * target: [B,240,320] u8 or int64; spat_w: [B,240,320]; dist_map: [B,3,240,320]; cond: [B,4];
 * pupil_c: [B,2] pixels; el_norm: [B,2,5]; el_out, el_pred: [B,10] (el_pred from egn_seg_post);
 * loss: [1] fp32.  Samples whose GT lacks more than one class make the reference raise
 * (loss.py:128-132); here their cross-entropy is simply taken over all pixels. */
int egn_forward_loss(egn_ctx* ctx, const float* logits, const void* target, int target_is_i64, const float* spat_w,
                     const float* dist_map, const float* cond, const float* pupil_c, const float* el_norm,
                     const float* el_out, const float* el_pred, float alpha, float* loss, int batch, void* stream);

/* Replaces: my_ellipse(norm).transform(H) + search_proper_parameter_iou_for_our_data
 * (evaluate.py:141-151, helperfunctions.py:25-63,102-129, utils.py:176-204,450-486).
 * ell_norm: [B,2,5] normalised ellipses (iris, pupil) = elPred; out: [B,2,5] fp64 pixel-space
 * (cx, cy, a, b, theta); refine == 0 returns the plain transform. */
int egn_ellipse_refine(egn_ctx* ctx, const uint8_t* argmax_u8, const float* ell_norm, double* out,
                       int refine, int batch, void* stream);

/* Replaces: the per-frame z-score of evaluate.preprocess_frame (evaluate.py:102-103) and of the
 * DataLoader (CurriculumLib.py:139-140): (img - img.mean()) / img.std() in float64, cast to fp32.
 * frames_u8: device [B,240,320] uint8 (already 240x320 grey); out: device [B,1,240,320] fp32. */
int egn_preprocess_u8(egn_ctx* ctx, const uint8_t* frames_u8, float* out, int batch, void* stream);

/* Per-launch timing of the convolution kernel (CUDA event pairs on the launching stream).
 * egn_profile_read returns the summed kernel milliseconds, the algorithmic FLOPs (2*MAC at the
 * reference's unpadded sizes) and the number of launches since the last reset. */
int egn_profile(egn_ctx* ctx, int enable);
int egn_profile_read(egn_ctx* ctx, double* conv_ms, double* conv_flops, long long* conv_launches, int reset);

/* CSV of per-layer kernel times gathered in profiling mode; returns the bytes needed. */
long long egn_profile_table(egn_ctx* ctx, char* out, long long capacity);

/* What the context actually runs: bench.py reports `dtype` and the precision ceiling from here, so a
 * tuning knob (EGN_NSPLIT, EGN_CONV) can never hide behind a hard-coded label. */
typedef struct egn_info_t {
  int device, num_sms, micro_batch;
  int products_per_mac;      /* bf16 tensor-core products per algorithmic MAC: 3 = hi*hi + lo*hi + hi*lo (default), 1 = plain bf16 */
  int tensor_core_path;      /* 1 = tcgen05 kernel, 0 = SIMT companion (EGN_CONV=simt, debugging only) */
  long long workspace_bytes; /* device memory held by the context so far */
  long long activation_bytes_unshared; /* what the activation planes would take without liveness sharing (engine.cuh commit_acts) */
  long long shared_pool_bytes; /* size of the per-device activation pool this context draws from (egn_share_workspace), not part of workspace_bytes; 0 = own arena */
  int lowered_layers;        /* convolution layers that run below the parity precision: fewer products than products_per_mac (EGN_PRODUCTS probe
                              * knob) or InstanceNorm folded into per-frame weights (EGN_IN_FOLD experiment); 0 in the parity configuration */
} egn_info_t;
int egn_info(egn_ctx* ctx, egn_info_t* out);

/* Introspection used by tests / bench. */
long long egn_launch_count(egn_ctx* ctx);            /* kernels launched so far */
double egn_flops_per_frame(egn_ctx* ctx, int net);    /* algorithmic 2*MAC of the built graph */
/* Copies an internal activation (by debug name) to HOST fp32 [frames][C][H][W]; returns the
 * number of floats written, or -1.  dims (host, 3 ints) receives C,H,W.  out may be NULL to query. */
long long egn_debug_read(egn_ctx* ctx, const char* name, float* out, long long capacity, int frames,
                         int* dims);
/* Runs one convolution layer (by name) through both the tcgen05 kernel and its SIMT companion on
 * the layer's current inputs and returns the max abs difference of the outputs in *max_diff. */
int egn_conv_selfcheck(egn_ctx* ctx, const char* layer, int frames, double* max_diff, double* max_ref);

#ifdef __cplusplus
}
#endif
#endif /* EGN_H_ */
