/* libsnrf - C ABI of the B200-native SAM-NeRF feature-field renderer.
 *
 * The reference (WangFeng18/Segment-Anything-in-NeRF) has no FFI of its own: its hot path sits behind the
 * nerfstudio Python module surface and reaches native code through the `tinycudann` PyTorch extension
 * (call sites: samnerf/sam_field.py:51,63,84,99; nerfstudio/fields/nerfacto_field.py:144-175,228-240;
 * nerfstudio/fields/density_fields.py:92-100) plus ~60 ATen launches per chunk.  This header is the boundary a
 * drop-in replaces that with: plain pointers and sizes, int status codes, no torch types.  Every pointer
 * argument is a DEVICE pointer unless it says "host"; tensors are contiguous, row-major, rays-major.
 * All calls are stream-ordered on the `cudaStream_t` passed as `void* stream` (0 = default stream).
 * One context per (process, device); a context is not thread-safe (the reference serialises its two threads
 * with train_lock: nerfstudio/viewer/server/render_state_machine.py:190, nerfstudio/engine/trainer.py:222).
 *
 * Return value: 0 on success, otherwise a negative SNRF_E_* code; snrf_last_error() has the message.
 */
#ifndef SNRF_H_
#define SNRF_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SNRF_MAX_LEVELS 16

#define SNRF_OK 0
#define SNRF_E_INVALID (-1) /* bad argument / unsupported configuration */
#define SNRF_E_CUDA (-2)    /* a CUDA call failed                        */
#define SNRF_E_STATE (-3)   /* parameters for the requested output were never uploaded */

typedef struct snrf_ctx snrf_ctx;

/* One level of a tcnn HashGrid. */
typedef struct snrf_level {
  float scale;     /* exp2f(l*log2f(per_level_scale))*base - 1 */
  uint32_t res;    /* ceil(scale)+1                             */
  uint32_t size;   /* entries: min(round_up(res^3,8), 2^log2T)  */
  uint32_t offset; /* first entry of the level                  */
  uint32_t hashed; /* res^3 > size                              */
} snrf_level;

typedef struct snrf_grid_desc {
  int32_t n_levels;
  int32_t n_features; /* 2 (density fields) or 8 (feature field) */
  snrf_level lv[SNRF_MAX_LEVELS];
} snrf_grid_desc;

/* Fill a descriptor with tcnn's geometry rule.  Replaces the `encoding_config` dict handed to
 * tcnn.Encoding / tcnn.NetworkWithInputEncoding (samnerf/sam_field.py:99-109, nerfacto_field.py:160-167,
 * density_fields.py:73-81). */
int snrf_grid_desc_init(snrf_grid_desc* d, int n_levels, int n_features, int log2_hashmap_size, int base_resolution,
                        float per_level_scale);

/* ---- lifetime ------------------------------------------------------------------------------------ */
int snrf_ctx_create(int device, snrf_ctx** out);
void snrf_ctx_destroy(snrf_ctx* ctx);
const char* snrf_last_error(snrf_ctx* ctx); /* valid until the next call on ctx; ctx may be NULL */
/* tensor-core engine for the 192->256->{256,192} feature MLP and the conv head:
 * 1 = tcgen05.mma + TMEM (default), 0 = mma.sync (the recompiled-legacy comparison path). */
int snrf_set_engine(snrf_ctx* ctx, int engine);
/* Early termination of the nerfacto field along a ray (opt-in, default 0 = exact): when the transmittance left after
 * the first 16 of the 32 nerf samples is below eps, the remaining 16 samples are not evaluated and get weight 0.
 * Bounds: |rgb|, |accumulation| move by <= eps; the median depth and the top-k features are unaffected for
 * eps < 0.5 up to the eps-weighted tail; per-sample debug outputs of skipped samples read 0.  The reference never
 * terminates early, hence opt-in. */
int snrf_set_early_termination(snrf_ctx* ctx, float eps);
/* HBM budget (bytes) for "bricks": cell-major copies of the leading levels of the proposal and nerfacto grids, in which
 * the 8 corner entries of a cell are 32 contiguous bytes (one sector, one 256-bit load per sample and level instead of
 * 8 scattered gathers; csrc/bricks.cu).  Pure re-layout: rendered values are bit-identical with any budget.  The
 * longest prefix of levels that fits is used (nerfacto 16..2048 x 16 levels: 11 levels = 3.2 GiB, 12 = 8.5 GiB,
 * 13 = 22.5 GiB, 14 = 59.3 GiB); rebuilt by every snrf_upload_proposal / snrf_upload_field_base.  Default 4 GiB
 * (environment SNRF_BRICK_GB overrides), 0 = off.  Optionally returns the number of bricked levels.  Synchronises. */
int snrf_set_brick_budget(snrf_ctx* ctx, int64_t bytes, int* prop_levels, int* field_levels);
/* Feature samples below the precision of their own sum.  cutoff >= 0: rays are bucketed by the number of leading slots
 * whose sharpened, renormalised weight (sam_model.py:244-248) is >= cutoff (> 0 when cutoff == 0) and only 1 / 2 / 4 / 8 /
 * 16 slots of a ray are gathered and pushed through the MLP accordingly; the weights dropped per ray sum to
 * < 16 * cutoff.  DEFAULT 2^-24: what is dropped is less than one fp32 ulp of the accumulated feature (which is then
 * rounded to fp16), 3.0 instead of 16 slots per ray on the 800x800 benchmark frame; the whole parity suite runs through
 * it.  cutoff = 0 drops exact zeros only; cutoff < 0 evaluates every one of the 16 picked samples of every ray with the
 * un-bucketed kernel (csrc/sam.cu).  tcgen05 engine only (the mma.sync engine always evaluates every slot).
 * snrf_feature_backward applies the same rule to its rows.  See csrc/sam_bucket.cu. */
int snrf_set_feature_cutoff(snrf_ctx* ctx, float cutoff);
/* Rays the bucketed feature kernel has put into its 1 / 2 / 4 / 8 / 16-slot buckets since the last reset
 * (rays_per_bucket[5], host): sum_b rays[b] << b is the number of (ray, sample) slots actually gathered, 3 072 B each -
 * the algorithmic bytes bench.py states the roofline on when the cut-off is on.  Synchronises the device. */
int snrf_feature_slot_stats(snrf_ctx* ctx, int64_t* rays_per_bucket, int reset);
/* eval-mode PDF sample positions u[33] = linspace(0, 1-1/33, 33) + 1/66 (ray_samplers.py:325-327).  The
 * library computes the same table itself; a host may override it so that both sides share the bits.  n = 66:
 * followed by the 33 linspace values without the offset (the base of the training-mode positions, :314-322). */
int snrf_set_pdf_u(snrf_ctx* ctx, const float* u_host, int n);
/* Training-mode stratified sampling (ray_samplers.py:104-112,314-322 with single_jitter, the nerfacto default):
 * jitter[n_rays,2] (DEVICE) = per ray {t_rand of the initial sampler, rand of the PDF sampler}, uniform in [0,1) -
 * the caller draws them (torch.rand in the reference).  One-shot: applies to the NEXT snrf_render / snrf_sample
 * call on ctx, which must carry exactly n_rays rays, and is cleared by it.  NULL clears.  Not for
 * snrf_render_frame / snrf_render_camera (whole frames are eval-mode renders). */
int snrf_set_jitter(snrf_ctx* ctx, const float* jitter, int64_t n_rays);
/* Proposal-weight annealing of training (ProposalNetworkSampler.set_anneal, ray_samplers.py:546-548,583;
 * nerfacto.py:248-256): the PDF sampler sees weights^anneal.  Only calls that carry jitter (training-mode sampling)
 * use it; eval-mode calls always resample from the raw weights.  Default 1. */
int snrf_set_anneal(snrf_ctx* ctx, float anneal);

/* ---- parameters: flat fp32 tensors in tcnn order, host OR device pointers --------------------------
 * Each call converts to fp16 and packs into the kernels' layouts once (replaces tcnn's `params` tensors:
 * network weights first, then the grid).  Re-upload after every optimiser step if training continues. */
/* proposal_networks.0.mlp_base.params : MLP 16->16->16 then 5x2 grid (density_fields.py:50-100) */
int snrf_upload_proposal(snrf_ctx* ctx, const float* params, int64_t n, const snrf_grid_desc* grid, void* stream);
/* field.mlp_base.params : MLP 32->64->16 then 16x2 grid (nerfacto_field.py:157-175) */
int snrf_upload_field_base(snrf_ctx* ctx, const float* params, int64_t n, const snrf_grid_desc* grid, void* stream);
/* field.mlp_head.params : MLP 32->64->64->16 (nerfacto_field.py:228-240) */
int snrf_upload_field_head(snrf_ctx* ctx, const float* params, int64_t n, void* stream);
/* sam_field.clip_encs.{idx}.params (which=0) / sam_field.clipseg_encs.{idx}.params (which=1) */
int snrf_upload_feature_grid(snrf_ctx* ctx, int which, int idx, const float* params, int64_t n,
                             const snrf_grid_desc* grid, void* stream);
/* sam_field.sam_net.params (which=0, n_out=256) / sam_field.clipseg_net.params (which=1, n_out=192) */
int snrf_upload_feature_net(snrf_ctx* ctx, int which, const float* params, int64_t n, int n_out, void* stream);
/* conv_head.{0,2}.{weight,bias} : two Conv2d(256,256,3,pad 1) (sam_model.py:202-208) */
int snrf_upload_conv_head(snrf_ctx* ctx, const float* w0, const float* b0, const float* w2, const float* b2,
                          void* stream);

/* ---- the hot path -------------------------------------------------------------------------------- */
#define SNRF_WANT_SAM 1u     /* also render the 256-d SAM feature (sam_model.py:243-265)            */
#define SNRF_WANT_CLIPSEG 2u /* also render the 192-d ClipSeg feature (sam_model.py:272-277)         */
#define SNRF_PATCH 4u        /* push p x p ray patches through the conv head (sam_model.py:260-265)   */
#define SNRF_BG_LAST_SAMPLE 0
#define SNRF_BG_FIXED 1

typedef struct snrf_render_opts {
  float near_plane;   /* used when `nears` is NULL (eval: 0, scene_colliders.py:185) */
  float far_plane;    /* used when `fars` is NULL (1000, nerfacto.py:73)             */
  float hist_padding; /* PDFSampler histogram_padding (0.01)                         */
  int32_t bg_mode;    /* SNRF_BG_*                                                   */
  float bg[3];
  int32_t k_sam;      /* num_sam_samples (16)                                        */
  float sharpen;      /* sharpening_temperature (10)                                 */
  int32_t patch_size; /* p for SNRF_PATCH (4)                                        */
} snrf_render_opts;

/* optional per-stage outputs (NULL = skip); back the component shims and the staged parity tests */
typedef struct snrf_debug_out {
  float* prop_weights; /* [N,64]   proposal weights                      */
  float* edges;        /* [N,33]   nerf bin edges (euclidean t)          */
  float* weights;      /* [N,32]   nerf weights                          */
  float* density;      /* [N,32]   nerf densities                        */
  float* rgb_samples;  /* [N,32,3] per-sample colour                     */
  float* sam_t;        /* [N,k]    2 x midpoint t of the picked samples  */
  float* sam_w;        /* [N,k]    sharpened renormalised weights        */
  void* sam_feat;      /* [N,k,192] fp16 encoder output at those samples */
} snrf_debug_out;

/* SAMModel.forward / get_outputs for one chunk of rays in eval mode (samnerf/sam_model.py:226-314).
 * nears/fars: NULL or [N].  rgb[N,3] depth[N] are always written; acc[N], prop_depth[N] may be NULL ("fast");
 * sam: [N,256] or [N/p^2,256] with SNRF_PATCH; clipseg: [N,192]; dbg may be NULL. */
int snrf_render(snrf_ctx* ctx, const float* origins, const float* dirs, const float* nears, const float* fars,
                int64_t n_rays, uint32_t flags, const snrf_render_opts* opts, float* rgb, float* depth, float* acc,
                float* prop_depth, float* sam, float* clipseg, const snrf_debug_out* dbg, void* stream);

/* ProposalNetworkSampler.generate_ray_samples (ray_samplers.py:558-599): proposal weights [N,64], nerf bin
 * edges [N,33] and optionally the proposal median depth [N]. */
int snrf_sample(snrf_ctx* ctx, const float* origins, const float* dirs, const float* nears, const float* fars,
                int64_t n_rays, const snrf_render_opts* opts, float* prop_weights, float* edges, float* prop_depth,
                void* stream);

/* conv head on patch-major feature rows: feat_in[P*p*p,256] -> out[P,256] (sam_model.py:260-265) */
int snrf_patch_aggregate(snrf_ctx* ctx, const float* feat_in, int64_t n_patches, int p, float* out, void* stream);

/* ---- component-level entry points (back Field.density_fn / SAMField.get_outputs / renderers) ------- */
/* which: 0 = proposal field (density_fields.py:102-125), 1 = nerfacto field (nerfacto_field.py:242-266);
 * geo_f16: [n,15] fp16 or NULL (nerfacto only). */
int snrf_query_density(snrf_ctx* ctx, int which, const float* xyz, int64_t n, float* density, void* geo_f16,
                       void* stream);
/* nerfacto colour head (nerfacto_field.py:268-351): dirs[n,3], geo_f16[n,15] -> rgb[n,3] */
int snrf_query_rgb(snrf_ctx* ctx, const float* dirs, const void* geo_f16, int64_t n, float* rgb, void* stream);
/* SAMField.get_outputs per sample (sam_field.py:112-140): which 0 = sam, 1 = clipseg;
 * hashgrid_f16: [n,192] fp16 or NULL; out: [n,n_out] fp32 */
int snrf_query_features(snrf_ctx* ctx, int which, const float* xyz, int64_t n, void* hashgrid_f16, float* out,
                        void* stream);
/* ray-wise renderer ops: 0 get_weights(a=deltas,b=densities)->[N,S] (rays.py:141-163); 1 accumulation(a=w)->[N]
 * (renderers.py:197-223); 2 median depth(a=w,b=starts,c=ends)->[N] (renderers.py:260-270); 3 rgb(a=rgb[N,S,3],
 * b=w)->[N,3] (renderers.py:69-140); 4 mean(a=embeds[N,S,C],b=w)->[N,C] (sam_model.py:126-137) */
int snrf_ray_op(snrf_ctx* ctx, int mode, const float* a, const float* b, const float* c, float* out, int64_t n, int S,
                int C, int bg_mode, const float* bg_host, void* stream);

/* Whole frame (or this rank's tile of it): n_rays rays in chunks of `chunk` (the reference's
 * eval_num_rays_per_chunk, sam_model.py:354-364), same outputs as snrf_render over all rays.  When pipelining is
 * active the three stages of a chunk run on three streams inside the library, so chunk c+1's march overlaps chunk
 * c's feature gather and its output layer (whose replicated stores are NVLink-bound); the caller's stream resumes
 * only after every chunk is complete, so the call is stream-ordered like any other. */
int snrf_render_frame(snrf_ctx* ctx, const float* origins, const float* dirs, const float* nears, const float* fars,
                      int64_t n_rays, int64_t chunk, uint32_t flags, const snrf_render_opts* opts, float* rgb,
                      float* depth, float* acc, float* prop_depth, float* sam, float* clipseg, void* stream);
/* Element type of the feature rows the render calls write (sam [N,256], clipseg [N,192]; not the patch-aggregated rows):
 * 0 = fp32 (default, the dtype of the reference's outputs["sam"]), 1 = fp16 - the precision tinycudann itself emits per
 * sample (sam_field.py:51-61): the fp32 result of the output layer is rounded once (2^-11 relative, against a stated
 * feature tolerance of 2e-2).  Halves the bytes of the tile exchange between GPUs and of the device-to-host copy; the
 * `sam` / `clipseg` pointers of snrf_render / snrf_render_frame / snrf_render_camera are then read as __half*.
 * Copy-engine and push exchange only (not with the fused multicast / peer stores). */
int snrf_set_feature_dtype(snrf_ctx* ctx, int f16);
/* Frame calls: 1 = one march launch over the whole tile first (the picked samples of every ray go to a scratch of 128 B
 * per ray), then the feature kernels chunk by chunk - fewer, larger march launches when a tile is cut into small chunks
 * for the exchange; 0 (default) = march and features alternate per chunk. */
int snrf_set_march_first(snrf_ctx* ctx, int enable);
/* chunk pipelining of snrf_render_frame: 0 = off (strictly sequential on the caller's stream), 1 = auto (default:
 * only when outputs are replicated to other ranks - on one GPU the stages cannot share an SM, so it gains nothing),
 * 2 = always */
int snrf_set_pipeline(snrf_ctx* ctx, int enable);

/* ---- multi-GPU: fused tile all-gather (SURVEY.md 8 e; the reference has no multi-GPU render) -------------
 * When a frame is cut into row blocks across ranks, the kernels can store every output row they produce not only
 * into this rank's frame buffer but, at the same byte offset, into every other rank's frame buffer, so the
 * all-gather of the rendered tiles rides on the kernels' own stores over NVLink instead of a separate collective.
 * which: 0 sam[.,256], 1 rgb[.,3], 2 depth, 3 accumulation, 4 prop_depth.  local_base / bytes: this rank's frame
 * buffer (output pointers later passed to snrf_render* must lie inside it for the replication to apply).
 * mc_base: NVSwitch multicast alias of that buffer (one multimem.st reaches every rank) or NULL;
 * peer_bases_host[n_peers]: peer-mapped aliases on the other ranks, used when mc_base is NULL.  The caller
 * synchronises the ranks after the frame (e.g. a symmetric-memory barrier).  local_base NULL clears. */
int snrf_set_replication(snrf_ctx* ctx, int which, void* local_base, int64_t bytes, void* mc_base,
                         void* const* peer_bases_host, int n_peers);
/* How snrf_render_frame moves replicated outputs: 0 (default) = the kernels' own stores (multimem.st through
 * mc_base, else st.global through the peer pointers); 1 = copy engines: kernels write locally and each chunk's rows
 * are pushed to every peer pointer with cudaMemcpyAsync on side streams, so the exchange costs no SM time at all;
 * 2 = push kernel (csrc/exchange.cu): kernels write locally and a small kernel on a high-priority side stream stores each
 * chunk's feature rows into the other ranks' buffers - with ONE multimem.st per 16 bytes through mc_base when the "sam"
 * descriptor has both mc_base and peer pointers, else with one store per peer pointer (the small per-ray outputs always
 * go through copy engines once per frame).  Fastest measured mode at N = 8 (profiles/r02_multi_gpu.txt). */
int snrf_set_replication_mode(snrf_ctx* ctx, int mode); /* 0 fused stores, 1 copy engines, 2 push kernel (csrc/exchange.cu) */

/* ---- camera ray generation fused in front of the render -------------------------------
 * Replaces Cameras.generate_rays(camera_indices=i, keep_shape=True) for one camera
 * (nerfstudio/cameras/cameras.py:312-482,490-726; distortion: camera_utils.py:298-401) and the pixel sub-grid
 * the feature map is rendered on (samnerf/sam_model.py:368-379), so that a frame call takes ~100 bytes of camera
 * instead of 24 bytes per ray. */
#define SNRF_CAM_PERSPECTIVE 1 /* CameraType.PERSPECTIVE, cameras.py:42-47 */
#define SNRF_CAM_FISHEYE 2
#define SNRF_CAM_EQUIRECTANGULAR 3

typedef struct snrf_camera {
  float fx, fy, cx, cy;
  int32_t width, height;
  int32_t camera_type;    /* SNRF_CAM_*                                            */
  int32_t has_distortion; /* 0: distortion_params is None                          */
  float distortion[6];    /* k1 k2 k3 k4 p1 p2 (camera_utils.py:320-325)           */
  float c2w[12];          /* camera_to_worlds, row-major 3x4                       */
  int32_t has_aabb;       /* viewer crop box (generate_rays(aabb_box=...), cameras.py:463-482): nears / fars of every
                             ray = its intersection with the box (nerfstudio/utils/math.py:201-238)                 */
  float aabb[6];          /* x_min y_min z_min x_max y_max z_max                   */
} snrf_camera;

/* Rays through the pixel grid rows x cols of `cam` (HOST pointers; NULL = every row / column of the image, in
 * which case n_rows / n_cols must equal height / width).  Ray order: row-major when patch <= 1, else patch-major
 * over patch x patch blocks with row-major order inside a block (n_rows, n_cols divisible by patch) - the order
 * sam_model.py:376-379 produces.  origins[n,3], dirs[n,3] (unit), pixel_area[n] or NULL; nears[n] / fars[n]: written
 * when cam->has_aabb (else ignored, may be NULL); n = n_rows * n_cols. */
int snrf_generate_rays(snrf_ctx* ctx, const snrf_camera* cam, const int32_t* rows_host, int n_rows,
                       const int32_t* cols_host, int n_cols, int patch, float* origins, float* dirs,
                       float* pixel_area, float* nears, float* fars, void* stream);
/* snrf_generate_rays into library scratch followed by snrf_render_frame over those rays: one call per loop of
 * SAMModel.get_outputs_for_camera_ray_bundle (sam_model.py:354-418).  With SNRF_PATCH in `flags` the ray order is
 * patch-major with p = opts->patch_size and sam is [n/p^2, 256].  With cam->has_aabb the rays are rendered between
 * their crop-box intersections (the collider is bypassed, scene_colliders.py:40-44). */
int snrf_render_camera(snrf_ctx* ctx, const snrf_camera* cam, const int32_t* rows_host, int n_rows,
                       const int32_t* cols_host, int n_cols, int64_t chunk, uint32_t flags,
                       const snrf_render_opts* opts, float* rgb, float* depth, float* acc, float* prop_depth,
                       float* sam, float* clipseg, void* stream);

/* ---- training side of the feature-field branch (first slice) ---------------------------
 * The `sam_field` parameter group (samnerf/sam_model.py:330-335) receives gradients only through
 * SAMField.get_outputs at the k picked samples (positions detached, sam_field.py:116) and the MeanRenderer with
 * detached weights (sam_model.py:258-277), so its forward / backward are closed over (origins, dirs, sam_t, sam_w) -
 * the picks snrf_render reports through snrf_debug_out.  which: 0 = sam (n_out 256), 1 = clipseg (n_out 192). */
/* out[N,n_out] = W2 . sum_k w_k fp16(relu(W1 x_k)); enc_f16: [N,16,192] fp16 encoder outputs, saved for the
 * backward pass (may be NULL when no backward follows).  Same kernels as the render path. */
int snrf_feature_forward(snrf_ctx* ctx, int which, const float* origins, const float* dirs, const float* sam_t,
                         const float* sam_w, int64_t n_rays, float* out, void* enc_f16, void* stream);
/* Given d_out[N,n_out], ACCUMULATE (+=, like torch's .grad) fp32 gradients in the reference's flat layouts:
 * grad_net [256*192 + n_out*256] for sam_field.{sam,clipseg}_net.params (layer-1 matrix, then layer-2),
 * grad_grid0 / grad_grid1 [entries*8] for sam_field.{clip,clipseg}_encs.{0,1}.params (16-byte aligned: the table
 * scatter uses vector reductions).  Any of the three may be NULL (that parameter is frozen).  Replaces tinycudann's autograd for these modules (sam_field.py:51,63,84,99). */
int snrf_feature_backward(snrf_ctx* ctx, int which, const float* origins, const float* dirs, const float* sam_t,
                          const float* sam_w, int64_t n_rays, const float* d_out, const void* enc_f16,
                          float* grad_net, float* grad_grid0, float* grad_grid1, void* stream);

/* Backward of snrf_patch_aggregate (conv head, sam_model.py:202-208,260-265; torch autograd + cuDNN in the reference):
 * feat_in[P*16,256] as given to the forward call, d_out[P,256].  grad_w0 / grad_w2 [256*256*9] and grad_b0 / grad_b2
 * [256] in torch's Conv2d layouts, ACCUMULATED (+=); d_feat[P*16,256] WRITTEN (NULL = not needed). */
int snrf_patch_aggregate_backward(snrf_ctx* ctx, const float* feat_in, int64_t n_patches, int p, const float* d_out,
                                  float* grad_w0, float* grad_b0, float* grad_w2, float* grad_b2, float* d_feat,
                                  void* stream);

/* Backward of the density fields at arbitrary sample positions - the counterpart of snrf_query_density /
 * snrf_query_rgb, i.e. what tinycudann's autograd does for HashMLPDensityField (density_fields.py:92-125) and
 * TCNNNerfactoField (nerfacto_field.py:157-175,228-351) while torch autograd handles weights, compositing and the
 * losses on [N,S] tensors.  which: 0 = proposal field, 1 = nerfacto field.  xyz[n,3] world positions; dirs[n,3]
 * (nerfacto with d_rgb, else NULL); d_density[n] / d_rgb[n,3]: upstream gradients, either may be NULL.
 * grad_base: flat fp32 gradient of proposal_networks.0.mlp_base.params / field.mlp_base.params (MLP matrices, then the
 * grid; 16-byte aligned), grad_head: of field.mlp_head.params (needed with d_rgb).  Gradients are ACCUMULATED (+=).
 * trunc_exp backward is g * exp(clamp(x, -15, 15)) (activations.py:33-37). */
int snrf_field_backward(snrf_ctx* ctx, int which, const float* xyz, const float* dirs, int64_t n,
                        const float* d_density, const float* d_rgb, float* grad_base, float* grad_head, void* stream);

/* Top-k pick + sharpening of the feature samples as a stand-alone ray op (samnerf/sam_model.py:244-255; the fused
 * snrf_render does the same inside the march kernel): weights / starts / ends [N,S] -> sam_t[N,k] = start + end of the
 * picked samples, sam_w[N,k] = w^sharpen renormalised, slots in descending weight order (ties by sample index). */
int snrf_pick_samples(snrf_ctx* ctx, const float* weights, const float* starts, const float* ends, int64_t n, int S,
                      int k, float sharpen, float* sam_t, float* sam_w, void* stream);

/* Backward of the two ray-wise ops that carry gradients in training (torch autograd in the reference):
 * mode 0  RaySamples.get_weights (rays.py:141-163): a = deltas[N,S], b = densities[N,S], g = dL/dweights[N,S]
 *         -> out_a = dL/ddensities[N,S] (out_b unused; deltas are detached, ray_samplers.py:357);  S <= 64
 * mode 3  RGBRenderer.combine_rgb (renderers.py:69-112): a = rgb[N,S,3], b = weights[N,S], g = dL/drgb_out[N,3]
 *         -> out_a = dL/drgb[N,S,3], out_b = dL/dweights[N,S]; bg_mode / bg_host as in snrf_ray_op.
 * Outputs are WRITTEN (not accumulated). */
int snrf_ray_op_backward(snrf_ctx* ctx, int mode, const float* a, const float* b, const float* g, float* out_a,
                         float* out_b, int64_t n, int S, int bg_mode, const float* bg_host, void* stream);

/* number of kernels this library has launched on ctx since creation (bench.py's gpu_launches) */
int64_t snrf_launch_count(snrf_ctx* ctx);
/* Bracket the three hot kernels of snrf_render with CUDA events on the launching stream (bench.py's roofline). */
int snrf_set_timing(snrf_ctx* ctx, int enable);
/* Synchronise on the recorded events; ms_out[3] / count_out[3] = accumulated duration and launches of
 * {0: per-ray march, 1: feature gather + first MLP layer, 2: tap GEMM} since the previous call. */
int snrf_kernel_times(snrf_ctx* ctx, double* ms_out, int64_t* count_out);

#ifdef __cplusplus
}
#endif
#endif /* SNRF_H_ */
