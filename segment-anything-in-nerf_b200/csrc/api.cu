// extern "C" entry points of libsnrf (see include/snrf.h): context, parameter packing, the render call and the
// component-level queries.  Host code only; every kernel lives in march.cu / sam.cu / gemm.cu / query.cu.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/snrf.h"
#include "kernels.cuh"
#include "raygen.cuh"
#include "backward.cuh"

using namespace snrf;

namespace {

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  cudaError_t ensure(size_t need) {
    if (need <= bytes) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
    cudaError_t e = cudaMalloc(&p, need);
    if (e == cudaSuccess) bytes = need;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
  }
  template <typename T>
  T* as() const {
    return reinterpret_cast<T*>(p);
  }
};

constexpr int kCopyStreams = 32;  // kMaxPeers destinations x up to 4 pieces

struct Replication {
  float* local = nullptr;
  size_t bytes = 0;
  float* mc = nullptr;
  float* peer[kMaxPeers] = {nullptr};
  int n = 0;
};

struct FeatureNet {
  DevBuf table[2];
  GridDev grid[2];
  bool have_grid[2] = {false, false};
  DevBuf w1_core, w2_core;  // tensor-core layouts
  DevBuf w1_rm, w2_rm;      // row-major fp16 (component queries)
  int n_out = 0;
  bool have_net = false;
};

}  // namespace

struct snrf_ctx {
  int device = 0;
  int sm_count = 148;
  int engine = 1;
  // cell-major copies of the leading grid levels for the march kernel (snrf_set_brick_budget, bricks.cu)
  int64_t brick_budget = 4ll << 30;
  DevBuf prop_brick_store, field_brick_store;
  BrickDev prop_bricks = {}, field_bricks = {};
  bool march_v1 = false;  // SNRF_MARCH=v1: the round-1 march kernel (A/B measurements only)
  // Two alternative structures of the march kernel, both parity-green on hardware and both measured SLOWER than the fused
  // mma.sync kernel on the headline frame (gpurun calls 14 / 15: 4.09 ms fused, 4.47 tcgen05 MLPs, 4.38 split), kept as
  // opt-in variants for A/B measurements: SNRF_MARCH_TC=1, SNRF_MARCH_SPLIT=1 (see the notes in march.cu).
  bool march_tc = false;
  bool march_split = false;
  DevBuf edges;             // [N,33] bin edges between the two halves
  float et_eps = 0.f;  // snrf_set_early_termination
  // snrf_set_feature_cutoff: >= 0 = bucketed kernel B' (default 2^-24: slots below one fp32 ulp of the ray's weight sum
  // are not evaluated), < 0 = kernel B on every slot
  float feat_cutoff = 5.9604645e-8f;
  const float* jitter = nullptr;  // snrf_set_jitter: training-mode draws for the next render / sample call
  int64_t jitter_rays = 0;
  float anneal = 1.f;  // snrf_set_anneal
  std::string err;
  int64_t launches = 0;
  // proposal field
  DevBuf prop_table, prop_w1f, prop_w2f, prop_w1_rm, prop_w2_rm;
  GridDev prop_grid;
  bool have_prop = false;
  // nerfacto field
  DevBuf field_table, base_w1_rm, base_w2_rm, head_w1_rm, head_w2_rm, head_w3_rm, wfrag, head_perm, wcore;
  GridDev field_grid;
  bool have_base = false, have_head = false;
  // feature nets: 0 = sam, 1 = clipseg
  FeatureNet feat[2];
  // conv head
  DevBuf conv_w[2], conv_b[2];
  DevBuf conv_w_rm[2];  // [256 x 2304] row-major fp16 = torch's Conv2d layout (backward pass)
  bool have_conv = false;
  // per-kernel CUDA-event timing (bench.py's roofline): 0 march, 1 feature gather+MLP1, 2 tap GEMM
  bool timing = false;
  std::vector<cudaEvent_t> ev_pool;
  std::vector<std::pair<int, std::pair<cudaEvent_t, cudaEvent_t>>> ev_used;
  double k_ms[3] = {0, 0, 0};
  int64_t k_count[3] = {0, 0, 0};
  // fused tile all-gather (snrf_set_replication): 0 sam, 1 rgb, 2 depth, 3 accumulation, 4 prop_depth
  Replication rep[5];
  int rep_mode = 0;  // 0: the kernels' own stores (multimem.st / peer pointers); 1: copy engines (cudaMemcpyAsync per peer);
                     // 2: a small push kernel (peer stores from a few CTAs on a side stream, see launch_push_rows)
  int dma_split = 1;      // copy-engine mode: pieces per destination and chunk (more copies in flight; SNRF_DMA_SPLIT)
  bool rows_f16 = false;  // snrf_set_feature_dtype: the sam / clipseg output rows are fp16 instead of fp32
  bool march_first = false;  // snrf_set_march_first: frame calls march the whole tile in one launch, then the feature chunks
  // one copy stream per destination (kMaxPeers): the copy engines serve the peers of a chunk concurrently
  cudaStream_t copy_stream[kCopyStreams] = {nullptr};
  cudaEvent_t ev_copy[kCopyStreams] = {nullptr}, ev_out[2] = {nullptr, nullptr};
  // frame-level pipelining (snrf_render_frame)
  int pipeline = 1;  // 0 off, 1 auto (only when outputs are replicated to other ranks), 2 always
  bool aux_ready = false;
  cudaStream_t aux_feat = nullptr, aux_out = nullptr;
  cudaEvent_t ev_march[2] = {nullptr, nullptr}, ev_feat[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};
  cudaEvent_t ev_feat_done[2] = {nullptr, nullptr}, ev_out_done[2] = {nullptr, nullptr}, ev_join = nullptr, ev_join2 = nullptr;
  // misc
  DevBuf pdf_u;
  DevBuf sam_t[2], sam_w[2], hbar[2][2], feat_f16, hid_f16, q_feat, q_h1, q_h2, q_sel, q_x;
  DevBuf cam_rows, cam_cols, cam_o, cam_d, cam_near, cam_far;  // snrf_generate_rays / snrf_render_camera
  DevBuf bwd_scratch, bwd_sink;             // snrf_feature_backward
  DevBuf bucket_lists[2], bucket_counts[2]; // bucketed feature kernel, per pipeline slot
  DevBuf bucket_totals;                     // rays per bucket since the last snrf_feature_slot_stats reset
};

namespace {

int fail(snrf_ctx* ctx, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (ctx) ctx->err = buf;
  return code;
}
std::string g_null_err = "null context";

#define CK(expr)                                                                                        \
  do {                                                                                                  \
    cudaError_t e__ = (expr);                                                                           \
    if (e__ != cudaSuccess) return fail(ctx, SNRF_E_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e__)); \
  } while (0)
#define LAUNCH(expr) \
  do {               \
    CK(expr);        \
    ctx->launches++; \
  } while (0)

cudaEvent_t get_event(snrf_ctx* ctx) {
  if (!ctx->ev_pool.empty()) {
    cudaEvent_t e = ctx->ev_pool.back();
    ctx->ev_pool.pop_back();
    return e;
  }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}
// LAUNCH, bracketed by events on the launching stream when timing is enabled
#define TIMED_LAUNCH(kid, s, expr)                                   \
  do {                                                               \
    cudaEvent_t e0__ = nullptr, e1__ = nullptr;                      \
    if (ctx->timing) {                                               \
      e0__ = get_event(ctx);                                         \
      e1__ = get_event(ctx);                                         \
      cudaEventRecord(e0__, s);                                      \
    }                                                                \
    LAUNCH(expr);                                                    \
    if (ctx->timing) {                                               \
      cudaEventRecord(e1__, s);                                      \
      ctx->ev_used.push_back({kid, {e0__, e1__}});                   \
    }                                                                \
  } while (0)

GridDev to_dev(const snrf_grid_desc* d, const __half* table) {
  GridDev g;
  memset(&g, 0, sizeof(g));
  g.table = table;
  g.n_levels = d->n_levels;
  g.n_features = d->n_features;
  for (int l = 0; l < d->n_levels; ++l) {
    g.lv[l].scale = d->lv[l].scale;
    g.lv[l].res = d->lv[l].res;
    g.lv[l].size = d->lv[l].size;
    g.lv[l].offset = d->lv[l].offset;
    g.lv[l].hashed = d->lv[l].hashed;
  }
  return g;
}
int64_t grid_entries(const snrf_grid_desc* d) {
  return static_cast<int64_t>(d->lv[d->n_levels - 1].offset) + d->lv[d->n_levels - 1].size;
}
bool grid_ok(const snrf_grid_desc* d, int levels, int feats) {
  if (!d || d->n_levels != levels || d->n_features != feats) return false;
  for (int l = 0; l < levels; ++l) {
    if (d->lv[l].size == 0 || d->lv[l].res < 2) return false;
    // the kernels replace `% size` by a mask on hashed levels (tcnn: hashed size == 2^log2_hashmap_size)
    if (d->lv[l].hashed && (d->lv[l].size & (d->lv[l].size - 1u)) != 0) return false;
  }
  return true;
}

// stage `n` floats (host or device source) in a temporary device buffer
int stage(snrf_ctx* ctx, const float* src, int64_t n, DevBuf& tmp, cudaStream_t s) {
  CK(tmp.ensure(static_cast<size_t>(n) * sizeof(float)));
  CK(cudaMemcpyAsync(tmp.p, src, static_cast<size_t>(n) * sizeof(float), cudaMemcpyDefault, s));
  return SNRF_OK;
}

// (Re)build the bricks of one F = 2 grid: the longest prefix of levels whose bricks fit `budget` bytes.
// Returns the bytes used through *used.
int build_bricks(snrf_ctx* ctx, const GridDev& G, int64_t budget, DevBuf& store, BrickDev& B, int64_t* used, cudaStream_t s) {
  memset(&B, 0, sizeof(B));
  *used = 0;
  int n = 0;
  size_t total = 0;
  size_t off[kMaxLevels];
  for (int l = 0; l < G.n_levels; ++l) {
    const uint64_t r = G.lv[l].res;
    const uint64_t cells = r * r * r;
    if (cells > 0xFFFFFFFFull / 2) break;  // cell indices (and 2 x cell) stay 32-bit in the kernel
    const size_t bytes = static_cast<size_t>(cells) * 32;
    if (static_cast<int64_t>(total + bytes) > budget) break;
    off[l] = total;
    total += bytes;
    n = l + 1;
  }
  if (n == 0) {
    store.release();
    return SNRF_OK;
  }
  if (store.bytes < total || store.bytes > total + (64u << 20)) store.release();  // shrink when the budget drops
  CK(store.ensure(total));
  for (int l = 0; l < n; ++l) {
    uint4* dst = reinterpret_cast<uint4*>(static_cast<char*>(store.p) + off[l]);
    LAUNCH(launch_brick_build(G, l, dst, s));
    B.lv[l] = dst;
  }
  B.n = n;
  *used = static_cast<int64_t>(total);
  return SNRF_OK;
}

int rebuild_all_bricks(snrf_ctx* ctx, cudaStream_t s) {
  int64_t used_p = 0, used_f = 0;
  if (ctx->have_prop) {
    int rc = build_bricks(ctx, ctx->prop_grid, ctx->brick_budget, ctx->prop_brick_store, ctx->prop_bricks, &used_p, s);
    if (rc) return rc;
  }
  if (ctx->have_base) {
    int rc = build_bricks(ctx, ctx->field_grid, ctx->brick_budget - used_p, ctx->field_brick_store, ctx->field_bricks, &used_f, s);
    if (rc) return rc;
  }
  CK(cudaStreamSynchronize(s));
  return SNRF_OK;
}

void default_pdf_u(float* u, int n_bins) {
  // u[0..n):   torch.linspace(0, 1 - 1/n_bins, n_bins) (float32, symmetric fill) + 1/(2 n_bins)   ray_samplers.py:325-327
  // u[n..2n):  the same linspace without the offset (training mode adds rand / n_bins instead, :314-322)
  const float end = static_cast<float>(1.0 - (1.0 / n_bins));
  const float step = (end - 0.f) / static_cast<float>(n_bins - 1);
  const int half = n_bins / 2;
  const float off = static_cast<float>(1.0 / (2 * n_bins));
  for (int i = 0; i < n_bins; ++i) {
    const float v = i < half ? 0.f + step * static_cast<float>(i) : end - step * static_cast<float>(n_bins - i - 1);
    u[i] = v + off;
    u[n_bins + i] = v;
  }
}

}  // namespace

extern "C" {

int snrf_grid_desc_init(snrf_grid_desc* d, int n_levels, int n_features, int log2_hashmap_size, int base_resolution,
                        float per_level_scale) {
  if (!d || n_levels < 1 || n_levels > SNRF_MAX_LEVELS) return SNRF_E_INVALID;
  memset(d, 0, sizeof(*d));
  d->n_levels = n_levels;
  d->n_features = n_features;
  const float log2_pls = log2f(per_level_scale);
  uint32_t offset = 0;
  for (int l = 0; l < n_levels; ++l) {
    const float scale = exp2f(static_cast<float>(l) * log2_pls) * static_cast<float>(base_resolution) - 1.0f;
    const uint32_t res = static_cast<uint32_t>(ceilf(scale)) + 1u;
    const uint64_t dense = static_cast<uint64_t>(res) * res * res;
    uint64_t size = (dense + 7) / 8 * 8;
    const uint64_t cap = 1ull << log2_hashmap_size;
    if (size > cap) size = cap;
    d->lv[l].scale = scale;
    d->lv[l].res = res;
    d->lv[l].size = static_cast<uint32_t>(size);
    d->lv[l].offset = offset;
    d->lv[l].hashed = dense > size ? 1u : 0u;
    offset += static_cast<uint32_t>(size);
  }
  return SNRF_OK;
}

int snrf_ctx_create(int device, snrf_ctx** out) {
  if (!out) return SNRF_E_INVALID;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) return SNRF_E_CUDA;
  if (cudaSetDevice(device) != cudaSuccess) return SNRF_E_CUDA;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return SNRF_E_CUDA;
  snrf_ctx* ctx = new snrf_ctx();
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  if (const char* e = getenv("SNRF_MARCH")) ctx->march_v1 = strcmp(e, "v1") == 0;
  if (const char* e = getenv("SNRF_MARCH_TC")) ctx->march_tc = atoi(e) != 0;
  if (const char* e = getenv("SNRF_MARCH_SPLIT")) ctx->march_split = atoi(e) != 0;
  if (const char* e = getenv("SNRF_BRICK_GB")) ctx->brick_budget = static_cast<int64_t>(atof(e) * (1ll << 30));
  if (prop.major != 10) {
    // this library is compiled for sm_100a only; refuse loudly instead of failing at the first launch
    fprintf(stderr, "libsnrf: device %d is sm_%d%d, built for sm_100a\n", device, prop.major, prop.minor);
    delete ctx;
    return SNRF_E_INVALID;
  }
  float u[66];
  default_pdf_u(u, 33);
  if (ctx->pdf_u.ensure(sizeof(u)) != cudaSuccess ||
      cudaMemcpy(ctx->pdf_u.p, u, sizeof(u), cudaMemcpyHostToDevice) != cudaSuccess) {
    delete ctx;
    return SNRF_E_CUDA;
  }
  cudaEvent_t* evs[] = {&ctx->ev_march[0], &ctx->ev_march[1], &ctx->ev_feat[0][0], &ctx->ev_feat[0][1], &ctx->ev_feat[1][0],
                        &ctx->ev_feat[1][1], &ctx->ev_feat_done[0], &ctx->ev_feat_done[1], &ctx->ev_out_done[0],
                        &ctx->ev_out_done[1], &ctx->ev_join, &ctx->ev_join2};
  for (cudaEvent_t* e : evs)
    if (cudaEventCreateWithFlags(e, cudaEventDisableTiming) != cudaSuccess) {
      delete ctx;
      return SNRF_E_CUDA;
    }
  *out = ctx;
  return SNRF_OK;
}

void snrf_ctx_destroy(snrf_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  DevBuf* bufs[] = {&ctx->prop_table, &ctx->prop_w1f,   &ctx->prop_w2f,   &ctx->prop_w1_rm, &ctx->prop_w2_rm,
                    &ctx->field_table, &ctx->base_w1_rm, &ctx->base_w2_rm, &ctx->head_w1_rm, &ctx->head_w2_rm,
                    &ctx->head_w3_rm, &ctx->wfrag,      &ctx->head_perm,  &ctx->pdf_u,      &ctx->sam_t[0],
                    &ctx->sam_w[0],   &ctx->hbar[0][0], &ctx->feat_f16,   &ctx->hid_f16,    &ctx->q_feat,
                    &ctx->q_h1,       &ctx->q_h2,       &ctx->q_sel,      &ctx->q_x,        &ctx->conv_w[0],
                    &ctx->conv_w[1],  &ctx->conv_b[0],  &ctx->conv_b[1]};
  for (DevBuf* b : bufs) b->release();
  ctx->sam_t[1].release(); ctx->sam_w[1].release(); ctx->wcore.release(); ctx->edges.release();
  ctx->cam_rows.release(); ctx->cam_cols.release(); ctx->cam_o.release(); ctx->cam_d.release();
  ctx->cam_near.release(); ctx->cam_far.release();
  ctx->conv_w_rm[0].release(); ctx->conv_w_rm[1].release();
  ctx->bwd_scratch.release(); ctx->bwd_sink.release();
  for (int i = 0; i < 2; ++i) { ctx->bucket_lists[i].release(); ctx->bucket_counts[i].release(); }
  ctx->bucket_totals.release();
  ctx->prop_brick_store.release(); ctx->field_brick_store.release();
  ctx->hbar[0][1].release(); ctx->hbar[1][0].release(); ctx->hbar[1][1].release();
  if (ctx->aux_feat) cudaStreamDestroy(ctx->aux_feat);
  if (ctx->aux_out) cudaStreamDestroy(ctx->aux_out);
  for (int i = 0; i < kCopyStreams; ++i) {
    if (ctx->copy_stream[i]) cudaStreamDestroy(ctx->copy_stream[i]);
    if (ctx->ev_copy[i]) cudaEventDestroy(ctx->ev_copy[i]);
  }
  cudaEvent_t own[] = {ctx->ev_out[0], ctx->ev_out[1], ctx->ev_march[0], ctx->ev_march[1], ctx->ev_feat[0][0],
                       ctx->ev_feat[0][1], ctx->ev_feat[1][0], ctx->ev_feat[1][1], ctx->ev_feat_done[0],
                       ctx->ev_feat_done[1], ctx->ev_out_done[0], ctx->ev_out_done[1], ctx->ev_join, ctx->ev_join2};
  for (cudaEvent_t e : own)
    if (e) cudaEventDestroy(e);
  for (auto& u : ctx->ev_used) {
    cudaEventDestroy(u.second.first);
    cudaEventDestroy(u.second.second);
  }
  for (cudaEvent_t e : ctx->ev_pool) cudaEventDestroy(e);
  for (int w = 0; w < 2; ++w) {
    FeatureNet& f = ctx->feat[w];
    f.table[0].release(); f.table[1].release();
    f.w1_core.release(); f.w2_core.release(); f.w1_rm.release(); f.w2_rm.release();
  }
  delete ctx;
}

const char* snrf_last_error(snrf_ctx* ctx) { return ctx ? ctx->err.c_str() : g_null_err.c_str(); }

int snrf_set_engine(snrf_ctx* ctx, int engine) {
  if (!ctx || (engine != 0 && engine != 1)) return fail(ctx, SNRF_E_INVALID, "engine must be 0 (mma.sync) or 1 (tcgen05)");
  ctx->engine = engine;
  return SNRF_OK;
}

int snrf_set_early_termination(snrf_ctx* ctx, float eps) {
  if (!ctx || !(eps >= 0.f) || eps >= 0.5f) return fail(ctx, SNRF_E_INVALID, "early-termination threshold must be in [0, 0.5)");
  ctx->et_eps = eps;
  return SNRF_OK;
}

int snrf_set_brick_budget(snrf_ctx* ctx, int64_t bytes, int* prop_levels, int* field_levels) {
  if (!ctx || bytes < 0) return fail(ctx, SNRF_E_INVALID, "brick budget must be >= 0 bytes");
  CK(cudaSetDevice(ctx->device));
  CK(cudaDeviceSynchronize());  // no render may still be reading the old bricks
  ctx->brick_budget = bytes;
  int rc = rebuild_all_bricks(ctx, nullptr);
  if (rc) return rc;
  if (prop_levels) *prop_levels = ctx->prop_bricks.n;
  if (field_levels) *field_levels = ctx->field_bricks.n;
  return SNRF_OK;
}

int snrf_set_feature_cutoff(snrf_ctx* ctx, float cutoff) {
  if (!ctx || cutoff != cutoff || cutoff > 1e-2f) return fail(ctx, SNRF_E_INVALID, "feature cut-off must be < 0 (off) or in [0, 1e-2]");
  ctx->feat_cutoff = cutoff;
  return SNRF_OK;
}

int snrf_feature_slot_stats(snrf_ctx* ctx, int64_t* rays_per_bucket, int reset) {
  if (!ctx || !rays_per_bucket) return fail(ctx, SNRF_E_INVALID, "null argument");
  CK(cudaSetDevice(ctx->device));
  unsigned long long host[kFeatBuckets] = {0, 0, 0, 0, 0};
  if (ctx->bucket_totals.p) {
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(host, ctx->bucket_totals.p, sizeof(host), cudaMemcpyDeviceToHost));
    if (reset) CK(cudaMemset(ctx->bucket_totals.p, 0, sizeof(host)));
  }
  for (int b = 0; b < kFeatBuckets; ++b) rays_per_bucket[b] = static_cast<int64_t>(host[b]);
  return SNRF_OK;
}

int snrf_set_pdf_u(snrf_ctx* ctx, const float* u_host, int n) {
  if (!ctx || !u_host || (n != 33 && n != 66)) return fail(ctx, SNRF_E_INVALID, "pdf_u must have 33 (or 33 + 33) entries");
  CK(cudaSetDevice(ctx->device));
  CK(cudaMemcpy(ctx->pdf_u.p, u_host, n * sizeof(float), cudaMemcpyHostToDevice));
  return SNRF_OK;
}

int snrf_set_anneal(snrf_ctx* ctx, float anneal) {
  if (!ctx || !(anneal >= 0.f) || anneal > 1.f) return fail(ctx, SNRF_E_INVALID, "anneal must be in [0, 1]");
  ctx->anneal = anneal;
  return SNRF_OK;
}

int snrf_set_jitter(snrf_ctx* ctx, const float* jitter, int64_t n_rays) {
  if (!ctx || n_rays < 0) return fail(ctx, SNRF_E_INVALID, "bad argument");
  ctx->jitter = jitter;
  ctx->jitter_rays = jitter ? n_rays : 0;
  return SNRF_OK;
}

int64_t snrf_launch_count(snrf_ctx* ctx) { return ctx ? ctx->launches : 0; }

int snrf_set_replication(snrf_ctx* ctx, int which, void* local_base, int64_t bytes, void* mc_base,
                         void* const* peer_bases_host, int n_peers) {
  if (!ctx) return SNRF_E_INVALID;
  if (which < 0 || which > 4 || n_peers < 0 || n_peers > kMaxPeers || (n_peers > 0 && !peer_bases_host) || bytes < 0)
    return fail(ctx, SNRF_E_INVALID, "bad replication descriptor");
  Replication& R = ctx->rep[which];
  R = Replication();
  if (!local_base || (n_peers == 0 && !mc_base)) return SNRF_OK;  // cleared
  R.local = reinterpret_cast<float*>(local_base);
  R.bytes = static_cast<size_t>(bytes);
  R.mc = reinterpret_cast<float*>(mc_base);
  R.n = n_peers;
  for (int i = 0; i < n_peers; ++i) R.peer[i] = reinterpret_cast<float*>(peer_bases_host[i]);
  return SNRF_OK;
}

int snrf_set_replication_mode(snrf_ctx* ctx, int mode) {
  if (!ctx || mode < 0 || mode > 2)
    return fail(ctx, SNRF_E_INVALID, "replication mode must be 0 (fused stores), 1 (copy engines) or 2 (push kernel)");
  ctx->rep_mode = mode;
  if (const char* e = getenv("SNRF_DMA_SPLIT")) ctx->dma_split = atoi(e) < 1 ? 1 : (atoi(e) > 4 ? 4 : atoi(e));
  return SNRF_OK;
}

int snrf_set_feature_dtype(snrf_ctx* ctx, int f16) {
  if (!ctx || (f16 != 0 && f16 != 1)) return fail(ctx, SNRF_E_INVALID, "feature dtype must be 0 (fp32) or 1 (fp16)");
  ctx->rows_f16 = f16 != 0;
  return SNRF_OK;
}

int snrf_set_march_first(snrf_ctx* ctx, int enable) {
  if (!ctx) return SNRF_E_INVALID;
  ctx->march_first = enable != 0;
  return SNRF_OK;
}

int snrf_set_pipeline(snrf_ctx* ctx, int enable) {
  if (!ctx) return SNRF_E_INVALID;
  ctx->pipeline = enable < 0 ? 0 : (enable > 2 ? 2 : enable);
  return SNRF_OK;
}

int snrf_set_timing(snrf_ctx* ctx, int enable) {
  if (!ctx) return SNRF_E_INVALID;
  ctx->timing = enable != 0;
  return SNRF_OK;
}

int snrf_kernel_times(snrf_ctx* ctx, double* ms_out, int64_t* count_out) {
  if (!ctx || !ms_out || !count_out) return fail(ctx, SNRF_E_INVALID, "null argument");
  CK(cudaSetDevice(ctx->device));
  for (auto& u : ctx->ev_used) {
    CK(cudaEventSynchronize(u.second.second));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, u.second.first, u.second.second));
    ctx->k_ms[u.first] += ms;
    ctx->k_count[u.first] += 1;
    ctx->ev_pool.push_back(u.second.first);
    ctx->ev_pool.push_back(u.second.second);
  }
  ctx->ev_used.clear();
  for (int k = 0; k < 3; ++k) {
    ms_out[k] = ctx->k_ms[k];
    count_out[k] = ctx->k_count[k];
    ctx->k_ms[k] = 0;
    ctx->k_count[k] = 0;
  }
  return SNRF_OK;
}

// ---------------------------------------------------------------------------------------------------
// uploads
// ---------------------------------------------------------------------------------------------------
int snrf_upload_proposal(snrf_ctx* ctx, const float* params, int64_t n, const snrf_grid_desc* grid, void* stream) {
  if (!ctx || !params) return fail(ctx, SNRF_E_INVALID, "null argument");
  if (!grid_ok(grid, 5, 2)) return fail(ctx, SNRF_E_INVALID, "proposal grid must be 5 levels x 2 features");
  const int64_t n_net = 16 * 16 + 16 * 16, n_grid = grid_entries(grid) * 2;
  if (n != n_net + n_grid) return fail(ctx, SNRF_E_INVALID, "proposal params: got %lld, expected %lld", (long long)n, (long long)(n_net + n_grid));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  CK(cudaSetDevice(ctx->device));
  DevBuf tmp;
  int rc = stage(ctx, params, n, tmp, s);
  if (rc) return rc;
  const float* src = tmp.as<float>();
  CK(ctx->prop_table.ensure(n_grid * 2));
  CK(ctx->prop_w1f.ensure(16 * 17 * 4));
  CK(ctx->prop_w2f.ensure(16 * 4));
  CK(ctx->prop_w1_rm.ensure(256 * 2));
  CK(ctx->prop_w2_rm.ensure(256 * 2));
  CK(cudaMemsetAsync(ctx->prop_w1f.p, 0, 16 * 17 * 4, s));
  LAUNCH(launch_f32_to_f16(src + n_net, ctx->prop_table.as<__half>(), n_grid, s));
  CK(ctx->wfrag.ensure(kMarchFragTiles * 256));
  LAUNCH(launch_pack_frag(src, 16, nullptr, ctx->wfrag.as<uint2>() + kFragProp1 * 32, 2, 1, s));
  LAUNCH(launch_pack_frag(src + 256, 16, nullptr, ctx->wfrag.as<uint2>() + kFragProp2 * 32, 1, 1, s));
  LAUNCH(launch_f32_to_f16(src, ctx->prop_w1_rm.as<__half>(), 256, s));
  LAUNCH(launch_f32_to_f16(src + 256, ctx->prop_w2_rm.as<__half>(), 256, s));
  CK(cudaStreamSynchronize(s));
  tmp.release();
  ctx->prop_grid = to_dev(grid, ctx->prop_table.as<__half>());
  ctx->have_prop = true;
  return rebuild_all_bricks(ctx, s);
}

int snrf_upload_field_base(snrf_ctx* ctx, const float* params, int64_t n, const snrf_grid_desc* grid, void* stream) {
  if (!ctx || !params) return fail(ctx, SNRF_E_INVALID, "null argument");
  if (!grid_ok(grid, 16, 2)) return fail(ctx, SNRF_E_INVALID, "field grid must be 16 levels x 2 features");
  const int64_t n_net = 64 * 32 + 16 * 64, n_grid = grid_entries(grid) * 2;
  if (n != n_net + n_grid) return fail(ctx, SNRF_E_INVALID, "field base params: got %lld, expected %lld", (long long)n, (long long)(n_net + n_grid));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  CK(cudaSetDevice(ctx->device));
  DevBuf tmp;
  int rc = stage(ctx, params, n, tmp, s);
  if (rc) return rc;
  const float* src = tmp.as<float>();
  CK(ctx->field_table.ensure(n_grid * 2));
  CK(ctx->base_w1_rm.ensure(64 * 32 * 2));
  CK(ctx->base_w2_rm.ensure(16 * 64 * 2));
  CK(ctx->wfrag.ensure(kMarchFragTiles * 256));
  LAUNCH(launch_f32_to_f16(src + n_net, ctx->field_table.as<__half>(), n_grid, s));
  LAUNCH(launch_f32_to_f16(src, ctx->base_w1_rm.as<__half>(), 64 * 32, s));
  LAUNCH(launch_f32_to_f16(src + 64 * 32, ctx->base_w2_rm.as<__half>(), 16 * 64, s));
  uint2* wf = ctx->wfrag.as<uint2>();
  LAUNCH(launch_pack_frag(src, 32, nullptr, wf + kFragBase1 * 32, 8, 2, s));
  LAUNCH(launch_pack_frag(src + 64 * 32, 64, nullptr, wf + kFragBase2 * 32, 2, 4, s));
  // the same two layers as tcgen05 B operands (march_kernel<.., TC>; byte plan in march.cu)
  CK(ctx->wcore.ensure(kMarchCoreBytes));
  LAUNCH(launch_pack_core(src, ctx->wcore.as<__half>(), 64, 32, s));
  LAUNCH(launch_pack_core(src + 64 * 32, ctx->wcore.as<__half>() + 64 * 32, 16, 64, s));
  CK(cudaStreamSynchronize(s));
  tmp.release();
  ctx->field_grid = to_dev(grid, ctx->field_table.as<__half>());
  ctx->have_base = true;
  return rebuild_all_bricks(ctx, s);
}

int snrf_upload_field_head(snrf_ctx* ctx, const float* params, int64_t n, void* stream) {
  if (!ctx || !params) return fail(ctx, SNRF_E_INVALID, "null argument");
  const int64_t expect = 64 * 32 + 64 * 64 + 16 * 64;
  if (n != expect) return fail(ctx, SNRF_E_INVALID, "field head params: got %lld, expected %lld", (long long)n, (long long)expect);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  CK(cudaSetDevice(ctx->device));
  DevBuf tmp;
  int rc = stage(ctx, params, n, tmp, s);
  if (rc) return rc;
  const float* src = tmp.as<float>();
  CK(ctx->head_w1_rm.ensure(64 * 32 * 2));
  CK(ctx->head_w2_rm.ensure(64 * 64 * 2));
  CK(ctx->head_w3_rm.ensure(16 * 64 * 2));
  CK(ctx->wfrag.ensure(kMarchFragTiles * 256));
  CK(ctx->head_perm.ensure(32 * sizeof(int)));
  // kernel-side input order [pad, geo0..14, SH0..15]  <-  tcnn order [SH0..15, geo0..14, pad]
  int perm[32];
  perm[0] = 31;
  for (int c = 1; c < 16; ++c) perm[c] = 15 + c;
  for (int c = 0; c < 16; ++c) perm[16 + c] = c;
  CK(cudaMemcpyAsync(ctx->head_perm.p, perm, sizeof(perm), cudaMemcpyHostToDevice, s));
  LAUNCH(launch_f32_to_f16(src, ctx->head_w1_rm.as<__half>(), 64 * 32, s));
  LAUNCH(launch_f32_to_f16(src + 64 * 32, ctx->head_w2_rm.as<__half>(), 64 * 64, s));
  LAUNCH(launch_f32_to_f16(src + 64 * 32 + 64 * 64, ctx->head_w3_rm.as<__half>(), 16 * 64, s));
  uint2* wf = ctx->wfrag.as<uint2>();
  LAUNCH(launch_pack_frag(src, 32, ctx->head_perm.as<int>(), wf + kFragHead1 * 32, 8, 2, s));
  LAUNCH(launch_pack_frag(src + 64 * 32, 64, nullptr, wf + kFragHead2 * 32, 8, 4, s));
  LAUNCH(launch_pack_frag(src + 64 * 32 + 64 * 64, 64, nullptr, wf + kFragHead3 * 32, 1, 4, s));
  CK(ctx->wcore.ensure(kMarchCoreBytes));
  __half* wc = ctx->wcore.as<__half>() + 64 * 32 + 16 * 64;
  LAUNCH(launch_pack_core(src, wc, 64, 32, s, ctx->head_perm.as<int>()));
  LAUNCH(launch_pack_core(src + 64 * 32, wc + 64 * 32, 64, 64, s));
  LAUNCH(launch_pack_core(src + 64 * 32 + 64 * 64, wc + 64 * 32 + 64 * 64, 16, 64, s));
  CK(cudaStreamSynchronize(s));
  tmp.release();
  ctx->have_head = true;
  return SNRF_OK;
}

int snrf_upload_feature_grid(snrf_ctx* ctx, int which, int idx, const float* params, int64_t n,
                             const snrf_grid_desc* grid, void* stream) {
  if (!ctx || !params || which < 0 || which > 1 || idx < 0 || idx > 1) return fail(ctx, SNRF_E_INVALID, "bad argument");
  if (!grid_ok(grid, 12, 8)) return fail(ctx, SNRF_E_INVALID, "feature grid must be 12 levels x 8 features");
  const int64_t n_grid = grid_entries(grid) * 8;
  if (n != n_grid) return fail(ctx, SNRF_E_INVALID, "feature grid params: got %lld, expected %lld", (long long)n, (long long)n_grid);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  CK(cudaSetDevice(ctx->device));
  DevBuf tmp;
  int rc = stage(ctx, params, n, tmp, s);
  if (rc) return rc;
  FeatureNet& f = ctx->feat[which];
  CK(f.table[idx].ensure(n_grid * 2));
  LAUNCH(launch_f32_to_f16(tmp.as<float>(), f.table[idx].as<__half>(), n_grid, s));
  CK(cudaStreamSynchronize(s));
  tmp.release();
  f.grid[idx] = to_dev(grid, f.table[idx].as<__half>());
  f.have_grid[idx] = true;
  return SNRF_OK;
}

int snrf_upload_feature_net(snrf_ctx* ctx, int which, const float* params, int64_t n, int n_out, void* stream) {
  if (!ctx || !params || which < 0 || which > 1) return fail(ctx, SNRF_E_INVALID, "bad argument");
  if (n_out != 256 && n_out != 192) return fail(ctx, SNRF_E_INVALID, "feature net output width must be 256 or 192");
  const int64_t expect = 256 * 192 + static_cast<int64_t>(n_out) * 256;
  if (n != expect) return fail(ctx, SNRF_E_INVALID, "feature net params: got %lld, expected %lld", (long long)n, (long long)expect);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  CK(cudaSetDevice(ctx->device));
  DevBuf tmp;
  int rc = stage(ctx, params, n, tmp, s);
  if (rc) return rc;
  const float* src = tmp.as<float>();
  FeatureNet& f = ctx->feat[which];
  CK(f.w1_core.ensure(256 * 192 * 2));
  CK(f.w2_core.ensure(static_cast<size_t>(n_out) * 256 * 2));
  CK(f.w1_rm.ensure(256 * 192 * 2));
  CK(f.w2_rm.ensure(static_cast<size_t>(n_out) * 256 * 2));
  LAUNCH(launch_pack_core(src, f.w1_core.as<__half>(), 256, 192, s));
  LAUNCH(launch_pack_core(src + 256 * 192, f.w2_core.as<__half>(), n_out, 256, s));
  LAUNCH(launch_f32_to_f16(src, f.w1_rm.as<__half>(), 256 * 192, s));
  LAUNCH(launch_f32_to_f16(src + 256 * 192, f.w2_rm.as<__half>(), static_cast<int64_t>(n_out) * 256, s));
  CK(cudaStreamSynchronize(s));
  tmp.release();
  f.n_out = n_out;
  f.have_net = true;
  return SNRF_OK;
}

int snrf_upload_conv_head(snrf_ctx* ctx, const float* w0, const float* b0, const float* w2, const float* b2,
                          void* stream) {
  if (!ctx || !w0 || !b0 || !w2 || !b2) return fail(ctx, SNRF_E_INVALID, "null argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  CK(cudaSetDevice(ctx->device));
  // Conv2d weight [out=256][in=256][3][3] -> 9 per-tap [out x in] matrices, each in core-matrix layout.
  // The tap split is done on the host side of the staging copy to keep the packing kernel generic.
  const float* ws[2] = {w0, w2};
  const float* bs[2] = {b0, b2};
  const int64_t n_w = 256LL * 256 * 9;
  std::vector<float> host(n_w), taps(n_w);
  for (int c = 0; c < 2; ++c) {
    CK(cudaMemcpyAsync(host.data(), ws[c], n_w * sizeof(float), cudaMemcpyDefault, s));
    CK(cudaStreamSynchronize(s));
    for (int o = 0; o < 256; ++o)
      for (int i = 0; i < 256; ++i)
        for (int t = 0; t < 9; ++t) taps[(static_cast<size_t>(t) * 256 + o) * 256 + i] = host[(static_cast<size_t>(o) * 256 + i) * 9 + t];
    DevBuf tmp;
    int rc = stage(ctx, host.data(), n_w, tmp, s);
    if (rc) return rc;
    CK(ctx->conv_w_rm[c].ensure(n_w * 2));
    LAUNCH(launch_f32_to_f16(tmp.as<float>(), ctx->conv_w_rm[c].as<__half>(), n_w, s));
    CK(cudaStreamSynchronize(s));
    rc = stage(ctx, taps.data(), n_w, tmp, s);
    if (rc) return rc;
    CK(ctx->conv_w[c].ensure(n_w * 2));
    for (int t = 0; t < 9; ++t)
      LAUNCH(launch_pack_core(tmp.as<float>() + static_cast<size_t>(t) * 65536,
                              ctx->conv_w[c].as<__half>() + static_cast<size_t>(t) * 65536, 256, 256, s));
    CK(ctx->conv_b[c].ensure(256 * sizeof(float)));
    CK(cudaMemcpyAsync(ctx->conv_b[c].p, bs[c], 256 * sizeof(float), cudaMemcpyDefault, s));
    CK(cudaStreamSynchronize(s));
    tmp.release();
  }
  ctx->have_conv = true;
  return SNRF_OK;
}

// ---------------------------------------------------------------------------------------------------
// render
// ---------------------------------------------------------------------------------------------------
static int fill_march(snrf_ctx* ctx, MarchParams& M, const float* origins, const float* dirs, const float* nears,
                      const float* fars, int64_t n, const snrf_render_opts* o) {
  if (!ctx->have_prop) return fail(ctx, SNRF_E_STATE, "proposal parameters not uploaded");
  memset(&M, 0, sizeof(M));
  M.origins = origins;
  M.dirs = dirs;
  M.nears = nears;
  M.fars = fars;
  M.n_rays = n;
  M.near_default = o->near_plane;
  M.far_default = o->far_plane;
  M.prop = ctx->prop_grid;
  M.field = ctx->field_grid;
  M.prop_bricks = ctx->prop_bricks;
  M.field_bricks = ctx->field_bricks;
  M.wfrag = ctx->wfrag.as<uint2>();
  M.wcore = ctx->have_base && ctx->have_head ? ctx->wcore.as<uint4>() : nullptr;
  M.use_tc = ctx->march_tc ? 1 : 0;
  if (ctx->march_split && !ctx->march_tc) {  // all march launches of a context are ordered on one stream: one buffer
    CK(ctx->edges.ensure(static_cast<size_t>(n) * 33 * sizeof(float)));
    M.edges = ctx->edges.as<float>();
    M.split = 1;
  }
  M.pdf_u = ctx->pdf_u.as<float>();
  M.hist_padding = o->hist_padding;
  M.bg_mode = o->bg_mode == SNRF_BG_FIXED ? kBgFixed : kBgLastSample;
  M.bg[0] = o->bg[0];
  M.bg[1] = o->bg[1];
  M.bg[2] = o->bg[2];
  M.k_sam = o->k_sam;
  M.sharpen = o->sharpen;
  M.et_eps = ctx->et_eps;
  M.pdf_u_base = ctx->pdf_u.as<float>() + 33;
  M.anneal = ctx->anneal;
  if (ctx->jitter) {  // one-shot: consumed by this call
    const float* j = ctx->jitter;
    const int64_t nj = ctx->jitter_rays;
    ctx->jitter = nullptr;
    ctx->jitter_rays = 0;
    if (nj != n) return fail(ctx, SNRF_E_INVALID, "snrf_set_jitter was given %lld rays, the call has %lld", (long long)nj, (long long)n);
    M.jitter = j;
  }
  return SNRF_OK;
}

int snrf_patch_aggregate(snrf_ctx* ctx, const float* feat_in, int64_t n_patches, int p, float* out, void* stream) {
  if (!ctx || !feat_in || !out) return fail(ctx, SNRF_E_INVALID, "null argument");
  if (p != 4) return fail(ctx, SNRF_E_INVALID, "patch_size %d unsupported (the conv head kernel is built for p = 4)", p);
  if (!ctx->have_conv) return fail(ctx, SNRF_E_STATE, "conv head parameters not uploaded");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  CK(cudaSetDevice(ctx->device));
  const int64_t m = n_patches * 16;
  if (m == 0) return SNRF_OK;
  CK(ctx->feat_f16.ensure(m * 256 * 2));
  CK(ctx->hid_f16.ensure(m * 256 * 2));
  LAUNCH(launch_f32_to_f16(feat_in, ctx->feat_f16.as<__half>(), m * 256, s));
  GemmParams G;
  memset(&G, 0, sizeof(G));
  G.a = ctx->feat_f16.as<__half>();
  G.w = ctx->conv_w[0].as<__half>();
  G.bias = ctx->conv_b[0].as<float>();
  G.out_f16 = ctx->hid_f16.as<__half>();
  G.m = m; G.n = 256; G.taps = 9; G.relu = 1; G.out_mode = 1;
  LAUNCH(launch_tapgemm(G, ctx->engine == 1, ctx->sm_count, s));
  G.a = ctx->hid_f16.as<__half>();
  G.w = ctx->conv_w[1].as<__half>();
  G.bias = ctx->conv_b[1].as<float>();
  G.out_f16 = nullptr;
  G.out_f32 = out;
  G.relu = 0; G.out_mode = 2;
  LAUNCH(launch_tapgemm(G, ctx->engine == 1, ctx->sm_count, s));
  return SNRF_OK;
}

// Streams of one chunk.  Plain snrf_render runs everything on the caller's stream; snrf_render_frame spreads the
// three stages over three streams so that chunk c+1's march overlaps chunk c's feature gather and output layer.
struct ChunkStreams {
  cudaStream_t march, feat, out;
  int slot;        // which copy of the per-chunk scratch (sam_t / sam_w / hbar) to use
  int out_grid;    // CTA cap of the output-layer kernel (0 = one per SM)
  bool small_cta;  // 8-warp / 145 KB feature kernel that shares an SM with march CTAs
  cudaEvent_t after_march;  // recorded right behind the march launch when not null (the per-ray outputs are complete there)
  int stage;           // 0: march + features; 1: march only (picks go to the scratch); 2: features only (picks come from it)
  int64_t scratch_off; // first ray of this call inside the sam_t / sam_w scratch (stage 1 / 2: the scratch spans the tile)
};

// offset-preserving aliases of an output pointer in the other ranks' frame buffers (snrf_set_replication)
static void replicate(const snrf_ctx* ctx, int which, const void* out, int64_t n_bytes, float** mc, float** peers,
                      int* n_peers) {
  *mc = nullptr;
  *n_peers = 0;
  const Replication& R = ctx->rep[which];
  if (ctx->rep_mode != 0) return;  // copy-engine / push modes: kernels write locally, snrf_render_frame moves the rows
  const char* lo = reinterpret_cast<const char*>(R.local);
  const char* p = reinterpret_cast<const char*>(out);
  if (!R.local || !out || p < lo || p + n_bytes > lo + R.bytes) return;
  const int64_t off = p - lo;
  if (R.mc) {
    *mc = reinterpret_cast<float*>(reinterpret_cast<char*>(R.mc) + off);
  } else {
    *n_peers = R.n;
    for (int i = 0; i < R.n; ++i) peers[i] = reinterpret_cast<float*>(reinterpret_cast<char*>(R.peer[i]) + off);
  }
}

static int render_chunk(snrf_ctx* ctx, const float* origins, const float* dirs, const float* nears, const float* fars,
                        int64_t n_rays, uint32_t flags, const snrf_render_opts* opts, float* rgb, float* depth, float* acc,
                        float* prop_depth, void* sam, void* clipseg, const snrf_debug_out* dbg, const ChunkStreams& cs) {
  if (n_rays == 0) return SNRF_OK;  // empty chunk: nothing to read or write, pointers may be null
  if (!origins || !dirs || !opts || !rgb || !depth) return fail(ctx, SNRF_E_INVALID, "null argument");
  if (!ctx->have_base || !ctx->have_head) return fail(ctx, SNRF_E_STATE, "nerfacto field parameters not uploaded");
  const bool want_sam = (flags & SNRF_WANT_SAM) != 0, want_clip = (flags & SNRF_WANT_CLIPSEG) != 0;
  const bool patch = (flags & SNRF_PATCH) != 0;
  if (want_sam && !sam) return fail(ctx, SNRF_E_INVALID, "SNRF_WANT_SAM without an output buffer");
  if (want_clip && !clipseg) return fail(ctx, SNRF_E_INVALID, "SNRF_WANT_CLIPSEG without an output buffer");
  if ((want_sam || want_clip) && opts->k_sam != 16)
    return fail(ctx, SNRF_E_INVALID, "num_sam_samples %d unsupported (the feature kernel is built for k = 16)", opts->k_sam);
  if (patch && (!want_sam || opts->patch_size != 4 || n_rays % 16 != 0))
    return fail(ctx, SNRF_E_INVALID, "SNRF_PATCH needs SNRF_WANT_SAM, patch_size 4 and a ray count divisible by 16");
  const int slot = cs.slot;

  MarchParams M;
  int rc = fill_march(ctx, M, origins, dirs, nears, fars, n_rays, opts);
  if (rc) return rc;
  M.rgb = rgb;
  M.depth = depth;
  M.acc = acc;
  M.prop_depth = prop_depth;
  replicate(ctx, 1, rgb, n_rays * 12, &M.rep[0].mc, M.rep[0].peer, &M.rep[0].n);
  replicate(ctx, 2, depth, n_rays * 4, &M.rep[1].mc, M.rep[1].peer, &M.rep[1].n);
  replicate(ctx, 3, acc, n_rays * 4, &M.rep[2].mc, M.rep[2].peer, &M.rep[2].n);
  replicate(ctx, 4, prop_depth, n_rays * 4, &M.rep[3].mc, M.rep[3].peer, &M.rep[3].n);
  const bool feats = want_sam || want_clip;
  if (feats || (dbg && dbg->sam_t)) {
    CK(ctx->sam_t[slot].ensure((cs.scratch_off + n_rays) * 16 * 4));
    CK(ctx->sam_w[slot].ensure((cs.scratch_off + n_rays) * 16 * 4));
    M.sam_t = ctx->sam_t[slot].as<float>() + cs.scratch_off * 16;
    M.sam_w = ctx->sam_w[slot].as<float>() + cs.scratch_off * 16;
  }
  if (dbg) {
    M.dbg_w0 = dbg->prop_weights;
    M.dbg_edges = dbg->edges;
    M.dbg_weights = dbg->weights;
    M.dbg_density = dbg->density;
    M.dbg_rgb = dbg->rgb_samples;
  }
  if (cs.stage != 2)
    TIMED_LAUNCH(0, cs.march, ctx->march_v1 ? launch_march_v1(M, ctx->sm_count, cs.march) : launch_march(M, ctx->sm_count, cs.march));
  if (cs.after_march) CK(cudaEventRecord(cs.after_march, cs.march));
  if (cs.stage == 1) return SNRF_OK;
  if (dbg && dbg->sam_t) {
    CK(cudaMemcpyAsync(dbg->sam_t, M.sam_t, n_rays * 16 * 4, cudaMemcpyDeviceToDevice, cs.march));
    if (dbg->sam_w) CK(cudaMemcpyAsync(dbg->sam_w, M.sam_w, n_rays * 16 * 4, cudaMemcpyDeviceToDevice, cs.march));
  }
  if (!feats) return SNRF_OK;
  if (cs.feat != cs.march) {
    CK(cudaEventRecord(ctx->ev_march[slot], cs.march));
    CK(cudaStreamWaitEvent(cs.feat, ctx->ev_march[slot], 0));
  }
  for (int which = 0; which < 2; ++which) {
    if (!(which == 0 ? want_sam : want_clip)) continue;
    FeatureNet& f = ctx->feat[which];
    if (!f.have_grid[0] || !f.have_grid[1] || !f.have_net)
      return fail(ctx, SNRF_E_STATE, "%s parameters not uploaded", which == 0 ? "sam_field" : "clipseg");
    DevBuf& hbar = ctx->hbar[slot][which];
    CK(hbar.ensure(n_rays * 256 * 2));
    SamParams S;
    memset(&S, 0, sizeof(S));
    S.origins = origins;
    S.dirs = dirs;
    S.sam_t = M.sam_t;
    S.sam_w = M.sam_w;
    S.n_rays = n_rays;
    S.enc[0] = f.grid[0];
    S.enc[1] = f.grid[1];
    S.w1 = f.w1_core.as<__half>();
    S.hbar = hbar.as<__half>();
    S.dbg_feat = (which == 0 && dbg) ? reinterpret_cast<__half*>(dbg->sam_feat) : nullptr;
    if (ctx->feat_cutoff >= 0.f && ctx->engine == 1 && !S.dbg_feat) {
      // bucketed variant (sam_bucket.cu): pre-pass + one launch per bucket, all on the feature stream
      CK(ctx->bucket_lists[slot].ensure(static_cast<size_t>(kFeatBuckets) * n_rays * sizeof(int)));
      CK(ctx->bucket_counts[slot].ensure((kFeatBuckets + 1) * sizeof(int)));
      SamBucketParams Bk;
      memset(&Bk, 0, sizeof(Bk));
      Bk.origins = origins; Bk.dirs = dirs; Bk.sam_t = M.sam_t; Bk.sam_w = M.sam_w;
      Bk.enc[0] = f.grid[0]; Bk.enc[1] = f.grid[1];
      Bk.w1 = f.w1_core.as<__half>();
      Bk.hbar = hbar.as<__half>();
      cudaEvent_t e0 = nullptr, e1 = nullptr;
      if (ctx->timing) {
        e0 = get_event(ctx); e1 = get_event(ctx);
        cudaEventRecord(e0, cs.feat);
      }
      if (!ctx->bucket_totals.p) {
        CK(ctx->bucket_totals.ensure(kFeatBuckets * sizeof(unsigned long long)));
        CK(cudaMemsetAsync(ctx->bucket_totals.p, 0, kFeatBuckets * sizeof(unsigned long long), cs.feat));
      }
      LAUNCH(launch_bucket_assign(M.sam_w, ctx->feat_cutoff, ctx->bucket_counts[slot].as<int>(),
                                  ctx->bucket_lists[slot].as<int>(), n_rays, ctx->bucket_totals.as<unsigned long long>(),
                                  cs.feat));
      Bk.lists = ctx->bucket_lists[slot].as<int>();
      Bk.counts = ctx->bucket_counts[slot].as<int>();
      Bk.n_rays = n_rays;
      LAUNCH(launch_sam_bucketed(Bk, ctx->sm_count, cs.feat));
      if (ctx->timing) {
        cudaEventRecord(e1, cs.feat);
        ctx->ev_used.push_back({1, {e0, e1}});
      }
    } else {
      TIMED_LAUNCH(1, cs.feat, launch_sam(S, ctx->engine == 1, cs.small_cta, ctx->sm_count, cs.feat));
    }
    if (cs.out != cs.feat) {
      CK(cudaEventRecord(ctx->ev_feat[slot][which], cs.feat));
      CK(cudaStreamWaitEvent(cs.out, ctx->ev_feat[slot][which], 0));
    }
    GemmParams G;
    memset(&G, 0, sizeof(G));
    G.a = hbar.as<__half>();
    G.w = f.w2_core.as<__half>();
    G.m = n_rays;
    G.n = f.n_out;
    G.taps = 1;
    const int out_sms = cs.out_grid > 0 && cs.out_grid < ctx->sm_count ? cs.out_grid : ctx->sm_count;
    const bool to_patch = which == 0 && patch;
    if (to_patch) {
      CK(ctx->feat_f16.ensure(n_rays * 256 * 2));
      G.out_f16 = ctx->feat_f16.as<__half>();
      G.out_mode = 1;
    } else if (ctx->rows_f16 && !patch) {
      // fp16 output rows (snrf_set_feature_dtype): tcnn's own output precision; only the copy-engine / push exchange
      if (ctx->rep_mode == 0 && which == 0 && ctx->rep[0].local && (ctx->rep[0].mc || ctx->rep[0].n > 0))
        return fail(ctx, SNRF_E_INVALID, "fp16 feature rows cannot be combined with the fused multicast / peer stores");
      G.out_f16 = reinterpret_cast<__half*>(which == 0 ? sam : clipseg);
      G.out_mode = 1;
    } else {
      G.out_f32 = reinterpret_cast<float*>(which == 0 ? sam : clipseg);
      G.out_mode = 0;
      // fused tile all-gather: mirror the rows into the other ranks' frame buffers at the same offset
      if (which == 0) replicate(ctx, 0, sam, n_rays * 256 * 4, &G.out_mc, G.out_peer, &G.n_peers);
    }
    TIMED_LAUNCH(2, cs.out, launch_tapgemm(G, ctx->engine == 1, out_sms, cs.out));
    if (to_patch) {
      if (!ctx->have_conv) return fail(ctx, SNRF_E_STATE, "conv head parameters not uploaded");
      CK(ctx->hid_f16.ensure(n_rays * 256 * 2));
      GemmParams C;
      memset(&C, 0, sizeof(C));
      C.a = ctx->feat_f16.as<__half>();
      C.w = ctx->conv_w[0].as<__half>();
      C.bias = ctx->conv_b[0].as<float>();
      C.out_f16 = ctx->hid_f16.as<__half>();
      C.m = n_rays; C.n = 256; C.taps = 9; C.relu = 1; C.out_mode = 1;
      LAUNCH(launch_tapgemm(C, ctx->engine == 1, ctx->sm_count, cs.out));
      C.a = ctx->hid_f16.as<__half>();
      C.w = ctx->conv_w[1].as<__half>();
      C.bias = ctx->conv_b[1].as<float>();
      C.out_f16 = nullptr;
      C.out_f32 = reinterpret_cast<float*>(sam);  // patch-aggregated rows are always fp32
      C.relu = 0; C.out_mode = 2;
      LAUNCH(launch_tapgemm(C, ctx->engine == 1, ctx->sm_count, cs.out));
    }
  }
  return SNRF_OK;
}

int snrf_render(snrf_ctx* ctx, const float* origins, const float* dirs, const float* nears, const float* fars,
                int64_t n_rays, uint32_t flags, const snrf_render_opts* opts, float* rgb, float* depth, float* acc,
                float* prop_depth, float* sam, float* clipseg, const snrf_debug_out* dbg, void* stream) {
  if (!ctx) return SNRF_E_INVALID;
  if (n_rays < 0) return fail(ctx, SNRF_E_INVALID, "negative ray count");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  CK(cudaSetDevice(ctx->device));
  const ChunkStreams cs = {s, s, s, 0, 0, false, nullptr, 0, 0};
  return render_chunk(ctx, origins, dirs, nears, fars, n_rays, flags, opts, rgb, depth, acc, prop_depth, sam, clipseg, dbg, cs);
}

int snrf_render_frame(snrf_ctx* ctx, const float* origins, const float* dirs, const float* nears, const float* fars,
                      int64_t n_rays, int64_t chunk, uint32_t flags, const snrf_render_opts* opts, float* rgb,
                      float* depth, float* acc, float* prop_depth, float* sam, float* clipseg, void* stream) {
  if (!ctx) return SNRF_E_INVALID;
  if (n_rays < 0 || chunk <= 0) return fail(ctx, SNRF_E_INVALID, "bad ray count / chunk size");
  if ((flags & SNRF_PATCH) && chunk % 16 != 0) return fail(ctx, SNRF_E_INVALID, "chunk must be a multiple of 16 with SNRF_PATCH");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  CK(cudaSetDevice(ctx->device));
  if (!ctx->aux_ready) {
    int lo = 0, hi = 0;
    CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CK(cudaStreamCreateWithPriority(&ctx->aux_feat, cudaStreamNonBlocking, hi));
    CK(cudaStreamCreateWithPriority(&ctx->aux_out, cudaStreamNonBlocking, hi));
    for (int i = 0; i < kCopyStreams; ++i) {
      CK(cudaStreamCreateWithPriority(&ctx->copy_stream[i], cudaStreamNonBlocking, hi));
      CK(cudaEventCreateWithFlags(&ctx->ev_copy[i], cudaEventDisableTiming));
    }
    CK(cudaEventCreateWithFlags(&ctx->ev_out[0], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&ctx->ev_out[1], cudaEventDisableTiming));
    ctx->aux_ready = true;
  }
  const bool feats = (flags & (SNRF_WANT_SAM | SNRF_WANT_CLIPSEG)) != 0;
  // rows leave through copy engines (mode 1) or the push kernel (mode 2): any output registered with a destination
  bool moved = false;
  for (int w = 0; w < 5; ++w) moved = moved || (ctx->rep_mode != 0 && ctx->rep[w].local && ctx->rep[w].n > 0);
  const bool move_sam = moved && ctx->rep[0].local && ctx->rep[0].n > 0;
  // auto: pipeline only when a kernel's own replicated stores are NVLink-bound; with copy engines the kernels run
  // back to back on one stream (cross-stream hops cost ~20 us per chunk and buy nothing when the exchange is off the SMs)
  const bool pipelined = feats && n_rays > chunk && (ctx->pipeline == 2 || (ctx->pipeline == 1 && ctx->rep[0].local && !moved));
  const int p2 = (flags & SNRF_PATCH) ? 16 : 1;
  const size_t esz = ctx->rows_f16 && !(flags & SNRF_PATCH) ? 2 : 4;  // bytes per feature element (patch rows stay fp32)
  // the aux streams start after everything already queued on the caller's stream
  if (pipelined) {
    CK(cudaEventRecord(ctx->ev_join, s));
    CK(cudaStreamWaitEvent(ctx->aux_feat, ctx->ev_join, 0));
    CK(cudaStreamWaitEvent(ctx->aux_out, ctx->ev_join, 0));
  }
  // with replicated outputs the output-layer kernel is throttled by NVLink, not by the SMs: keep its footprint small
  // (multicast: one store reaches every rank, a handful of SMs saturate the link; peer pointers: n_peers stores)
  int out_grid = 0;
  if (pipelined && ctx->rep[0].local && !moved) {
    out_grid = 32;
    if (const char* e = getenv("SNRF_OUT_GRID")) out_grid = atoi(e);
  }
  // march-first (snrf_set_march_first): one march launch over the whole tile - the picks of every ray go to the scratch -
  // then the feature kernels chunk by chunk.  Fewer, larger march launches; the bulk rows still leave chunk by chunk.
  const bool march_first = ctx->march_first && feats && !pipelined && n_rays > chunk;
  if (march_first) {
    ChunkStreams cs = {s, s, s, 0, 0, false, moved ? ctx->ev_join : nullptr, 1, 0};
    int rc = render_chunk(ctx, origins, dirs, nears, fars, n_rays, flags, opts, rgb, depth, acc, prop_depth, sam, clipseg, nullptr, cs);
    if (rc) return rc;
  }
  uint32_t used = 0;  // copy streams that carry work of this frame
  int64_t c = 0;
  for (int64_t i = 0; i < n_rays; i += chunk, ++c) {
    const int64_t n = n_rays - i < chunk ? n_rays - i : chunk;
    const int slot = pipelined ? static_cast<int>(c & 1) : 0;
    // rgb / depth / accumulation / proposal depth of the whole tile are complete behind the LAST chunk's march kernel,
    // so their copies start there and overlap that chunk's feature kernels
    const bool last_chunk = i + chunk >= n_rays;
    ChunkStreams cs = {s, pipelined ? ctx->aux_feat : s, pipelined ? ctx->aux_out : s, slot, out_grid,
                       pipelined && getenv("SNRF_CORESIDENT") != nullptr,
                       (moved && last_chunk && !march_first) ? ctx->ev_join : nullptr, march_first ? 2 : 0, march_first ? i : 0};
    if (pipelined && c >= 2) {
      CK(cudaStreamWaitEvent(s, ctx->ev_feat_done[slot], 0));             // sam_t / sam_w of this slot are free again
      CK(cudaStreamWaitEvent(ctx->aux_feat, ctx->ev_out_done[slot], 0));  // hbar of this slot is free again
    }
    char* sam_c = sam ? reinterpret_cast<char*>(sam) + (i / p2) * 256 * esz : nullptr;
    char* clip_c = clipseg ? reinterpret_cast<char*>(clipseg) + i * 192 * esz : nullptr;
    int rc = render_chunk(ctx, origins + 3 * i, dirs + 3 * i, nears ? nears + i : nullptr, fars ? fars + i : nullptr, n,
                          flags, opts, rgb + 3 * i, depth + i, acc ? acc + i : nullptr,
                          prop_depth ? prop_depth + i : nullptr, sam_c, clip_c, nullptr, cs);
    if (rc) return rc;
    if (pipelined) {
      CK(cudaEventRecord(ctx->ev_feat_done[slot], ctx->aux_feat));
      CK(cudaEventRecord(ctx->ev_out_done[slot], ctx->aux_out));
    }
    // exchange: as soon as this chunk's feature rows exist, move them into every destination's frame buffer
    if (move_sam && sam && (flags & SNRF_WANT_SAM) && !(flags & SNRF_PATCH)) {
      const Replication& R = ctx->rep[0];
      const char* lo = reinterpret_cast<const char*>(R.local);
      const size_t bytes = static_cast<size_t>(n) * 256 * esz;
      if (sam_c >= lo && sam_c + bytes <= lo + R.bytes) {
        const int64_t off = sam_c - lo;
        cudaStream_t so = pipelined ? ctx->aux_out : s;
        CK(cudaEventRecord(ctx->ev_out[c & 1], so));
        if (ctx->rep_mode == 2) {
          cudaStream_t cp = ctx->copy_stream[0];
          used |= 1u;
          CK(cudaStreamWaitEvent(cp, ctx->ev_out[c & 1], 0));
          if (R.mc) {  // one multicast store reaches every rank
            LAUNCH(launch_push_rows_mc(sam_c, reinterpret_cast<char*>(R.mc) + off, bytes, cp));
          } else {
            void* dst[kMaxPeers] = {nullptr};
            for (int q = 0; q < R.n; ++q) dst[q] = reinterpret_cast<char*>(R.peer[q]) + off;
            LAUNCH(launch_push_rows(sam_c, dst, R.n, bytes, cp));
          }
        } else {
          const int pieces = ctx->dma_split;
          const size_t piece = ((bytes / pieces) + 255) & ~static_cast<size_t>(255);
          for (int q = 0; q < R.n; ++q)
            for (int k = 0; k < pieces; ++k) {
              const size_t b0 = k * piece;
              if (b0 >= bytes) break;
              const size_t nb = bytes - b0 < piece ? bytes - b0 : piece;
              const int si = (q * pieces + k) % kCopyStreams;
              cudaStream_t cp = ctx->copy_stream[si];
              used |= 1u << si;
              CK(cudaStreamWaitEvent(cp, ctx->ev_out[c & 1], 0));
              CK(cudaMemcpyAsync(reinterpret_cast<char*>(R.peer[q]) + off + b0, sam_c + b0, nb, cudaMemcpyDefault, cp));
            }
        }
      }
    }
  }
  if (moved) {
    // the small per-ray outputs (24 B/ray) go once per frame, behind the last march kernel
    float* outs[4] = {rgb, depth, acc, prop_depth};
    const int64_t widths[4] = {3, 1, 1, 1};
    if (n_rays == 0) CK(cudaEventRecord(ctx->ev_join, s));  // otherwise recorded behind the last march (see above)
    for (int w = 0; w < 4; ++w) {
      const Replication& R = ctx->rep[w + 1];
      if (!outs[w] || !R.local) continue;
      const char* lo = reinterpret_cast<const char*>(R.local);
      const char* p = reinterpret_cast<const char*>(outs[w]);
      if (p < lo || p + n_rays * widths[w] * sizeof(float) > lo + R.bytes) continue;
      const int64_t off = (p - lo) / static_cast<int64_t>(sizeof(float));
      for (int q = 0; q < R.n; ++q) {
        const int si = (kCopyStreams - 1 - q) % kCopyStreams;
        cudaStream_t cp = ctx->copy_stream[si];
        used |= 1u << si;
        CK(cudaStreamWaitEvent(cp, ctx->ev_join, 0));
        CK(cudaMemcpyAsync(R.peer[q] + off, outs[w], n_rays * widths[w] * sizeof(float), cudaMemcpyDefault, cp));
      }
    }
    for (int q = 0; q < kCopyStreams; ++q) {  // the caller's stream resumes after every copy has been queued and finished
      if (!((used >> q) & 1u)) continue;
      CK(cudaEventRecord(ctx->ev_copy[q], ctx->copy_stream[q]));
      CK(cudaStreamWaitEvent(s, ctx->ev_copy[q], 0));
    }
  }
  if (pipelined) {  // join: the caller's stream continues only when every chunk's outputs are complete
    CK(cudaEventRecord(ctx->ev_join, ctx->aux_feat));
    CK(cudaStreamWaitEvent(s, ctx->ev_join, 0));
    CK(cudaEventRecord(ctx->ev_join2, ctx->aux_out));
    CK(cudaStreamWaitEvent(s, ctx->ev_join2, 0));
  }
  return SNRF_OK;
}

// ---------------------------------------------------------------------------------------------------
// camera ray generation
// ---------------------------------------------------------------------------------------------------
static int fill_raygen(snrf_ctx* ctx, RayGenParams& R, const snrf_camera* cam, const int32_t* rows_host, int n_rows,
                       const int32_t* cols_host, int n_cols, int patch, cudaStream_t s) {
  if (!cam) return fail(ctx, SNRF_E_INVALID, "null camera");
  if (cam->camera_type < SNRF_CAM_PERSPECTIVE || cam->camera_type > SNRF_CAM_EQUIRECTANGULAR)
    return fail(ctx, SNRF_E_INVALID, "Camera type %d not supported.", cam->camera_type);
  if (cam->width <= 0 || cam->height <= 0 || n_rows < 0 || n_cols < 0)
    return fail(ctx, SNRF_E_INVALID, "bad image / grid size");
  if (!rows_host && n_rows != cam->height) return fail(ctx, SNRF_E_INVALID, "rows NULL needs n_rows == height");
  if (!cols_host && n_cols != cam->width) return fail(ctx, SNRF_E_INVALID, "cols NULL needs n_cols == width");
  if (patch > 1 && (n_rows % patch != 0 || n_cols % patch != 0))
    return fail(ctx, SNRF_E_INVALID, "a %d x %d pixel grid does not divide into %d x %d patches", n_rows, n_cols, patch, patch);
  for (int i = 0; rows_host && i < n_rows; ++i)
    if (rows_host[i] < 0 || rows_host[i] >= cam->height) return fail(ctx, SNRF_E_INVALID, "row index %d outside the image", rows_host[i]);
  for (int i = 0; cols_host && i < n_cols; ++i)
    if (cols_host[i] < 0 || cols_host[i] >= cam->width) return fail(ctx, SNRF_E_INVALID, "column index %d outside the image", cols_host[i]);
  memset(&R, 0, sizeof(R));
  R.cam.fx = cam->fx; R.cam.fy = cam->fy; R.cam.cx = cam->cx; R.cam.cy = cam->cy;
  R.cam.type = cam->camera_type;
  R.cam.has_dist = cam->has_distortion != 0;
  memcpy(R.cam.dist, cam->distortion, sizeof(R.cam.dist));
  memcpy(R.cam.c2w, cam->c2w, sizeof(R.cam.c2w));
  R.has_aabb = cam->has_aabb != 0;
  memcpy(R.aabb, cam->aabb, sizeof(R.aabb));
  R.n_rows = n_rows;
  R.n_cols = n_cols;
  R.patch = patch > 1 ? patch : 1;
  if (rows_host && n_rows > 0) {
    CK(ctx->cam_rows.ensure(n_rows * sizeof(int32_t)));
    CK(cudaMemcpyAsync(ctx->cam_rows.p, rows_host, n_rows * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    R.rows = ctx->cam_rows.as<int>();
  }
  if (cols_host && n_cols > 0) {
    CK(ctx->cam_cols.ensure(n_cols * sizeof(int32_t)));
    CK(cudaMemcpyAsync(ctx->cam_cols.p, cols_host, n_cols * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    R.cols = ctx->cam_cols.as<int>();
  }
  return SNRF_OK;
}

int snrf_generate_rays(snrf_ctx* ctx, const snrf_camera* cam, const int32_t* rows_host, int n_rows,
                       const int32_t* cols_host, int n_cols, int patch, float* origins, float* dirs,
                       float* pixel_area, float* nears, float* fars, void* stream) {
  if (!ctx) return SNRF_E_INVALID;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  CK(cudaSetDevice(ctx->device));
  RayGenParams R;
  int rc = fill_raygen(ctx, R, cam, rows_host, n_rows, cols_host, n_cols, patch, s);
  if (rc) return rc;
  if (static_cast<int64_t>(n_rows) * n_cols == 0) return SNRF_OK;
  if (!origins || !dirs) return fail(ctx, SNRF_E_INVALID, "null output");
  R.origins = origins;
  R.dirs = dirs;
  R.pixel_area = pixel_area;
  if (R.has_aabb && (!nears || !fars)) return fail(ctx, SNRF_E_INVALID, "a crop box needs nears / fars outputs");
  R.nears = nears;
  R.fars = fars;
  LAUNCH(launch_raygen(R, s));
  return SNRF_OK;
}

int snrf_render_camera(snrf_ctx* ctx, const snrf_camera* cam, const int32_t* rows_host, int n_rows,
                       const int32_t* cols_host, int n_cols, int64_t chunk, uint32_t flags,
                       const snrf_render_opts* opts, float* rgb, float* depth, float* acc, float* prop_depth,
                       float* sam, float* clipseg, void* stream) {
  if (!ctx) return SNRF_E_INVALID;
  if (!opts) return fail(ctx, SNRF_E_INVALID, "null argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  CK(cudaSetDevice(ctx->device));
  const int patch = (flags & SNRF_PATCH) ? opts->patch_size : 1;
  RayGenParams R;
  int rc = fill_raygen(ctx, R, cam, rows_host, n_rows, cols_host, n_cols, patch, s);
  if (rc) return rc;
  const int64_t n = static_cast<int64_t>(n_rows) * n_cols;
  if (n == 0) return SNRF_OK;
  CK(ctx->cam_o.ensure(n * 3 * sizeof(float)));
  CK(ctx->cam_d.ensure(n * 3 * sizeof(float)));
  R.origins = ctx->cam_o.as<float>();
  R.dirs = ctx->cam_d.as<float>();
  if (R.has_aabb) {
    CK(ctx->cam_near.ensure(n * sizeof(float)));
    CK(ctx->cam_far.ensure(n * sizeof(float)));
    R.nears = ctx->cam_near.as<float>();
    R.fars = ctx->cam_far.as<float>();
  }
  LAUNCH(launch_raygen(R, s));
  return snrf_render_frame(ctx, R.origins, R.dirs, R.nears, R.fars, n, chunk, flags, opts, rgb, depth, acc, prop_depth,
                           sam, clipseg, stream);
}

// ---------------------------------------------------------------------------------------------------
// training side of the feature-field branch
// ---------------------------------------------------------------------------------------------------
int snrf_feature_forward(snrf_ctx* ctx, int which, const float* origins, const float* dirs, const float* sam_t,
                         const float* sam_w, int64_t n_rays, float* out, void* enc_f16, void* stream) {
  if (!ctx) return SNRF_E_INVALID;
  if (which < 0 || which > 1 || n_rays < 0) return fail(ctx, SNRF_E_INVALID, "bad argument");
  if (n_rays == 0) return SNRF_OK;
  if (!origins || !dirs || !sam_t || !sam_w || !out) return fail(ctx, SNRF_E_INVALID, "null argument");
  FeatureNet& f = ctx->feat[which];
  if (!f.have_grid[0] || !f.have_grid[1] || !f.have_net)
    return fail(ctx, SNRF_E_STATE, "%s parameters not uploaded", which == 0 ? "sam_field" : "clipseg");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  CK(cudaSetDevice(ctx->device));
  DevBuf& hbar = ctx->hbar[0][which];
  CK(hbar.ensure(n_rays * 256 * 2));
  SamParams S;
  memset(&S, 0, sizeof(S));
  S.origins = origins;
  S.dirs = dirs;
  S.sam_t = sam_t;
  S.sam_w = sam_w;
  S.n_rays = n_rays;
  S.enc[0] = f.grid[0];
  S.enc[1] = f.grid[1];
  S.w1 = f.w1_core.as<__half>();
  S.hbar = hbar.as<__half>();
  S.dbg_feat = reinterpret_cast<__half*>(enc_f16);
  TIMED_LAUNCH(1, s, launch_sam(S, ctx->engine == 1, false, ctx->sm_count, s));
  GemmParams G;
  memset(&G, 0, sizeof(G));
  G.a = hbar.as<__half>();
  G.w = f.w2_core.as<__half>();
  G.m = n_rays;
  G.n = f.n_out;
  G.taps = 1;
  G.out_f32 = out;
  G.out_mode = 0;
  TIMED_LAUNCH(2, s, launch_tapgemm(G, ctx->engine == 1, ctx->sm_count, s));
  return SNRF_OK;
}

int snrf_feature_backward(snrf_ctx* ctx, int which, const float* origins, const float* dirs, const float* sam_t,
                          const float* sam_w, int64_t n_rays, const float* d_out, const void* enc_f16,
                          float* grad_net, float* grad_grid0, float* grad_grid1, void* stream) {
  if (!ctx) return SNRF_E_INVALID;
  if (which < 0 || which > 1 || n_rays < 0) return fail(ctx, SNRF_E_INVALID, "bad argument");
  if (n_rays == 0) return SNRF_OK;
  if (!origins || !dirs || !sam_t || !sam_w || !d_out || !enc_f16) return fail(ctx, SNRF_E_INVALID, "null argument");
  if ((reinterpret_cast<uintptr_t>(grad_grid0) | reinterpret_cast<uintptr_t>(grad_grid1)) & 15u)
    return fail(ctx, SNRF_E_INVALID, "grid gradients must be 16-byte aligned");
  FeatureNet& f = ctx->feat[which];
  if (!f.have_grid[0] || !f.have_grid[1] || !f.have_net)
    return fail(ctx, SNRF_E_STATE, "%s parameters not uploaded", which == 0 ? "sam_field" : "clipseg");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  CK(cudaSetDevice(ctx->device));
  CK(ctx->bwd_scratch.ensure(feat_bwd_scratch_floats(n_rays) * sizeof(float)));
  FeatBwdParams B;
  memset(&B, 0, sizeof(B));
  B.origins = origins;
  B.dirs = dirs;
  B.sam_t = sam_t;
  B.sam_w = sam_w;
  B.d_out = d_out;
  B.x = reinterpret_cast<const __half*>(enc_f16);
  B.w1 = f.w1_rm.as<__half>();
  B.w2 = f.w2_rm.as<__half>();
  B.n_rays = n_rays;
  B.n_out = f.n_out;
  B.enc[0] = f.grid[0];
  B.enc[1] = f.grid[1];
  B.d_hbar = ctx->bwd_scratch.as<float>();
  B.cutoff = ctx->feat_cutoff;  // same significance rule as the forward pass (snrf_set_feature_cutoff); < 0 = all 16 slots
  // a frozen parameter's gradient goes to a sink so that the kernels stay branch-free
  const size_t n_net = 256 * 192 + static_cast<size_t>(f.n_out) * 256;
  const size_t n_g0 = static_cast<size_t>(f.grid[0].lv[11].offset + f.grid[0].lv[11].size) * 8;
  const size_t n_g1 = static_cast<size_t>(f.grid[1].lv[11].offset + f.grid[1].lv[11].size) * 8;
  size_t sink = 0;
  if (!grad_net) sink = n_net;
  if (!grad_grid0 && n_g0 > sink) sink = n_g0;
  if (!grad_grid1 && n_g1 > sink) sink = n_g1;
  if (sink) CK(ctx->bwd_sink.ensure(sink * sizeof(float)));
  float* sinkp = ctx->bwd_sink.as<float>();
  B.g_w1 = grad_net ? grad_net : sinkp;
  B.g_w2 = grad_net ? grad_net + 256 * 192 : sinkp;
  B.g_table[0] = grad_grid0 ? grad_grid0 : sinkp;
  B.g_table[1] = grad_grid1 ? grad_grid1 : sinkp;
  int64_t launched = 0;
  const cudaError_t e = launch_feat_backward(B, s, &launched);
  ctx->launches += launched;
  if (e != cudaSuccess) return fail(ctx, SNRF_E_CUDA, "feature backward failed: %s", cudaGetErrorString(e));
  return SNRF_OK;
}

int snrf_patch_aggregate_backward(snrf_ctx* ctx, const float* feat_in, int64_t n_patches, int p, const float* d_out,
                                  float* grad_w0, float* grad_b0, float* grad_w2, float* grad_b2, float* d_feat,
                                  void* stream) {
  if (!ctx) return SNRF_E_INVALID;
  if (!feat_in || !d_out || !grad_w0 || !grad_b0 || !grad_w2 || !grad_b2 || n_patches < 0)
    return fail(ctx, SNRF_E_INVALID, "null argument");
  if (p != 4) return fail(ctx, SNRF_E_INVALID, "patch_size %d unsupported (the conv head kernels are built for p = 4)", p);
  if (!ctx->have_conv) return fail(ctx, SNRF_E_STATE, "conv head parameters not uploaded");
  if (n_patches == 0) return SNRF_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  CK(cudaSetDevice(ctx->device));
  CK(ctx->bwd_scratch.ensure(conv_bwd_scratch_bytes(n_patches)));
  ConvBwdParams B;
  memset(&B, 0, sizeof(B));
  B.feat_in = feat_in;
  B.d_out = d_out;
  B.rows = n_patches * 16;
  B.w1 = ctx->conv_w_rm[0].as<__half>();
  B.w2 = ctx->conv_w_rm[1].as<__half>();
  B.b1 = ctx->conv_b[0].as<float>();
  B.b2 = ctx->conv_b[1].as<float>();
  B.g_w1 = grad_w0; B.g_b1 = grad_b0; B.g_w2 = grad_w2; B.g_b2 = grad_b2;
  B.d_feat = d_feat;
  int64_t launched = 0;
  const cudaError_t e = launch_conv_backward(B, ctx->bwd_scratch.p, s, &launched);
  ctx->launches += launched;
  if (e != cudaSuccess) return fail(ctx, SNRF_E_CUDA, "conv head backward failed: %s", cudaGetErrorString(e));
  return SNRF_OK;
}

int snrf_field_backward(snrf_ctx* ctx, int which, const float* xyz, const float* dirs, int64_t n,
                        const float* d_density, const float* d_rgb, float* grad_base, float* grad_head, void* stream) {
  if (!ctx) return SNRF_E_INVALID;
  if (which < 0 || which > 1 || n < 0) return fail(ctx, SNRF_E_INVALID, "bad argument");
  if (n == 0) return SNRF_OK;
  if (!xyz || !grad_base || (!d_density && !d_rgb)) return fail(ctx, SNRF_E_INVALID, "null argument");
  if (d_rgb && (which != 1 || !dirs || !grad_head))
    return fail(ctx, SNRF_E_INVALID, "d_rgb needs the nerfacto field (which = 1), dirs and grad_head");
  if (reinterpret_cast<uintptr_t>(grad_base) & 15u) return fail(ctx, SNRF_E_INVALID, "grad_base must be 16-byte aligned");
  if (which == 0 ? !ctx->have_prop : !(ctx->have_base && (ctx->have_head || !d_rgb)))
    return fail(ctx, SNRF_E_STATE, "field parameters not uploaded");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  CK(cudaSetDevice(ctx->device));
  CK(ctx->bwd_scratch.ensure(field_bwd_scratch_bytes(n)));
  FieldBwdParams B;
  memset(&B, 0, sizeof(B));
  B.xyz = xyz;
  B.dirs = dirs;
  B.d_density = d_density;
  B.d_rgb = d_rgb;
  B.n = n;
  B.which = which;
  B.grid = which == 0 ? ctx->prop_grid : ctx->field_grid;
  B.w1 = which == 0 ? ctx->prop_w1_rm.as<__half>() : ctx->base_w1_rm.as<__half>();
  B.w2 = which == 0 ? ctx->prop_w2_rm.as<__half>() : ctx->base_w2_rm.as<__half>();
  B.wh1 = ctx->head_w1_rm.as<__half>();
  B.wh2 = ctx->head_w2_rm.as<__half>();
  B.wh3 = ctx->head_w3_rm.as<__half>();
  B.g_base = grad_base;
  B.g_head = grad_head;
  int64_t launched = 0;
  const cudaError_t e = launch_field_backward(B, ctx->bwd_scratch.p, s, &launched);
  ctx->launches += launched;
  if (e != cudaSuccess) return fail(ctx, SNRF_E_CUDA, "field backward failed: %s", cudaGetErrorString(e));
  return SNRF_OK;
}

int snrf_pick_samples(snrf_ctx* ctx, const float* weights, const float* starts, const float* ends, int64_t n, int S,
                      int k, float sharpen, float* sam_t, float* sam_w, void* stream) {
  if (!ctx) return SNRF_E_INVALID;
  if (!weights || !starts || !ends || !sam_t || !sam_w || n < 0 || S < 1 || k < 1 || k > S)
    return fail(ctx, SNRF_E_INVALID, "bad argument");
  if (n == 0) return SNRF_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  CK(cudaSetDevice(ctx->device));
  LAUNCH(launch_pick_samples(weights, starts, ends, n, S, k, sharpen, sam_t, sam_w, s));
  return SNRF_OK;
}

int snrf_ray_op_backward(snrf_ctx* ctx, int mode, const float* a, const float* b, const float* g, float* out_a,
                         float* out_b, int64_t n, int S, int bg_mode, const float* bg_host, void* stream) {
  if (!ctx) return SNRF_E_INVALID;
  if ((mode != 0 && mode != 3) || !a || !b || !g || !out_a || S <= 0 || n < 0) return fail(ctx, SNRF_E_INVALID, "bad argument");
  if (mode == 0 && S > kMaxRaySamples) return fail(ctx, SNRF_E_INVALID, "get_weights backward supports up to %d samples per ray", kMaxRaySamples);
  if (mode == 3 && !out_b) return fail(ctx, SNRF_E_INVALID, "missing output for dL/dweights");
  if (n == 0) return SNRF_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  CK(cudaSetDevice(ctx->device));
  if (mode == 0)
    LAUNCH(launch_weights_bwd(a, b, g, out_a, n, S, s));
  else
    LAUNCH(launch_rgb_bwd(a, b, g, bg_mode == SNRF_BG_FIXED ? 1 : 0, bg_host, out_a, out_b, n, S, s));
  return SNRF_OK;
}

int snrf_sample(snrf_ctx* ctx, const float* origins, const float* dirs, const float* nears, const float* fars,
                int64_t n_rays, const snrf_render_opts* opts, float* prop_weights, float* edges, float* prop_depth,
                void* stream) {
  if (!ctx) return SNRF_E_INVALID;
  if (n_rays <= 0) return n_rays == 0 ? SNRF_OK : fail(ctx, SNRF_E_INVALID, "negative ray count");
  if (!origins || !dirs || !opts) return fail(ctx, SNRF_E_INVALID, "null argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  CK(cudaSetDevice(ctx->device));
  MarchParams M;
  int rc = fill_march(ctx, M, origins, dirs, nears, fars, n_rays, opts);
  if (rc) return rc;
  M.flags = kFlagSamplesOnly;
  M.dbg_w0 = prop_weights;
  M.dbg_edges = edges;
  M.prop_depth = prop_depth;
  LAUNCH(ctx->march_v1 ? launch_march_v1(M, ctx->sm_count, s) : launch_march(M, ctx->sm_count, s));
  return SNRF_OK;
}

// ---------------------------------------------------------------------------------------------------
// component-level queries
// ---------------------------------------------------------------------------------------------------
int snrf_query_density(snrf_ctx* ctx, int which, const float* xyz, int64_t n, float* density, void* geo_f16,
                       void* stream) {
  if (!ctx) return SNRF_E_INVALID;
  if (!xyz || !density || which < 0 || which > 1) return fail(ctx, SNRF_E_INVALID, "bad argument");
  if (n == 0) return SNRF_OK;
  if (which == 0 ? !ctx->have_prop : !ctx->have_base) return fail(ctx, SNRF_E_STATE, "field parameters not uploaded");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  CK(cudaSetDevice(ctx->device));
  const int width = which == 0 ? 10 : 32, hidden = which == 0 ? 16 : 64, k_w = which == 0 ? 16 : 32;
  CK(ctx->q_feat.ensure(n * width * 2));
  CK(ctx->q_h1.ensure(n * hidden * 2));
  CK(ctx->q_h2.ensure(n * 16 * 2));
  CK(ctx->q_sel.ensure(n * 4));
  QueryParams Q;
  memset(&Q, 0, sizeof(Q));
  Q.xyz = xyz;
  Q.n = n;
  Q.grid[0] = which == 0 ? ctx->prop_grid : ctx->field_grid;
  Q.n_grids = 1;
  Q.linf = 1;
  Q.selector = 1;
  Q.feat = ctx->q_feat.as<__half>();
  Q.sel = ctx->q_sel.as<float>();
  LAUNCH(launch_encode(Q, s));
  const __half* w1 = which == 0 ? ctx->prop_w1_rm.as<__half>() : ctx->base_w1_rm.as<__half>();
  const __half* w2 = which == 0 ? ctx->prop_w2_rm.as<__half>() : ctx->base_w2_rm.as<__half>();
  LAUNCH(launch_dense(Q.feat, width, width, w1, k_w, 0.f, ctx->q_h1.as<__half>(), hidden, hidden, 1, n, s));
  LAUNCH(launch_dense(ctx->q_h1.as<__half>(), hidden, hidden, w2, hidden, 0.f, ctx->q_h2.as<__half>(), 16, 16, 0, n, s));
  LAUNCH(launch_density_finish(ctx->q_h2.as<__half>(), 16, Q.sel, density,
                               which == 1 ? reinterpret_cast<__half*>(geo_f16) : nullptr, 15, n, s));
  return SNRF_OK;
}

int snrf_query_rgb(snrf_ctx* ctx, const float* dirs, const void* geo_f16, int64_t n, float* rgb, void* stream) {
  if (!ctx) return SNRF_E_INVALID;
  if (!dirs || !geo_f16 || !rgb) return fail(ctx, SNRF_E_INVALID, "null argument");
  if (n == 0) return SNRF_OK;
  if (!ctx->have_head) return fail(ctx, SNRF_E_STATE, "field head parameters not uploaded");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  CK(cudaSetDevice(ctx->device));
  CK(ctx->q_x.ensure(n * 32 * 2));
  CK(ctx->q_h1.ensure(n * 64 * 2));
  CK(ctx->q_h2.ensure(n * 64 * 2));
  CK(ctx->q_feat.ensure(n * 16 * 2));
  LAUNCH(launch_head_input(dirs, reinterpret_cast<const __half*>(geo_f16), ctx->q_x.as<__half>(), n, s));
  LAUNCH(launch_dense(ctx->q_x.as<__half>(), 32, 32, ctx->head_w1_rm.as<__half>(), 32, 1.f, ctx->q_h1.as<__half>(), 64, 64, 1, n, s));
  LAUNCH(launch_dense(ctx->q_h1.as<__half>(), 64, 64, ctx->head_w2_rm.as<__half>(), 64, 0.f, ctx->q_h2.as<__half>(), 64, 64, 1, n, s));
  LAUNCH(launch_dense(ctx->q_h2.as<__half>(), 64, 64, ctx->head_w3_rm.as<__half>(), 64, 0.f, ctx->q_feat.as<__half>(), 16, 16, 2, n, s));
  LAUNCH(launch_half_to_float(ctx->q_feat.as<__half>(), 16, rgb, 3, 3, n, s));
  return SNRF_OK;
}

int snrf_query_features(snrf_ctx* ctx, int which, const float* xyz, int64_t n, void* hashgrid_f16, float* out,
                        void* stream) {
  if (!ctx) return SNRF_E_INVALID;
  if (!xyz || !out || which < 0 || which > 1) return fail(ctx, SNRF_E_INVALID, "bad argument");
  if (n == 0) return SNRF_OK;
  FeatureNet& f = ctx->feat[which];
  if (!f.have_grid[0] || !f.have_grid[1] || !f.have_net) return fail(ctx, SNRF_E_STATE, "feature field parameters not uploaded");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  CK(cudaSetDevice(ctx->device));
  __half* feat = reinterpret_cast<__half*>(hashgrid_f16);
  if (!feat) {
    CK(ctx->q_feat.ensure(n * 192 * 2));
    feat = ctx->q_feat.as<__half>();
  }
  CK(ctx->q_h1.ensure(n * 256 * 2));
  CK(ctx->q_h2.ensure(n * 256 * 2));
  QueryParams Q;
  memset(&Q, 0, sizeof(Q));
  Q.xyz = xyz;
  Q.n = n;
  Q.grid[0] = f.grid[0];
  Q.grid[1] = f.grid[1];
  Q.n_grids = 2;
  Q.linf = 0;
  Q.selector = 0;
  Q.feat = feat;
  LAUNCH(launch_encode(Q, s));
  LAUNCH(launch_dense(feat, 192, 192, f.w1_rm.as<__half>(), 192, 0.f, ctx->q_h1.as<__half>(), 256, 256, 1, n, s));
  LAUNCH(launch_dense(ctx->q_h1.as<__half>(), 256, 256, f.w2_rm.as<__half>(), 256, 0.f, ctx->q_h2.as<__half>(), f.n_out, f.n_out, 0, n, s));
  LAUNCH(launch_half_to_float(ctx->q_h2.as<__half>(), f.n_out, out, f.n_out, f.n_out, n, s));
  return SNRF_OK;
}

int snrf_ray_op(snrf_ctx* ctx, int mode, const float* a, const float* b, const float* c, float* out, int64_t n, int S,
                int C, int bg_mode, const float* bg_host, void* stream) {
  if (!ctx) return SNRF_E_INVALID;
  if (mode < 0 || mode > 4 || !a || !out || S <= 0) return fail(ctx, SNRF_E_INVALID, "bad argument");
  if ((mode == 0 || mode == 2 || mode == 3 || mode == 4) && !b) return fail(ctx, SNRF_E_INVALID, "missing operand b");
  if (mode == 2 && !c) return fail(ctx, SNRF_E_INVALID, "missing operand c");
  if (n == 0) return SNRF_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  CK(cudaSetDevice(ctx->device));
  LAUNCH(launch_ray_ops(mode, a, b, c, out, n, S, C, bg_mode == SNRF_BG_FIXED ? kBgFixed : kBgLastSample, bg_host, s));
  return SNRF_OK;
}

}  // extern "C"
