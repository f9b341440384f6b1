// Kernel parameter blocks and launch entry points (internal to libsnrf).
#pragma once
#include "common.cuh"

namespace snrf {

// ---- kernel A: per-ray march -----------------------------------------------------------------
// packed mma.sync B-fragment tiles (32 lanes x uint2 = 256 B each), in this order:
constexpr int kFragBase1 = 0;    // base MLP  32 -> 64 : 8 n-tiles x 2 k-steps
constexpr int kFragBase2 = 16;   // base MLP  64 -> 16 : 2 x 4
constexpr int kFragHead1 = 24;   // head MLP  32 -> 64 : 8 x 2   (input columns permuted: [pad, geo(15), SH(16)])
constexpr int kFragHead2 = 40;   // head MLP  64 -> 64 : 8 x 4
constexpr int kFragHead3 = 72;   // head MLP  64 -> 8  : 1 x 4   (only rgb = outputs 0..2 are used)
constexpr int kFragProp1 = 76;   // proposal MLP 16 -> 16 : 2 x 1   (input columns 10..15 are tcnn's zero padding)
constexpr int kFragProp2 = 78;   // proposal MLP 16 -> 8  : 1 x 1   (only output 0, the density, is used)
constexpr int kMarchFragTiles = 79;
// the five field layers once more in the core-matrix layout (tcgen05 B operands of march_kernel<.., TC>; plan in march.cu)
constexpr uint32_t kMarchCoreBytes = (64 * 32 + 16 * 64 + 64 * 32 + 64 * 64 + 16 * 64) * 2;  // 20480

constexpr uint32_t kFlagSamplesOnly = 1u;  // stop after the PDF resample (backs ProposalNetworkSampler)
constexpr int kBgLastSample = 0;           // RGBRenderer background "last_sample" (renderers.py:102-103)
constexpr int kBgFixed = 1;                // fixed colour (black / white / background_color_override_context)

constexpr int kMaxPeers = 8;
// Aliases of one output buffer in the other ranks' frame buffers (fused tile all-gather): one NVSwitch multicast
// address (multimem.st reaches every rank, this one included) or up to 8 peer-mapped pointers.
struct OutRep {
  float* mc;
  float* peer[kMaxPeers];
  int n;
};
__device__ __forceinline__ void store_rep(float* local, const OutRep& R, int64_t off, float v) {
  if (R.mc) {
    asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;\n" ::"l"(R.mc + off), "f"(v) : "memory");
  } else {
    local[off] = v;
    for (int p = 0; p < R.n; ++p) R.peer[p][off] = v;
  }
}

// "Bricks": a level of an F = 2 grid re-laid-out by cell.  Brick c = gx + gy res + gz res^2 holds the 8 corner
// entries of cell (gx,gy,gz) - corner dx + 2 dy + 4 dz, fetched with the level's own index rule (dense with wrap, or
// the hash) - as 32 contiguous, 32-byte aligned bytes: one sector and one LDG.256 per (sample, level) instead of 8
// scattered 4-byte gathers.  Values are copies of table entries, so results are bit-identical; the price is HBM
// capacity (res^3 x 32 B per level: 9.3 GB for levels 0-11 of the nerfacto grid), which a 180 GB part has to spare.
// Built by snrf_set_brick_budget / the upload calls (bricks.cu); lv[l] is valid for l < n (a prefix of the levels).
struct BrickDev {
  const uint4* lv[kMaxLevels];
  int n;
};
cudaError_t launch_brick_build(const GridDev& G, int level, uint4* out, cudaStream_t stream);

struct MarchParams {
  const float* origins;  // [N,3]
  const float* dirs;     // [N,3]
  const float* nears;    // [N] or null -> near_default
  const float* fars;     // [N] or null -> far_default
  int64_t n_rays;
  float near_default, far_default;
  GridDev prop;          // 5 levels x 2
  GridDev field;         // 16 levels x 2
  BrickDev prop_bricks;  // cell-major copies of the leading levels (n = 0: none)
  BrickDev field_bricks;
  const uint2* wfrag;    // kMarchFragTiles x 32 fragment words
  const uint4* wcore;    // kMarchCoreBytes of core-matrix weights, or null: the mma.sync kernel is the only choice
  int use_tc;            // run the field MLPs on tcgen05 (march_kernel<.., TC>) where an instantiation exists
  float* edges;          // [N,33] scratch for the bin edges between the two halves of a split march, or null
  int split;             // run the sampling half and the field half as two launches (needs `edges`)
  const float* pdf_u;    // [33] eval-mode PDF sample positions
  float hist_padding;
  int bg_mode;
  float bg[3];
  int k_sam;
  float sharpen;
  uint32_t flags;
  float* rgb;         // [N,3]
  float* depth;       // [N]
  float* acc;         // [N] or null
  float* prop_depth;  // [N] or null
  OutRep rep[4];      // replication of rgb / depth / acc / prop_depth into the other ranks' frame buffers
  float* sam_t;       // [N,k] 2 x midpoint of the picked samples, or null
  float* sam_w;       // [N,k] sharpened, renormalised weights
  float* dbg_w0;      // [N,64] proposal weights, or null
  float* dbg_edges;   // [N,33] nerf bin edges, or null
  float* dbg_weights; // [N,32]
  float* dbg_density; // [N,32]
  float* dbg_rgb;     // [N,32,3]
  float et_eps;       // early-termination transmittance threshold, 0 = exact (see march_kernel<.., ET>)
  const float* jitter;      // [N,2] training-mode single-jitter draws, or null = eval (see march_kernel<.., JIT>)
  const float* pdf_u_base;  // [33] linspace(0, 1 - 1/33, 33) without the eval-mode half-bin offset
  float anneal;             // proposal-weight annealing exponent (training instantiation only; 1 = off)
};
cudaError_t launch_march(const MarchParams& P, int sm_count, cudaStream_t stream);
// round-1 kernel (march_v1.cu), kept for A/B measurements: SNRF_MARCH=v1
cudaError_t launch_march_v1(const MarchParams& P, int sm_count, cudaStream_t stream);

// ---- kernel B: feature-field gather + first MLP layer + weighted sample reduction ---------------
struct SamParams {
  const float* origins;
  const float* dirs;
  const float* sam_t;  // [N,16]
  const float* sam_w;  // [N,16]
  int64_t n_rays;
  GridDev enc[2];      // 12 levels x 8 each
  const __half* w1;    // [256 x 192] fp16 in core-matrix layout (98304 B)
  __half* hbar;        // [N,256] fp16: sum_k w_k * fp16(relu(W1 x_k))
  __half* dbg_feat;    // [N,16,192] encoder output, or null
};
cudaError_t launch_sam(const SamParams& P, bool tcgen05, bool small_cta, int sm_count, cudaStream_t stream);

// ---- kernel B': kernel B over rays bucketed by their number of significant slots (sam_bucket.cu) ---------
constexpr int kFeatBuckets = 5;  // rays with <= 1, <= 2, <= 4, <= 8, <= 16 significant slots (SLOTS = 1 << bucket)
struct SamBucketParams {
  const float* origins;
  const float* dirs;
  const float* sam_t;  // [N,16]
  const float* sam_w;  // [N,16] descending per ray (march.cu stores slot = rank)
  GridDev enc[2];
  const __half* w1;    // core-matrix layout, as SamParams
  __half* hbar;        // [N,256]
  const int* lists;    // [kFeatBuckets][n_rays] ray indices per bucket
  int* counts;         // [kFeatBuckets] entries per bucket (device memory, written by the pre-pass) + [1] tile counter
  int64_t n_rays;
};
// Pre-pass, one thread per ray: significant slots = 1 + index of the last slot whose weight is not below the
// cut-off (NaN weights - 0/0 rays, sam_model.py:248 - count as significant, so such a ray keeps all 16 slots and
// yields the same NaN row as kernel B); cut-off 0 keeps every non-zero weight.  lists: [kFeatBuckets][n].
// totals (optional): running number of rays per bucket over all launches (snrf_feature_slot_stats).
SNRF_HD void bucket_assign_one(const float* sam_w, float eps, int* counts, int* lists, int64_t n, int64_t ray,
                               unsigned long long* totals = nullptr) {
  int k = 0;
  for (int s = 0; s < 16; ++s) {
    const float w = sam_w[ray * 16 + s];
    if (!(w < eps) && w != 0.f) k = s + 1;
  }
  const int b = k <= 1 ? 0 : k <= 2 ? 1 : k <= 4 ? 2 : k <= 8 ? 3 : 4;
#ifdef __CUDA_ARCH__
  const int pos = atomicAdd(counts + b, 1);
  if (totals) atomicAdd(totals + b, 1ull);
#else
  const int pos = counts[b]++;
  if (totals) totals[b] += 1ull;
#endif
  lists[static_cast<int64_t>(b) * n + pos] = static_cast<int>(ray);
}
cudaError_t launch_bucket_assign(const float* sam_w, float eps, int* counts, int* lists, int64_t n,
                                 unsigned long long* totals, cudaStream_t stream);
cudaError_t launch_sam_bucketed(const SamBucketParams& P, int sm_count, cudaStream_t stream);

// ---- kernel C/D: tap GEMM  out = act(sum_t A_t[M,256] x W_t[N,256]^T + bias) ----------------------
struct GemmParams {
  const __half* a;     // [M,256] fp16 rows
  const __half* w;     // taps x [N x 256] fp16 in core-matrix layout
  const float* bias;   // [N] or null
  float* out_f32;      // out_mode 0: [M,N];  out_mode 2: [M/16,N] mean over each 16-row patch
  __half* out_f16;     // out_mode 1: [M,N]
  int64_t m;
  int n;               // 256 or 192
  int taps;            // 1 (plain GEMM) or 9 (3x3 conv over 4x4 patches of 16 consecutive rows)
  int relu;
  int out_mode;
  // Fused tile all-gather (out_mode 0 only): besides out_f32, every output row is stored into the frame buffers of
  // the other ranks at the same byte offset - either through one NVSwitch multicast address (multimem.st) or through
  // n_peers peer-mapped pointers (plain st.global over NVLink).  All null / 0 = single-GPU behaviour.
  float* out_mc;        // multicast alias of out_f32 (covers every rank, including this one), or null
  float* out_peer[8];   // peer aliases of out_f32 on the other ranks
  int n_peers;
};
cudaError_t launch_tapgemm(const GemmParams& P, bool tcgen05, int sm_count, cudaStream_t stream);
// TMA-fed, warp-specialised version for the plain output layer (taps = 1, fp32 rows, no fused replication): gemm_tma.cu.
// cudaErrorNotSupported = this launch is not its case; launch_tapgemm falls through to the kernel above.
cudaError_t launch_tapgemm_tma(const GemmParams& P, int sm_count, cudaStream_t stream);

// ---- tile exchange by peer stores from a small kernel (exchange.cu) --------------------------------
cudaError_t launch_push_rows(const void* src, void* const* dst, int n_dst, size_t bytes, cudaStream_t stream);
cudaError_t launch_push_rows_mc(const void* src, void* mc_dst, size_t bytes, cudaStream_t stream);

// ---- stand-alone field queries (back Field.density_fn / SAMField.get_outputs) --------------------
struct QueryParams {
  const float* xyz;  // [N,3] world positions
  int64_t n;
  GridDev grid[2];
  int n_grids;
  int linf;          // contraction order
  int selector;
  __half* feat;      // [N, sum(L*F)] encoder output (fp16)
  float* sel;        // [N] selector or null
};
cudaError_t launch_encode(const QueryParams& P, cudaStream_t stream);
cudaError_t launch_dense(const __half* in, int ld_in, int k_in, const __half* w, int k_w, float pad_value, __half* out,
                         int ld_out, int n_out, int act, int64_t n, cudaStream_t stream);
cudaError_t launch_density_finish(const __half* h, int ld, const float* sel, float* density, __half* geo, int n_geo,
                                  int64_t n, cudaStream_t stream);
cudaError_t launch_head_input(const float* dirs, const __half* geo, __half* x, int64_t n, cudaStream_t stream);
cudaError_t launch_half_to_float(const __half* in, int ld_in, float* out, int ld_out, int cols, int64_t n,
                                 cudaStream_t stream);
cudaError_t launch_ray_ops(int mode, const float* a, const float* b, const float* c, float* out, int64_t n, int S, int C,
                           int bg_mode, const float* bg, cudaStream_t stream);
cudaError_t launch_f32_to_f16(const float* in, __half* out, int64_t n, cudaStream_t stream);
cudaError_t launch_pack_core(const float* w, __half* out, int rows, int cols, cudaStream_t stream, const int* perm = nullptr);
cudaError_t launch_pack_frag(const float* w, int k_w, const int* perm, uint2* out, int n_tiles, int k_steps,
                             cudaStream_t stream);
cudaError_t launch_pack_prop(const float* w1, int k_w, const float* w2, float* o1, float* o2, cudaStream_t stream);

}  // namespace snrf
