// TEST INFRASTRUCTURE ONLY - never loaded by the package.
//
// Host emulation of libsnrf's simple kernels (one thread = one item, no shared memory, no warp intrinsics): the very
// same __host__ __device__ bodies the CUDA kernels call are run in a plain loop on the CPU, so that their arithmetic,
// indexing and layouts can be checked against the oracle in a container without a GPU.  It says nothing about launch
// configuration, races between threads or device intrinsics - the `-m gpu` parity tests remain the real gate.
// Built by tests/emu/build_emu.py with nvcc (host code only; no CUDA call is ever made).
#include "../../segment-anything-in-nerf_b200/csrc/raygen.cuh"
#include "../../segment-anything-in-nerf_b200/csrc/backward.cuh"
#include "../../segment-anything-in-nerf_b200/csrc/kernels.cuh"

#include <string.h>

#include <vector>

using namespace snrf;

extern "C" {

// mirrors snrf_generate_rays: all pointers are HOST pointers here
int emu_generate_rays(const float* intr /*fx fy cx cy*/, int type, int has_dist, const float* dist, const float* c2w,
                      const int* rows, int n_rows, const int* cols, int n_cols, int patch, float* origins, float* dirs,
                      float* pixel_area, const float* aabb /*6 or null*/, float* nears, float* fars) {
  RayGenParams P;
  P.has_aabb = aabb != nullptr;
  for (int i = 0; i < 6; ++i) P.aabb[i] = aabb ? aabb[i] : 0.f;
  P.nears = nears; P.fars = fars;
  P.cam.fx = intr[0]; P.cam.fy = intr[1]; P.cam.cx = intr[2]; P.cam.cy = intr[3];
  P.cam.type = type;
  P.cam.has_dist = has_dist;
  for (int i = 0; i < 6; ++i) P.cam.dist[i] = has_dist ? dist[i] : 0.f;
  for (int i = 0; i < 12; ++i) P.cam.c2w[i] = c2w[i];
  P.rows = rows; P.cols = cols; P.n_rows = n_rows; P.n_cols = n_cols; P.patch = patch > 1 ? patch : 1;
  P.origins = origins; P.dirs = dirs; P.pixel_area = pixel_area;
  const int64_t n = static_cast<int64_t>(n_rows) * n_cols;
  for (int64_t i = 0; i < n; ++i) raygen_one(P, i);
  return 0;
}

// executor of the backward chains of backward.cuh: every step is a plain loop over the items of one kernel
struct HostExec {
  void dgrad(const float* dY, int ldy, int ny, const __half* W, int ldw, const __half* X, int ldx, float* dX, int lddx,
             int nx, int64_t rows, const int* rows_dev) {
    for (int64_t i = 0; i < rows * nx; ++i) mlp_dgrad_one(dY, ldy, ny, W, ldw, X, ldx, dX, lddx, nx, rows_dev, i);
  }
  void wgrad_h(const float* A, int lda, int na, const __half* B, int ldb, int nb, int64_t rows, float* C, int ldc,
               const int* rows_dev, const int* bmap) {
    for (int64_t i = 0; i < mlp_wgrad_items(rows, na, nb); ++i)
      mlp_wgrad_one<__half>(A, lda, na, B, ldb, nb, rows, C, ldc, rows_dev, bmap, i);
  }
  void wgrad_f(const float* A, int lda, int na, const float* B, int ldb, int nb, int64_t rows, float* C, int ldc) {
    for (int64_t i = 0; i < mlp_wgrad_items(rows, na, nb); ++i)
      mlp_wgrad_one<float>(A, lda, na, B, ldb, nb, rows, C, ldc, nullptr, nullptr, i);
  }
  void scatter2(const GridDev& G, bool linf, bool sel, const float* xyz, const float* dX, int lddx, int col0, float* g,
                int64_t points, const int* rows_dev) {
    for (int64_t i = 0; i < points * G.n_levels; ++i) grid_scatter_one<2>(G, linf, sel, xyz, dX, lddx, col0, g, rows_dev, i);
  }
  void scatter8(const GridDev& G, bool linf, bool sel, const float* xyz, const float* dX, int lddx, int col0, float* g,
                int64_t points, const int* rows_dev) {
    for (int64_t i = 0; i < points * G.n_levels; ++i) grid_scatter_one<8>(G, linf, sel, xyz, dX, lddx, col0, g, rows_dev, i);
  }
  void feat_rows_assign(const FeatBwdParams& P, int64_t items) { for (int64_t i = 0; i < items; ++i) feat_rows_assign_one(P, i); }
  void im2col_f32(const float* X, __half* Xcol, int64_t items) { for (int64_t i = 0; i < items; ++i) conv_im2col_one<float>(X, Xcol, i); }
  void im2col_f16(const __half* X, __half* Xcol, int64_t items) { for (int64_t i = 0; i < items; ++i) conv_im2col_one<__half>(X, Xcol, i); }
  void conv_fwd(const __half* Xcol, const __half* W, const float* b, __half* Y, int relu, int64_t items) {
    for (int64_t i = 0; i < items; ++i) conv_fwd_one(Xcol, W, b, Y, relu, i);
  }
  void conv_mean_bwd(const float* d_out, float* d_y, int64_t items) { for (int64_t i = 0; i < items; ++i) conv_mean_bwd_one(d_out, d_y, i); }
  void col2im(const float* dXcol, const __half* mask, float* dX, int64_t items) {
    for (int64_t i = 0; i < items; ++i) conv_col2im_one(dXcol, mask, dX, i);
  }
  void bias_grad(const float* dY, int64_t rows, float* g_b) {
    for (int64_t i = 0; i < ((rows + kSlabRows - 1) / kSlabRows) * kConvC; ++i) conv_bias_grad_one(dY, rows, g_b, i);
  }
  void feat_hidden(const FeatBwdParams& P, int64_t items) { for (int64_t i = 0; i < items; ++i) feat_hidden_one(P, i); }
  void feat_positions(const FeatBwdParams& P, int64_t items) { for (int64_t i = 0; i < items; ++i) feat_positions_one(P, i); }
  void sigmoid_bwd(const float* d_rgb, const __half* pre, int ldp, float* d_pre, int64_t items) {
    for (int64_t i = 0; i < items; ++i) sigmoid_bwd_one(d_rgb, pre, ldp, d_pre, i);
  }
  void density_bwd(const float* d_density, const __half* o, int ldo, const float* sel, const float* d_geo, int ldg,
                   int col_geo, float* d_o, int n_o, int64_t items) {
    for (int64_t i = 0; i < items; ++i) density_bwd_one(d_density, o, ldo, sel, d_geo, ldg, col_geo, d_o, n_o, i);
  }
};

static void fill_grid(GridDev& G, const double* levels, int n_levels, int n_features) {
  G.table = nullptr; G.n_levels = n_levels; G.n_features = n_features;
  for (int l = 0; l < n_levels; ++l) {
    const double* v = levels + l * 5;  // {scale, res, size, offset, hashed}
    G.lv[l].scale = static_cast<float>(v[0]);
    G.lv[l].res = static_cast<uint32_t>(v[1]);
    G.lv[l].size = static_cast<uint32_t>(v[2]);
    G.lv[l].offset = static_cast<uint32_t>(v[3]);
    G.lv[l].hashed = static_cast<uint32_t>(v[4]);
  }
}

// mirrors snrf_feature_backward (HOST pointers): levels[e][l] = {scale, res, size, offset, hashed} as doubles
int emu_feature_backward(const float* origins, const float* dirs, const float* sam_t, const float* sam_w, long long n_rays,
                         const float* d_out, int n_out, const unsigned short* x_f16, const unsigned short* w1_f16,
                         const unsigned short* w2_f16, const double* levels /*[2][12][5]*/, float* g_w1, float* g_w2,
                         float* g_table0, float* g_table1, float* hbar_out /*[N,256] or null*/, float cutoff,
                         int* n_rows_out /*[1] or null*/) {
  FeatBwdParams P;
  P.cutoff = cutoff;
  P.origins = origins; P.dirs = dirs; P.sam_t = sam_t; P.sam_w = sam_w; P.d_out = d_out;
  P.x = reinterpret_cast<const __half*>(x_f16);
  P.w1 = reinterpret_cast<const __half*>(w1_f16);
  P.w2 = reinterpret_cast<const __half*>(w2_f16);
  P.n_rays = n_rays; P.n_out = n_out;
  for (int e = 0; e < 2; ++e) fill_grid(P.enc[e], levels + e * 12 * 5, 12, 8);
  const int64_t n = n_rays, rows = n * kBwdK;
  std::vector<float> d_hbar(n * kBwdHid), hbar(n * kBwdHid), dh(rows * kBwdHid), dx(rows * kBwdIn), xyz(rows * 3);
  P.d_hbar = d_hbar.data(); P.hbar = hbar.data(); P.dh = dh.data(); P.dx = dx.data(); P.xyz = xyz.data();
  std::vector<int> row_map(rows), row_start(n), row_k(n);
  int n_rows = 0;
  P.n_rows = &n_rows; P.row_map = row_map.data(); P.row_start = row_start.data(); P.row_k = row_k.data();
  P.g_w1 = g_w1; P.g_w2 = g_w2; P.g_table[0] = g_table0; P.g_table[1] = g_table1;
  HostExec ex;
  feat_backward_chain(P, ex);
  if (hbar_out) for (int64_t i = 0; i < n * kBwdHid; ++i) hbar_out[i] = hbar[i];
  if (n_rows_out) *n_rows_out = n_rows;
  return 0;
}

// pre-pass of the bucketed feature kernel (sam_bucket.cu): counts[4], lists[4][n]
int emu_bucket_assign(const float* sam_w, float eps, long long n, int* counts, int* lists) {
  for (int b = 0; b < kFeatBuckets; ++b) counts[b] = 0;
  for (int64_t r = 0; r < n; ++r) bucket_assign_one(sam_w, eps, counts, lists, n, r, nullptr);
  return 0;
}

// training-mode spacing bins of the initial sampler (march.cu, JIT instantiation): out[n, n_bins + 1]
int emu_jittered_bins(const float* t_rand, long long n, int n_bins, float* out) {
  for (int64_t r = 0; r < n; ++r)
    for (int j = 0; j <= n_bins; ++j) out[r * (n_bins + 1) + j] = jittered_bin(j, n_bins, t_rand[r]);
  return 0;
}

// mirrors snrf_pick_samples
int emu_pick_samples(const float* w, const float* starts, const float* ends, long long n, int S, int k, float sharpen,
                     float* sam_t, float* sam_w) {
  for (int64_t i = 0; i < n; ++i) pick_samples_one(w, starts, ends, S, k, sharpen, sam_t, sam_w, i);
  return 0;
}

// mirror snrf_ray_op_backward modes 0 and 3
int emu_weights_backward(const float* deltas, const float* dens, const float* g_w, float* d_dens, long long n, int S) {
  for (int64_t i = 0; i < n; ++i) weights_bwd_one(deltas, dens, g_w, d_dens, S, i);
  return 0;
}
int emu_rgb_backward(const float* rgb, const float* w, const float* g_out, int bg_fixed, const float* bg, float* d_rgb,
                     float* d_w, long long n, int S) {
  for (int64_t i = 0; i < n * S; ++i)
    rgb_bwd_one(rgb, w, g_out, bg_fixed, bg ? bg[0] : 0.f, bg ? bg[1] : 0.f, bg ? bg[2] : 0.f, d_rgb, d_w, S, i);
  return 0;
}

// mirrors snrf_patch_aggregate_backward (HOST pointers; w1 / w2 are fp16 bit patterns of torch's [256,256,3,3] weights)
int emu_conv_backward(const float* feat_in, long long n_patches, const float* d_out, const unsigned short* w1,
                      const float* b1, const unsigned short* w2, const float* b2, float* g_w1, float* g_b1, float* g_w2,
                      float* g_b2, float* d_feat, float* out_fwd /*[P,256] forward value of the recomputation, or null*/) {
  ConvBwdParams P;
  memset(&P, 0, sizeof(P));
  const int64_t n = n_patches * kConvPos;
  P.feat_in = feat_in; P.d_out = d_out; P.rows = n;
  P.w1 = reinterpret_cast<const __half*>(w1); P.w2 = reinterpret_cast<const __half*>(w2); P.b1 = b1; P.b2 = b2;
  std::vector<__half> xcol1(n * kConvK), xcol2(n * kConvK), hid(n * kConvC);
  std::vector<float> d_y(n * kConvC), d_xcol(n * kConvK), d_hid(n * kConvC);
  P.xcol1 = xcol1.data(); P.xcol2 = xcol2.data(); P.hid = hid.data();
  P.d_y = d_y.data(); P.d_xcol = d_xcol.data(); P.d_hid = d_hid.data();
  P.g_w1 = g_w1; P.g_b1 = g_b1; P.g_w2 = g_w2; P.g_b2 = g_b2; P.d_feat = d_feat;
  HostExec ex;
  conv_backward_chain(P, ex);
  if (out_fwd) {  // second conv + patch mean from the recomputed activations, to check the recomputation itself
    std::vector<__half> y2(n * kConvC);
    for (int64_t i = 0; i < n * kConvC; ++i) {
      const int64_t row = i / kConvC; const int o = static_cast<int>(i % kConvC);
      float a = 0.f;
      for (int k = 0; k < kConvK; ++k) a += __half2float(xcol2[row * kConvK + k]) * __half2float(P.w2[static_cast<size_t>(o) * kConvK + k]);
      out_fwd[(row / kConvPos) * kConvC + o] += (a + b2[o]) / static_cast<float>(kConvPos);
    }
  }
  return 0;
}

// mirrors the backward half of snrf_field_backward: the forward activations (what launch_field_backward recomputes
// with the query kernels) are supplied by the caller - the test takes them from the oracle.
// acts: fp16 bit patterns x[n,width] h1[n,hidden] o[n,16] and, for nerfacto colour, hx[n,32] g1[n,64] g2[n,64] pre3[n,16]
int emu_field_backward(int which, const float* xyz, long long n, const float* d_density, const float* d_rgb,
                       const double* levels, int n_levels, const unsigned short* w1, const unsigned short* w2,
                       const unsigned short* wh1, const unsigned short* wh2, const unsigned short* wh3,
                       const unsigned short* x, const unsigned short* h1, const unsigned short* o,
                       const unsigned short* hx, const unsigned short* g1, const unsigned short* g2,
                       const unsigned short* pre3, const float* sel, float* g_base, float* g_head) {
  FieldBwdParams P;
  memset(&P, 0, sizeof(P));
  P.xyz = xyz; P.d_density = d_density; P.d_rgb = d_rgb; P.n = n; P.which = which;
  fill_grid(P.grid, levels, n_levels, 2);
  auto H = [](const unsigned short* p) { return reinterpret_cast<__half*>(const_cast<unsigned short*>(p)); };
  P.w1 = H(w1); P.w2 = H(w2); P.wh1 = H(wh1); P.wh2 = H(wh2); P.wh3 = H(wh3);
  P.x = H(x); P.h1 = H(h1); P.o = H(o); P.hx = H(hx); P.g1 = H(g1); P.g2 = H(g2); P.pre3 = H(pre3);
  P.sel = const_cast<float*>(sel);
  std::vector<float> d_pre3(n * 4), d_g2(n * 64), d_g1(n * 64), d_hx(n * 32), d_o(n * 16), d_h1(n * 64), d_x(n * 32);
  P.d_pre3 = d_pre3.data(); P.d_g2 = d_g2.data(); P.d_g1 = d_g1.data(); P.d_hx = d_hx.data(); P.d_o = d_o.data();
  P.d_h1 = d_h1.data(); P.d_x = d_x.data();
  P.g_base = g_base; P.g_head = g_head;
  HostExec ex;
  field_backward_chain(P, ex);
  return 0;
}

}  // extern "C"
