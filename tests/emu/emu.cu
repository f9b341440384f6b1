// TEST INFRASTRUCTURE ONLY - never loaded by the package.
//
// Host emulation of libsnrf's simple kernels (one thread = one item, no shared memory, no warp intrinsics): the very
// same __host__ __device__ bodies the CUDA kernels call are run in a plain loop on the CPU, so that their arithmetic,
// indexing and layouts can be checked against the oracle in a container without a GPU.  It says nothing about launch
// configuration, races between threads or device intrinsics - the `-m gpu` parity tests remain the real gate.
// Built by tests/emu/build_emu.py with nvcc (host code only; no CUDA call is ever made).
#include "../../segment-anything-in-nerf_b200/csrc/raygen.cuh"
#include "../../segment-anything-in-nerf_b200/csrc/backward.cuh"

#include <vector>

using namespace snrf;

extern "C" {

// mirrors snrf_generate_rays: all pointers are HOST pointers here
int emu_generate_rays(const float* intr /*fx fy cx cy*/, int type, int has_dist, const float* dist, const float* c2w,
                      const int* rows, int n_rows, const int* cols, int n_cols, int patch, float* origins, float* dirs,
                      float* pixel_area) {
  RayGenParams P;
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

// mirrors snrf_feature_backward (HOST pointers): levels[e][l] = {scale, res, size, offset, hashed} as doubles
int emu_feature_backward(const float* origins, const float* dirs, const float* sam_t, const float* sam_w, long long n_rays,
                         const float* d_out, int n_out, const unsigned short* x_f16, const unsigned short* w1_f16,
                         const unsigned short* w2_f16, const double* levels /*[2][12][5]*/, float* g_w1, float* g_w2,
                         float* g_table0, float* g_table1, float* hbar_out /*[N,256] or null*/) {
  FeatBwdParams P;
  P.origins = origins; P.dirs = dirs; P.sam_t = sam_t; P.sam_w = sam_w; P.d_out = d_out;
  P.x = reinterpret_cast<const __half*>(x_f16);
  P.w1 = reinterpret_cast<const __half*>(w1_f16);
  P.w2 = reinterpret_cast<const __half*>(w2_f16);
  P.n_rays = n_rays; P.n_out = n_out;
  for (int e = 0; e < 2; ++e) {
    P.enc[e].table = nullptr; P.enc[e].n_levels = 12; P.enc[e].n_features = 8;
    for (int l = 0; l < 12; ++l) {
      const double* v = levels + (e * 12 + l) * 5;
      P.enc[e].lv[l].scale = static_cast<float>(v[0]);
      P.enc[e].lv[l].res = static_cast<uint32_t>(v[1]);
      P.enc[e].lv[l].size = static_cast<uint32_t>(v[2]);
      P.enc[e].lv[l].offset = static_cast<uint32_t>(v[3]);
      P.enc[e].lv[l].hashed = static_cast<uint32_t>(v[4]);
    }
  }
  const int64_t n = n_rays, rows = n * kBwdK;
  std::vector<float> d_hbar(n * kBwdHid), hbar(n * kBwdHid), dh(rows * kBwdHid), dx(rows * kBwdIn);
  P.d_hbar = d_hbar.data(); P.hbar = hbar.data(); P.dh = dh.data(); P.dx = dx.data();
  P.g_w1 = g_w1; P.g_w2 = g_w2; P.g_table[0] = g_table0; P.g_table[1] = g_table1;
  const int slab = 256;  // kSlabRows of backward.cu
  for (int64_t i = 0; i < n * kBwdHid; ++i) bwd_dhbar_one(P, i);
  for (int64_t i = 0; i < n * kBwdHid; ++i) bwd_hidden_one(P, i);
  for (int64_t i = 0; i < rows * kBwdIn; ++i) bwd_dx_one(P, i);
  for (int64_t i = 0; i < ((rows + slab - 1) / slab) * kBwdHid * kBwdIn; ++i)
    bwd_wgrad_one<__half>(P.dh, kBwdHid, P.x, kBwdIn, rows, slab, P.g_w1, i);
  for (int64_t i = 0; i < ((n + slab - 1) / slab) * n_out * kBwdHid; ++i)
    bwd_wgrad_one<float>(P.d_out, n_out, P.hbar, kBwdHid, n, slab, P.g_w2, i);
  for (int64_t i = 0; i < rows * 24; ++i) bwd_scatter_one(P, i);
  if (hbar_out) for (int64_t i = 0; i < n * kBwdHid; ++i) hbar_out[i] = hbar[i];
  return 0;
}

}  // extern "C"
