// TEST INFRASTRUCTURE ONLY - never loaded by the package.
//
// Host emulation of libsnrf's simple kernels (one thread = one item, no shared memory, no warp intrinsics): the very
// same __host__ __device__ bodies the CUDA kernels call are run in a plain loop on the CPU, so that their arithmetic,
// indexing and layouts can be checked against the oracle in a container without a GPU.  It says nothing about launch
// configuration, races between threads or device intrinsics - the `-m gpu` parity tests remain the real gate.
// Built by tests/emu/build_emu.py with nvcc (host code only; no CUDA call is ever made).
#include "../../segment-anything-in-nerf_b200/csrc/raygen.cuh"

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

}  // extern "C"
