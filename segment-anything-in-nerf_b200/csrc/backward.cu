// Kernels F1-F5 - backward of the feature-field branch (see backward.cuh for the math and the reference lines).
// First, simple version: one thread per output item, fp32 throughout, atomics for the reductions.  The dominant
// cost is the table scatter (16 x 24 x 8 corners x 8 features = 24 576 fp32 atomics per ray, ~98 KB of
// read-modify-write traffic per ray against 49 KB gathered by the forward pass); the fused tensor-core version that
// mirrors sam.cu is the next step once this one is parity-green on hardware.
#include "backward.cuh"

namespace snrf {
namespace {

constexpr int kThreads = 256;
constexpr int kSlabRows = 256;  // rows reduced per thread before one atomicAdd (weight gradients)

__global__ void bwd_dhbar_kernel(const FeatBwdParams P, int64_t items) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < items) bwd_dhbar_one(P, i);
}
__global__ void bwd_hidden_kernel(const FeatBwdParams P, int64_t items) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < items) bwd_hidden_one(P, i);
}
__global__ void bwd_dx_kernel(const FeatBwdParams P, int64_t items) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < items) bwd_dx_one(P, i);
}
template <typename TB>
__global__ void bwd_wgrad_kernel(const float* A, int na, const TB* B, int nb, int64_t rows, float* C, int64_t items) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < items) bwd_wgrad_one<TB>(A, na, B, nb, rows, kSlabRows, C, i);
}
__global__ void bwd_scatter_kernel(const FeatBwdParams P, int64_t items) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < items) bwd_scatter_one(P, i);
}

inline unsigned blocks_for(int64_t items) { return static_cast<unsigned>((items + kThreads - 1) / kThreads); }

}  // namespace

size_t feat_bwd_scratch_floats(int64_t n_rays) {
  const int64_t b = n_rays < kBwdBlockRays ? n_rays : kBwdBlockRays;
  return static_cast<size_t>(b) * (2 * kBwdHid + kBwdK * kBwdHid + kBwdK * kBwdIn);
}

// P.d_hbar must point at feat_bwd_scratch_floats(P.n_rays) floats; hbar / dh / dx are carved out of it here.
cudaError_t launch_feat_backward(const FeatBwdParams& P0, cudaStream_t stream, int64_t* launches) {
  float* scratch = P0.d_hbar;
  for (int64_t r0 = 0; r0 < P0.n_rays; r0 += kBwdBlockRays) {
    const int64_t n = P0.n_rays - r0 < kBwdBlockRays ? P0.n_rays - r0 : kBwdBlockRays;
    FeatBwdParams P = P0;
    P.origins += 3 * r0;
    P.dirs += 3 * r0;
    P.sam_t += kBwdK * r0;
    P.sam_w += kBwdK * r0;
    P.d_out += static_cast<int64_t>(P.n_out) * r0;
    P.x += kBwdK * kBwdIn * r0;
    P.n_rays = n;
    P.d_hbar = scratch;
    P.hbar = P.d_hbar + n * kBwdHid;
    P.dh = P.hbar + n * kBwdHid;
    P.dx = P.dh + n * kBwdK * kBwdHid;
    const int64_t rows = n * kBwdK;
    int64_t items = n * kBwdHid;
    bwd_dhbar_kernel<<<blocks_for(items), kThreads, 0, stream>>>(P, items);
    bwd_hidden_kernel<<<blocks_for(items), kThreads, 0, stream>>>(P, items);
    items = rows * kBwdIn;
    bwd_dx_kernel<<<blocks_for(items), kThreads, 0, stream>>>(P, items);
    items = ((rows + kSlabRows - 1) / kSlabRows) * kBwdHid * kBwdIn;
    bwd_wgrad_kernel<__half><<<blocks_for(items), kThreads, 0, stream>>>(P.dh, kBwdHid, P.x, kBwdIn, rows, P.g_w1, items);
    items = ((n + kSlabRows - 1) / kSlabRows) * P.n_out * kBwdHid;
    bwd_wgrad_kernel<float><<<blocks_for(items), kThreads, 0, stream>>>(P.d_out, P.n_out, P.hbar, kBwdHid, n, P.g_w2, items);
    items = rows * 24;
    bwd_scatter_kernel<<<blocks_for(items), kThreads, 0, stream>>>(P, items);
    if (launches) *launches += 6;
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

}  // namespace snrf
