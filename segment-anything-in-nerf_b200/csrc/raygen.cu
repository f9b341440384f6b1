// Kernel E - camera ray generation (see raygen.cuh for the reference lines restated).  24 bytes written per ray
// against ~75 KB gathered by the render that follows: HBM-trivial, launched once per frame.
#include "raygen.cuh"

namespace snrf {

__global__ void raygen_kernel(const RayGenParams P, int64_t n) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) raygen_one(P, i);
}

cudaError_t launch_raygen(const RayGenParams& P, cudaStream_t stream) {
  const int64_t n = static_cast<int64_t>(P.n_rows) * P.n_cols;
  if (n <= 0) return cudaSuccess;
  raygen_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(P, n);
  return cudaGetLastError();
}

}  // namespace snrf
