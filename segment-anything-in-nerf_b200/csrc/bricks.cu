// Brick builder: cell-major re-layout of one level of an F = 2 hash-grid table (see BrickDev in kernels.cuh).
// One thread per cell: the 8 corner entries are fetched with the level's own index rule (common.cuh grid_index: dense
// x + y res + z res^2 with tcnn's `% size` wrap, or the coherent-prime hash), so a brick is a verbatim copy of what
// the 8 gathers of the table path would have returned.
#include "kernels.cuh"

namespace snrf {
namespace {

__global__ void brick_build_kernel(const GridDev G, int level, uint4* __restrict__ out, uint32_t n_cells) {
  const uint32_t cell = blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= n_cells) return;
  const GridLevel& L = G.lv[level];
  const uint32_t r = L.res;
  const uint32_t gx = cell % r, gy = (cell / r) % r, gz = cell / (r * r);
  const uint32_t* table = reinterpret_cast<const uint32_t*>(G.table);
  uint32_t v[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) v[c] = table[grid_index(L, gx + (c & 1), gy + ((c >> 1) & 1), gz + (c >> 2))];
  out[2 * static_cast<size_t>(cell)] = make_uint4(v[0], v[1], v[2], v[3]);
  out[2 * static_cast<size_t>(cell) + 1] = make_uint4(v[4], v[5], v[6], v[7]);
}

}  // namespace

cudaError_t launch_brick_build(const GridDev& G, int level, uint4* out, cudaStream_t stream) {
  const uint64_t r = G.lv[level].res;
  const uint64_t n = r * r * r;
  if (G.n_features != 2 || n == 0 || n > 0xFFFFFFFFull) return cudaErrorInvalidValue;
  brick_build_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(G, level, out, static_cast<uint32_t>(n));
  return cudaGetLastError();
}

}  // namespace snrf
