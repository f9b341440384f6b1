// Tile exchange by a push kernel (replication mode 2): small CTAs on a high-priority side stream read this rank's freshly
// rendered feature rows and store them into the frame buffers of the other ranks through their peer-mapped addresses
// (plain st.global over NVLink / NVSwitch).  Measured on 8 B200: the copy engines move a chunk to 7 peers at ~350 GB/s
// in aggregate (one engine per destination), less than half of what the links carry; stores issued from SMs are limited
// by the link instead.  The kernel is tiny on purpose - 128 threads at no more than 32 registers, no shared memory: the
// resident march CTAs (3 x 256 threads x 80 registers) and the feature kernel's CTA (512 threads x 120 registers) both
// leave exactly 4096 registers of an SM free, so a push CTA starts beside them at once instead of waiting for one of
// those persistent CTAs to retire, and the exchange of chunk c overlaps the march / feature kernels of chunk c + 1.
// (The first version used 256-thread CTAs at 26 registers = 8192 allocated: it did not fit beside either kernel.)
#include <stdlib.h>

#include "kernels.cuh"

namespace snrf {
namespace {

struct PushDst {
  uint4* p[kMaxPeers];
};

constexpr int kPushThreads = 128;

__global__ void __launch_bounds__(kPushThreads, 16) push_rows_kernel(const uint4* __restrict__ src, PushDst dst, int n_dst,
                                                                      size_t n16) {
  const size_t stride = static_cast<size_t>(gridDim.x) * kPushThreads;
  size_t i = static_cast<size_t>(blockIdx.x) * kPushThreads + threadIdx.x;
  // two rows pieces in flight per thread (streamed loads: the rows were written moments ago and are not read again here)
  for (; i + stride < n16; i += 2 * stride) {
    const uint4 a = __ldcs(src + i), b = __ldcs(src + i + stride);
#pragma unroll
    for (int d = 0; d < kMaxPeers; ++d)
      if (d < n_dst) {
        dst.p[d][i] = a;
        dst.p[d][i + stride] = b;
      }
  }
  if (i < n16) {
    const uint4 a = __ldcs(src + i);
#pragma unroll
    for (int d = 0; d < kMaxPeers; ++d)
      if (d < n_dst) dst.p[d][i] = a;
  }
}

}  // namespace

cudaError_t launch_push_rows(const void* src, void* const* dst, int n_dst, size_t bytes, cudaStream_t stream) {
  if (n_dst <= 0 || bytes == 0) return cudaSuccess;
  if (n_dst > kMaxPeers || (bytes & 15u) || (reinterpret_cast<uintptr_t>(src) & 15u)) return cudaErrorInvalidValue;
  PushDst D;
  for (int d = 0; d < kMaxPeers; ++d) {
    D.p[d] = d < n_dst ? reinterpret_cast<uint4*>(dst[d]) : nullptr;
    if (d < n_dst && (reinterpret_cast<uintptr_t>(dst[d]) & 15u)) return cudaErrorInvalidValue;
  }
  static const int grid = [] { const char* e = getenv("SNRF_PUSH_GRID"); return e ? atoi(e) : 64; }();
  push_rows_kernel<<<grid > 0 ? grid : 64, kPushThreads, 0, stream>>>(reinterpret_cast<const uint4*>(src), D, n_dst, bytes / 16);
  return cudaGetLastError();
}

}  // namespace snrf
