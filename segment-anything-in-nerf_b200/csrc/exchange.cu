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
#include <string.h>

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

// The same exchange driven by the TMA unit (opt-in: SNRF_PUSH=tma): one thread per CTA streams 8 KB slices of the rows
// global -> shared with cp.async.bulk, and from shared memory to every destination with cp.async.bulk stores (three
// stages: the load of slice i + 1 is in flight while the stores of slices i and i - 1 drain).  The idea: no LSU
// instruction touches the data, so a link that pushes back stalls the TMA queue of that CTA and not the load / store
// pipe of the march or feature warps resident on the same SM (with plain stores the co-resident kernels lose 0.16 ms
// per frame at N = 8).  Measured at N = 2 (gpurun call 18, bit-identity tests green): 193.8 Mrays/s against 218.1 with
// the store kernel - its 24 KB of shared memory per CTA push the SM into the next shared-memory carve-out, and the march
// kernel's gathers live on the L1 that is left (the same effect that makes 4 march CTAs per SM slower than 3).  Kept
// for the comparison; the store kernel is the default.
constexpr int kTmaStages = 3;
constexpr uint32_t kTmaSlice = 8192;
constexpr uint32_t kTmaSmem = kTmaStages * kTmaSlice + kTmaStages * 8;

__device__ __forceinline__ void bulk_store_1d(void* dst, uint32_t src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(dst), "r"(src_smem), "r"(bytes) : "memory");
}

struct PushDstBytes {
  char* p[kMaxPeers];
};

__global__ void __launch_bounds__(32) push_rows_tma_kernel(const char* __restrict__ src, PushDstBytes dst, int n_dst, size_t bytes) {
  extern __shared__ __align__(128) unsigned char sm[];
  if (threadIdx.x != 0) return;
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + kTmaStages * kTmaSlice);
  for (int s = 0; s < kTmaStages; ++s) mbar_init(smem_u32(&bar[s]), 1);
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  const size_t n_slices = (bytes + kTmaSlice - 1) / kTmaSlice;
  // this CTA's slices: blockIdx.x, blockIdx.x + gridDim.x, ...
  const size_t mine = blockIdx.x < n_slices ? (n_slices - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  auto slice_off = [&](size_t i) { return (blockIdx.x + i * gridDim.x) * static_cast<size_t>(kTmaSlice); };
  auto slice_len = [&](size_t i) {
    const size_t off = slice_off(i);
    return static_cast<uint32_t>(bytes - off < kTmaSlice ? bytes - off : kTmaSlice);
  };
  auto load = [&](size_t i) {
    const int st = static_cast<int>(i % kTmaStages);
    mbar_expect_tx(smem_u32(&bar[st]), slice_len(i));
    bulk_load_1d(smem_u32(sm + st * kTmaSlice), src + slice_off(i), slice_len(i), smem_u32(&bar[st]));
  };
  if (mine > 0) load(0);
  for (size_t i = 0; i < mine; ++i) {
    const int st = static_cast<int>(i % kTmaStages);
    if (i + 1 < mine) {
      // slice i + 1 goes into the stage slice i - 2 was stored from: all but the newest store group must have been read
      if (i >= 2) asm volatile("cp.async.bulk.wait_group.read 1;\n" ::: "memory");
      load(i + 1);
    }
    mbar_wait(smem_u32(&bar[st]), static_cast<uint32_t>((i / kTmaStages) & 1));
    const size_t off = slice_off(i);
    const uint32_t len = slice_len(i);
    for (int d = 0; d < n_dst; ++d) bulk_store_1d(dst.p[d] + off, smem_u32(sm + st * kTmaSlice), len);
    asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
  }
  asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory");  // the writes are complete when the kernel is
}

// Push through the NVSwitch multicast alias of the frame buffer: ONE multimem.st per 16 bytes reaches every rank (the
// switch replicates it), so the SM issues a seventh of the store instructions of the peer-pointer kernel at N = 8 and the
// rank's NVLink egress carries the rows once instead of seven times.  The ingress of every rank is unchanged.
__global__ void __launch_bounds__(kPushThreads, 16) push_rows_mc_kernel(const uint4* __restrict__ src, uint4* mc, size_t n16) {
  const size_t stride = static_cast<size_t>(gridDim.x) * kPushThreads;
  for (size_t i = static_cast<size_t>(blockIdx.x) * kPushThreads + threadIdx.x; i < n16; i += stride) {
    const uint4 v = __ldcs(src + i);
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};\n" ::"l"(mc + i), "f"(__uint_as_float(v.x)),
                 "f"(__uint_as_float(v.y)), "f"(__uint_as_float(v.z)), "f"(__uint_as_float(v.w))
                 : "memory");
  }
}

}  // namespace

cudaError_t launch_push_rows_mc(const void* src, void* mc_dst, size_t bytes, cudaStream_t stream) {
  if (bytes == 0) return cudaSuccess;
  if ((bytes & 15u) || (reinterpret_cast<uintptr_t>(src) & 15u) || (reinterpret_cast<uintptr_t>(mc_dst) & 15u)) return cudaErrorInvalidValue;
  static const int grid = [] { const char* e = getenv("SNRF_PUSH_GRID"); return e ? atoi(e) : 64; }();
  push_rows_mc_kernel<<<grid > 0 ? grid : 64, kPushThreads, 0, stream>>>(reinterpret_cast<const uint4*>(src),
                                                                        reinterpret_cast<uint4*>(mc_dst), bytes / 16);
  return cudaGetLastError();
}

cudaError_t launch_push_rows(const void* src, void* const* dst, int n_dst, size_t bytes, cudaStream_t stream) {
  if (n_dst <= 0 || bytes == 0) return cudaSuccess;
  if (n_dst > kMaxPeers || (bytes & 15u) || (reinterpret_cast<uintptr_t>(src) & 15u)) return cudaErrorInvalidValue;
  PushDst D;
  for (int d = 0; d < kMaxPeers; ++d) {
    D.p[d] = d < n_dst ? reinterpret_cast<uint4*>(dst[d]) : nullptr;
    if (d < n_dst && (reinterpret_cast<uintptr_t>(dst[d]) & 15u)) return cudaErrorInvalidValue;
  }
  static const bool tma = [] { const char* e = getenv("SNRF_PUSH"); return e && strcmp(e, "tma") == 0; }();
  if (tma) {
    static const int tgrid = [] { const char* e = getenv("SNRF_PUSH_GRID"); return e ? atoi(e) : 48; }();
    PushDstBytes B;
    for (int d = 0; d < kMaxPeers; ++d) B.p[d] = d < n_dst ? static_cast<char*>(dst[d]) : nullptr;
    push_rows_tma_kernel<<<tgrid > 0 ? tgrid : 48, 32, kTmaSmem, stream>>>(static_cast<const char*>(src), B, n_dst, bytes);
    return cudaGetLastError();
  }
  static const int grid = [] { const char* e = getenv("SNRF_PUSH_GRID"); return e ? atoi(e) : 64; }();
  push_rows_kernel<<<grid > 0 ? grid : 64, kPushThreads, 0, stream>>>(reinterpret_cast<const uint4*>(src), D, n_dst, bytes / 16);
  return cudaGetLastError();
}

}  // namespace snrf
