// Kernel C (round 2) - the feature MLP's output layer as a warp-specialised, TMA-fed tcgen05 GEMM:
//     out[M, N] fp32 = act(A[M, 256] fp16 x W[N, 256]^T + bias),   N = 256 (SAM) or 192 (ClipSeg)
// Reference: samnerf/sam_field.py:51-61,84-94 (the CutlassMLP output layer), applied once per ray to the weighted hidden
// sum that sam.cu / sam_bucket.cu produce.
//
// The layer is a streaming op - 512 B read and 1 024 B written per ray against 131 kFLOP - so the design goal is to
// keep HBM busy, not the tensor cores.  The round-1 kernel (gemm.cu, still used for the 3x3 conv taps and the fused
// multicast / peer stores) staged A with LDG -> STS by all threads, ran the MMAs, and wrote the tile out, one after the
// other: 0.25 of the HBM roofline.  Here three roles run concurrently in a persistent CTA (one per SM):
//
//   warp 0  producer : one lane issues cp.async.bulk.tensor.2d (TMA) loads of A, 128 rows x 64 columns (16 KB,
//                      SWIZZLE_128B) per stage, 4-stage ring with full / empty mbarriers; W (<= 128 KB, already in the
//                      tcgen05 core-matrix layout in HBM) arrives once per CTA through one cp.async.bulk copy.
//   warp 1  MMA      : one lane issues 16 tcgen05.mma (M = 128, N, K = 16) per tile into one of two TMEM accumulators
//                      (2 x 256 columns); tcgen05.commit releases the A stage and hands the accumulator over.
//   warps 2-5 epilogue: tcgen05.ld 32 columns at a time -> + bias / ReLU -> 128-byte-swizzled staging tile in shared
//                      memory (conflict-free 16-byte stores, one row per lane) -> TMA store of the 128 x 32 fp32 box
//                      (cp.async.bulk.tensor.2d.global.shared::cta); two staging buffers, so the store of chunk c
//                      overlaps the TMEM read of chunk c + 1, and the whole epilogue of tile i overlaps the loads and
//                      MMAs of tile i + 1.
// Tail tiles need no code: TMA zero-fills rows past M on load and clips them on store.
#include <cuda.h>

#include "kernels.cuh"

namespace snrf {
namespace {

constexpr int kK = 256;
constexpr int kStages = 4;                        // A ring: one stage = one 64-column k-block of a 128-row tile
constexpr uint32_t kStageBytes = 128 * 64 * 2;    // 16384
constexpr uint32_t kStoreBytes = 128 * 32 * 4;    // 16384: one 128 x 32 fp32 box
constexpr uint32_t kWBytesMax = 256 * kK * 2;     // 131072
constexpr uint32_t kOffA = 0, kOffStore = kStages * kStageBytes, kOffW = kOffStore + 2 * kStoreBytes,
                   kOffBias = kOffW + kWBytesMax, kOffBar = kOffBias + 1024;
constexpr uint32_t kSmemBytes = kOffBar + 256 + 1024;  // + slack for the 1024-byte alignment of the swizzled tiles
constexpr int kThreads = 192;

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int x, int y, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(dst),
               "l"(map), "r"(x), "r"(y), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int x, int y) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];\n" ::"l"(map), "r"(src), "r"(x),
               "r"(y)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;\n" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory"); }
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 128;\n" ::: "memory"); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
  return pred != 0;
}

// shared-memory descriptor of a K-major operand tile in the 128-byte-swizzled layout TMA writes (rows of 128 B, 8-row
// groups of 1024 B): SBO = 1024 B, LBO unused (1), layout type 2 = SWIZZLE_128B, version 1 (cute::UMMA::SmemDescriptor)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr >> 4) & 0x3FFFu);
  d |= static_cast<uint64_t>(1u) << 16;
  d |= static_cast<uint64_t>((1024u >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;
  d |= 2ull << 61;
  return d;
}

struct TmaGemmParams {
  const __half* w;    // [N x 256] fp16, core-matrix layout (as GemmParams::w)
  const float* bias;  // [N] or null
  int64_t m;
  int n;
  int relu;
};

// F16 = false: fp32 output rows (boxes of 128 x 32 floats); F16 = true: fp16 output rows (boxes of 128 x 64 halfs, two
// 32-column TMEM chunks per box) - the fp16 feature rows of snrf_set_feature_dtype and the input rows of the conv head
template <bool F16>
__global__ void __launch_bounds__(kThreads, 1)
tapgemm_tma_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_out, const TmaGemmParams P) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B tiles want 1024-byte alignment
  unsigned char* gen = smem_raw + (base - smem_u32(smem_raw));
  float* s_bias = reinterpret_cast<float*>(gen + kOffBias);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(gen + kOffBar + 128);
  const uint32_t bar0 = base + kOffBar;
  auto a_full = [&](int s) { return bar0 + 8u * s; };
  auto a_empty = [&](int s) { return bar0 + 8u * (kStages + s); };
  auto t_full = [&](int a) { return bar0 + 8u * (2 * kStages + a); };
  auto t_empty = [&](int a) { return bar0 + 8u * (2 * kStages + 2 + a); };
  const uint32_t w_full = bar0 + 8u * (2 * kStages + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t w_bytes = static_cast<uint32_t>(P.n) * kK * 2;
  for (int i = threadIdx.x; i < 256; i += kThreads) s_bias[i] = (P.bias && i < P.n) ? P.bias[i] : 0.f;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(a_full(s), 1);
      mbar_init(a_empty(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(t_full(a), 1);
      mbar_init(t_empty(a), 128);
    }
    mbar_init(w_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];\n" ::"l"(&map_out) : "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32(s_tmem), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  const int64_t n_tiles = (P.m + 127) / 128;
  const int n_chunks = P.n / 32;  // 8 or 6

  if (warp == 0) {
    // ===== producer =====
    if (elect_one()) {
      mbar_expect_tx(w_full, w_bytes);
      bulk_load_1d(base + kOffW, P.w, w_bytes, w_full);
      uint32_t it = 0;
      for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int kb = 0; kb < kK / 64; ++kb, ++it) {
          const int s = it % kStages;
          const uint32_t ph = (it / kStages) & 1u;
          mbar_wait(a_empty(s), ph ^ 1u);  // first pass through the ring: passes at once
          mbar_expect_tx(a_full(s), kStageBytes);
          tma_load_2d(base + kOffA + s * kStageBytes, &map_a, kb * 64, static_cast<int>(tile * 128), a_full(s));
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    const uint32_t idesc = umma_idesc_f16(128, static_cast<uint32_t>(P.n));
    const uint32_t b_addr = base + kOffW;
    const bool leader = elect_one();
    if (leader) mbar_wait(w_full, 0);
    __syncwarp();
    uint32_t it = 0, t = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++t) {
      const int acc = t & 1;
      if (leader) {
        mbar_wait(t_empty(acc), ((t >> 1) & 1u) ^ 1u);  // the epilogue has drained this accumulator
        tc_fence_after();
        for (int kb = 0; kb < kK / 64; ++kb, ++it) {
          const int s = it % kStages;
          mbar_wait(a_full(s), (it / kStages) & 1u);
          tc_fence_after();
          const uint32_t a_addr = base + kOffA + s * kStageBytes;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int ks = kb * 4 + j;  // global k-step of 16
            umma_f16(tmem_base + acc * 256, umma_desc_sw128(a_addr + j * 32), umma_desc(b_addr + ks * 256, 128, kK * 16), idesc,
                     ks > 0 ? 1u : 0u);
          }
          umma_commit(a_empty(s));  // the stage is free once these MMAs have read it
        }
        umma_commit(t_full(acc));
      } else {
        it += kK / 64;
      }
      __syncwarp();
    }
  } else {
    // ===== epilogue: warps 2..5; a warp reads the TMEM lane quarter (warp % 4) =====
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;  // tile row = TMEM lane
    const bool issuer = threadIdx.x == 64;  // first lane of warp 2
    uint32_t t = 0, n_store = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++t) {
      const int acc = t & 1;
      mbar_wait(t_full(acc), (t >> 1) & 1u);
      tc_fence_after();
      for (int c = 0; c < n_chunks; ++c) {
        const int sb = n_store & 1;
        const bool opens = !F16 || (c & 1) == 0;               // first chunk of a staging box
        const bool closes = !F16 || (c & 1) == 1 || c == n_chunks - 1;  // last chunk of a staging box
        if (opens) {
          if (issuer) bulk_wait_read<1>();  // the store issued two boxes ago has finished reading this staging buffer
          epi_bar();
        }
        float v[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * 256 + c * 32, v);
        if (c == n_chunks - 1) {  // accumulator fully read: hand it back to the MMA warp
          tc_fence_before();
          mbar_arrive(t_empty(acc));
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          v[i] += s_bias[c * 32 + i];
          if (P.relu) v[i] = fmaxf(v[i], 0.f);
        }
        unsigned char* st = gen + kOffStore + sb * kStoreBytes + row * 128;
        if (F16) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {  // 8 halfs per 16-byte piece; this chunk fills pieces (c & 1) * 4 .. + 3 of the row
            const uint4 o = make_uint4(f2_to_h2(v[8 * q + 0], v[8 * q + 1]), f2_to_h2(v[8 * q + 2], v[8 * q + 3]),
                                       f2_to_h2(v[8 * q + 4], v[8 * q + 5]), f2_to_h2(v[8 * q + 6], v[8 * q + 7]));
            *reinterpret_cast<uint4*>(st + ((((c & 1) * 4 + q) ^ (row & 7)) << 4)) = o;
          }
        } else {
#pragma unroll
          for (int q = 0; q < 8; ++q)  // 128-byte swizzle: 16-byte piece index ^= row % 8
            *reinterpret_cast<float4*>(st + ((q ^ (row & 7)) << 4)) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        }
        if (closes) {
          fence_async_smem();
          epi_bar();
          if (issuer) {
            tma_store_2d(&map_out, base + kOffStore + sb * kStoreBytes, F16 ? (c >> 1) * 64 : c * 32, static_cast<int>(tile * 128));
            bulk_commit();
          }
          ++n_store;
        }
      }
    }
    if (issuer) bulk_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                              const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeFn get_encode() {
  static EncodeFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeFn>(p);
  }
  return fn;
}

bool make_map(CUtensorMap* map, CUtensorMapDataType type, size_t elem, void* ptr, uint64_t inner, uint64_t rows, uint32_t box_inner,
              uint32_t box_rows) {
  EncodeFn enc = get_encode();
  if (!enc) return false;
  const cuuint64_t dims[2] = {inner, rows};
  const cuuint64_t strides[1] = {inner * elem};
  const cuuint32_t box[2] = {box_inner, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  return enc(map, type, 2, ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

// Returns cudaErrorNotSupported when this launch cannot take the TMA path (the caller then uses gemm.cu).
cudaError_t launch_tapgemm_tma(const GemmParams& P, int sm_count, cudaStream_t stream) {
  if (P.taps != 1 || (P.out_mode != 0 && P.out_mode != 1) || P.out_mc || P.n_peers > 0 || (P.n != 256 && P.n != 192))
    return cudaErrorNotSupported;
  if (P.m <= 0) return cudaSuccess;
  const bool f16 = P.out_mode == 1;
  void* out = f16 ? static_cast<void*>(P.out_f16) : static_cast<void*>(P.out_f32);
  if ((reinterpret_cast<uintptr_t>(P.a) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(P.w)) & 15u)
    return cudaErrorNotSupported;  // TMA wants 16-byte aligned global addresses
  CUtensorMap map_a, map_out;
  if (!make_map(&map_a, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(P.a), kK, static_cast<uint64_t>(P.m), 64, 128))
    return cudaErrorNotSupported;
  if (f16 ? !make_map(&map_out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, out, static_cast<uint64_t>(P.n), static_cast<uint64_t>(P.m), 64, 128)
          : !make_map(&map_out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, out, static_cast<uint64_t>(P.n), static_cast<uint64_t>(P.m), 32, 128))
    return cudaErrorNotSupported;
  static bool configured_dev[64] = {false};
  int dev_id = 0;
  if (cudaGetDevice(&dev_id) != cudaSuccess || dev_id < 0 || dev_id >= 64) return cudaErrorInvalidDevice;
  if (!configured_dev[dev_id]) {
    cudaError_t e = cudaFuncSetAttribute(tapgemm_tma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(tapgemm_tma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
    if (e != cudaSuccess) return e;
    configured_dev[dev_id] = true;
  }
  TmaGemmParams T;
  T.w = P.w;
  T.bias = P.bias;
  T.m = P.m;
  T.n = P.n;
  T.relu = P.relu;
  const int64_t n_tiles = (P.m + 127) / 128;
  const int grid = static_cast<int>(n_tiles < sm_count ? n_tiles : sm_count);
  if (f16)
    tapgemm_tma_kernel<true><<<grid, kThreads, kSmemBytes, stream>>>(map_a, map_out, T);
  else
    tapgemm_tma_kernel<false><<<grid, kThreads, kSmemBytes, stream>>>(map_a, map_out, T);
  return cudaGetLastError();
}

}  // namespace snrf
