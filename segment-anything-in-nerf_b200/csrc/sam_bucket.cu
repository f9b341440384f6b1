// Kernel B' - kernel B (sam.cu) over rays bucketed by how many of their k = 16 picked samples still carry weight.
//
// The MeanRenderer weights are sharpened with w^10 and renormalised (samnerf/sam_model.py:244-248), so most of a ray's
// 16 slots end up far below the unit roundoff of the fp32 sum they are accumulated into: on the 800x800 scene-like
// frame 35 % of the rays have 1 slot with w >= 2^-24, 25 % have 2, 28 % at most 4, 12 % at most 8, 0.4 % more (oracle
// count; the slots are stored in descending weight order, so the significant ones are a prefix).  Kernel B spends the
// same 24 levels x 8 corners x 16 B of gathers and the same 128-row tensor-core tile on every slot.  Here a pre-pass
// sorts the rays into five buckets by their significant-slot count (<= 1, <= 2, <= 4, <= 8, 16) and one persistent launch
// runs tiles of 128 / SLOTS rays x SLOTS slots, bucket after bucket: same gather code, same tcgen05 tile, a log2(SLOTS)-step row
// reduction in the epilogue.  Every slot below SLOTS is evaluated with its real weight; only slots >= SLOTS - all
// of them below the cut-off - are dropped.  cut-off 0 drops exact zeros only (bit-for-bit the same sum up to fp32
// summation order); the default 2^-24 drops at most 1.2e-7 of total weight per ray (measured), i.e. less than one
// fp32 ulp of the accumulated sum, which is then rounded to fp16 anyway.  Rays whose weights are NaN (0/0,
// sam_model.py:248) count as fully significant and produce the same NaN row as kernel B.
//
// Default feature path since round 2 (snrf_set_feature_cutoff(ctx, < 0) selects kernel B): the whole GPU parity suite
// runs through it.  The gather, MMA issue, TMEM epilogue and barrier protocol are kernel B's; what is new is the
// row -> (ray, slot) mapping through the bucket list, the per-lane ray loads and the generalised reduction width.
#include "kernels.cuh"

namespace snrf {
namespace {

constexpr int kK = 16;      // slots per ray in sam_t / sam_w
constexpr int kIn = 192;    // encoder width
constexpr int kHid = 256;   // hidden width
constexpr uint32_t kSBO = kIn * 16;
constexpr uint32_t kW1Bytes = kHid * kIn * 2;    // 98304
constexpr uint32_t kATileBytes = 128 * kIn * 2;  // 49152
// One A tile, not two: the gathers of this kernel (and of the march kernel) live on the L1 that the shared-memory
// carve-out leaves, and 145 KB instead of 193 KB moves the carve-out down a step.  The accumulator stays double
// buffered in TMEM, so the epilogue of tile i - 1 still overlaps the MMAs of tile i; what is lost is only the overlap of
// the 12 MMAs of tile i (< 1 us) with the first gathers of tile i + 1 (-DSNRF_SAMB_A_TILES=2 restores it).
#ifndef SNRF_SAMB_A_TILES
#define SNRF_SAMB_A_TILES 1
#endif
constexpr uint32_t kATiles = SNRF_SAMB_A_TILES;
constexpr uint32_t kSmemBytes = kW1Bytes + kATiles * kATileBytes + 2 * 128 * 4 + 64;
constexpr int kThreads = 512;

// 12 levels x 4 (y,z) corners x 16 B for the sample this lane pair owns (identical to sam.cu's gather_f8)
template <uint32_t MASK>
__device__ __forceinline__ void gather_f8(const GridDev& G, int k0, float x, float y, float z, int xb, int row,
                                          unsigned char* a_tile) {
  const unsigned FULL = 0xffffffffu;
#pragma unroll
  for (int l = 0; l < 12; ++l) {
    const GridLevel& L = G.lv[l];
    const float qx = __fadd_rn(__fmul_rn(x, L.scale), 0.5f);
    const float qy = __fadd_rn(__fmul_rn(y, L.scale), 0.5f);
    const float qz = __fadd_rn(__fmul_rn(z, L.scale), 0.5f);
    const float fx = floorf(qx), fy = floorf(qy), fz = floorf(qz);
    const float rx = qx - fx, ry = qy - fy, rz = qz - fz;
    const uint32_t gx = static_cast<uint32_t>(static_cast<int>(fx)) + xb;
    const uint32_t gy = static_cast<uint32_t>(static_cast<int>(fy));
    const uint32_t gz = static_cast<uint32_t>(static_cast<int>(fz));
    const float wx = xb ? rx : 1.f - rx;
    uint32_t idx[4];
    uint4 v[4];
    corner_indices(L, level_hashed<MASK>(L, l), gx, gy, gz, idx);
#pragma unroll
    for (int c = 0; c < 4; ++c) v[c] = ldg_u128(G.table + 8 * static_cast<size_t>(idx[c]));
    float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float w = wx * ((c & 1) ? ry : 1.f - ry);
      w *= ((c >> 1) ? rz : 1.f - rz);
      const float2 f0 = h2_to_f2(v[c].x), f1 = h2_to_f2(v[c].y), f2 = h2_to_f2(v[c].z), f3 = h2_to_f2(v[c].w);
      a[0] += w * f0.x; a[1] += w * f0.y; a[2] += w * f1.x; a[3] += w * f1.y;
      a[4] += w * f2.x; a[5] += w * f2.y; a[6] += w * f3.x; a[7] += w * f3.y;
    }
    float o[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float send = xb ? a[i] : a[i + 4];
      const float mine = xb ? a[i + 4] : a[i];
      o[i] = mine + __shfl_xor_sync(FULL, send, 1);
    }
    const uint2 pk = make_uint2(f2_to_h2(o[0], o[1]), f2_to_h2(o[2], o[3]));
    *reinterpret_cast<uint2*>(a_tile + core_offset(row, k0 + l * 8 + xb * 4, kIn)) = pk;
  }
}

// Sum v[0..31] over groups of 2^LOG consecutive lanes; on return v[0 .. (32 >> LOG)) hold the sums of columns
// base .. base + (32 >> LOG), base = sum over steps of (lane bit `step` ? 16 >> step : 0).  LOG = 4 is sam.cu's
// halving_reduce16.
template <int LOG>
__device__ __forceinline__ int halving_reduce(float (&v)[32], int lane) {
  const unsigned FULL = 0xffffffffu;
  int base = 0;
#pragma unroll
  for (int step = 0; step < LOG; ++step) {
    const int half = 16 >> step;
    const bool up = (lane >> step) & 1;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const float send = up ? v[i] : v[i + half];
      const float keep = up ? v[i + half] : v[i];
      v[i] = keep + __shfl_xor_sync(FULL, send, 1 << step);
    }
    base += up ? half : 0;
  }
  return base;
}

// ---- epilogue of one tile out of TMEM: relu, fp16 round, slot weight, sum over the 2^LOG rows of each ray ------
template <int LOG>
__device__ __forceinline__ void epilogue_tile(const SamBucketParams& P, uint32_t tmem_base, const float* s_sw,
                                              uint64_t* s_bar, const int* list, int n_list, int64_t tile, int buf,
                                              uint32_t parity, int warp, int lane) {
  constexpr int kKeep = 32 >> LOG;  // columns a lane holds after the row reduction
  mbar_wait(smem_u32(&s_bar[buf]), parity);
  tc_fence_after();
  const int quarter = warp & 3, cq = warp >> 2;
  const int erow = quarter * 32 + lane;
  const float wgt = s_sw[buf * 128 + erow];
  const int64_t li = tile * (128 >> LOG) + (erow >> LOG);
  const bool valid = li < n_list;
  const int64_t ray = valid ? list[li] : 0;
#pragma unroll 1
  for (int chunk = 0; chunk < 2; ++chunk) {
    float v[32];
    const int col0 = cq * 64 + chunk * 32;
    tmem_ld32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + buf * 256 + col0, v);
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = round_f16(fmaxf(v[i], 0.f)) * wgt;
    const int base = halving_reduce<LOG>(v, lane);
    if (valid) {
      // kKeep consecutive fp16 columns of this ray; col0 and base are multiples of kKeep, so the widest store that
      // covers them is aligned (kKeep * 2 bytes: 64 / 32 / 16 / 8 / 4 for LOG = 0..4)
      __half* dst = P.hbar + ray * kHid + col0 + base;
      uint32_t pk[kKeep / 2];
#pragma unroll
      for (int i = 0; i < kKeep / 2; ++i) pk[i] = f2_to_h2(v[2 * i], v[2 * i + 1]);
      if (kKeep >= 8) {
#pragma unroll
        for (int i = 0; i < kKeep / 8; ++i)
          reinterpret_cast<uint4*>(dst)[i] = make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
      } else if (kKeep == 4) {
        *reinterpret_cast<uint2*>(dst) = make_uint2(pk[0], pk[1]);
      } else {
        *reinterpret_cast<uint32_t*>(dst) = pk[0];
      }
    }
  }
  tc_fence_before();
}

// One persistent launch covers all five buckets: tiles are numbered across the buckets (bucket b contributes
// ceil(count[b] / (128 >> b)) tiles of 128 >> b rays x 1 << b slots) and strided over the CTAs, so W1 is staged and
// TMEM allocated once per CTA and the buckets balance against each other.  The gather is the same code for every
// bucket (row -> (ray, slot) is a shift and a mask); only the epilogue's reduction width is a compile-time variant.
// 120 registers, not the 128 a 512-thread CTA could have: the 4096 left over are where a CTA of the push kernel
// (exchange.cu) lives while this kernel occupies the SM.
template <uint32_t M0, uint32_t M1>
__global__ void __maxnreg__(120) sam_bucket_kernel(const SamBucketParams P) {
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* s_w1 = smem;
  unsigned char* s_a = smem + kW1Bytes;
  float* s_sw = reinterpret_cast<float*>(smem + kW1Bytes + kATiles * kATileBytes);  // [2][128]
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_sw + 256);                // [2] MMA done, [1] W1 landed
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + 3);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  // W1 (96 KB, core-matrix layout in HBM) arrives through one bulk copy of the TMA unit, issued right behind the barrier
  // set-up and awaited only in front of the first MMA: it overlaps the first tile's gather instead of preceding it
  if (tid == 0) {
    mbar_init(smem_u32(&s_bar[0]), 1);
    mbar_init(smem_u32(&s_bar[1]), 1);
    mbar_init(smem_u32(&s_bar[2]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 0) tmem_alloc(smem_u32(s_tmem), 512);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  bool w1_ready = false;
  if (tid == 0) {
    mbar_expect_tx(smem_u32(&s_bar[2]), kW1Bytes);
    bulk_load_1d(smem_u32(s_w1), P.w1, kW1Bytes, smem_u32(&s_bar[2]));
  }

  // bucket sizes (written by the pre-pass on the same stream) -> first global tile of every bucket
  const int c0 = P.counts[0], c1 = P.counts[1], c2 = P.counts[2], c3 = P.counts[3], c4 = P.counts[4];
  const int64_t t1 = (c0 + 127) >> 7, t2 = t1 + ((c1 + 63) >> 6), t3 = t2 + ((c2 + 31) >> 5), t4 = t3 + ((c3 + 15) >> 4),
                t5 = t4 + ((c4 + 7) >> 3);
  auto bucket_of = [&](int64_t t, int64_t& first, int& count) -> int {
    if (t < t1) { first = 0; count = c0; return 0; }
    if (t < t2) { first = t1; count = c1; return 1; }
    if (t < t3) { first = t2; count = c2; return 2; }
    if (t < t4) { first = t3; count = c3; return 3; }
    first = t4; count = c4; return 4;
  };
  auto run_epilogue = [&](int b, int count, int64_t tile, int buf, uint32_t parity) {
    const int* list = P.lists + static_cast<int64_t>(b) * P.n_rays;
    switch (b) {
      case 0: epilogue_tile<0>(P, tmem_base, s_sw, s_bar, list, count, tile, buf, parity, warp, lane); break;
      case 1: epilogue_tile<1>(P, tmem_base, s_sw, s_bar, list, count, tile, buf, parity, warp, lane); break;
      case 2: epilogue_tile<2>(P, tmem_base, s_sw, s_bar, list, count, tile, buf, parity, warp, lane); break;
      case 3: epilogue_tile<3>(P, tmem_base, s_sw, s_bar, list, count, tile, buf, parity, warp, lane); break;
      default: epilogue_tile<4>(P, tmem_base, s_sw, s_bar, list, count, tile, buf, parity, warp, lane); break;
    }
  };

  const int g = warp & 7, e = warp >> 3;
  const int s16 = lane >> 1, xb = lane & 1;
  const int row = g * 16 + s16;  // tile row this lane pair gathers

  // Tiles are handed out by a device-wide counter (P.counts[kFeatBuckets], zeroed with the bucket sizes): a CTA whose SM
  // is shared with another kernel - the push kernel of the tile exchange runs beside this one - simply takes fewer
  // tiles instead of becoming the tail of the launch.  Thread 0 fetches the next index while the current tile is being
  // gathered; the __syncthreads in the loop body publishes it.
  __shared__ int s_next[2];
  if (tid == 0) s_next[0] = atomicAdd(P.counts + kFeatBuckets, 1);
  __syncthreads();
  int prev_b = 0, prev_count = 0;
  int64_t prev_tile = -1;
  int it = 0;
  for (int64_t t = s_next[0]; t < t5; t = s_next[it & 1]) {
    const int buf = it & 1;
    if (tid == 0) s_next[buf ^ 1] = atomicAdd(P.counts + kFeatBuckets, 1);
    unsigned char* a_tile = s_a + (kATiles == 2 ? buf : 0) * kATileBytes;
    // single A tile: the MMAs of the previous tile must have read it before this tile's rows are written (they were
    // issued before the previous iteration's epilogue ran, so this wait is normally over before it starts)
    if (kATiles == 1 && it > 0) mbar_wait(smem_u32(&s_bar[buf ^ 1]), static_cast<uint32_t>(((it - 1) >> 1) & 1));
    int64_t first;
    int count;
    const int b = bucket_of(t, first, count);  // LOG of this tile: 1 << b slots per ray, 128 >> b rays
    const int64_t tile = t - first;
    const int r_loc = row >> b, slot = row & ((1 << b) - 1);
    const int64_t li = tile * (128 >> b) + r_loc;
    const bool valid = li < count;
    // rows past the end of the list recompute the last ray with weight 0 (their results are never stored), so that
    // every lane of the warp runs the same gather and its full-mask shuffles
    const int64_t ray = (P.lists + static_cast<int64_t>(b) * P.n_rays)[valid ? li : count - 1];
    const float tm2 = P.sam_t[ray * kK + slot];
    if (e == 0 && xb == 0) s_sw[buf * 128 + row] = valid ? P.sam_w[ray * kK + slot] : 0.f;
    const float px = __fadd_rn(P.origins[3 * ray + 0], __fmul_rn(P.dirs[3 * ray + 0], tm2) / 2.f);
    const float py = __fadd_rn(P.origins[3 * ray + 1], __fmul_rn(P.dirs[3 * ray + 1], tm2) / 2.f);
    const float pz = __fadd_rn(P.origins[3 * ray + 2], __fmul_rn(P.dirs[3 * ray + 2], tm2) / 2.f);
    float x, y, z, sel;
    contract_normalize(px, py, pz, false, false, x, y, z, sel);
    if (e == 0)
      gather_f8<M0>(P.enc[0], 0, x, y, z, xb, row, a_tile);
    else
      gather_f8<M1>(P.enc[1], 96, x, y, z, xb, row, a_tile);
    fence_async_smem();
    __syncthreads();

    if (tid == 0) {
      if (!w1_ready) {
        mbar_wait(smem_u32(&s_bar[2]), 0);
        w1_ready = true;
      }
      tc_fence_after();
      const uint32_t a_addr = smem_u32(a_tile), b_addr = smem_u32(s_w1);
      const uint32_t idesc = umma_idesc_f16(128, 256);
#pragma unroll
      for (int ks = 0; ks < kIn / 16; ++ks) {
        umma_f16(tmem_base + buf * 256, umma_desc(a_addr + ks * 256, 128, kSBO), umma_desc(b_addr + ks * 256, 128, kSBO),
                 idesc, ks > 0 ? 1u : 0u);
      }
      umma_commit(smem_u32(&s_bar[buf]));
    }
    __syncwarp();
    if (it > 0) run_epilogue(prev_b, prev_count, prev_tile, buf ^ 1, static_cast<uint32_t>(((it - 1) >> 1) & 1));
    prev_b = b;
    prev_count = count;
    prev_tile = tile;
    ++it;
  }
  if (it > 0) run_epilogue(prev_b, prev_count, prev_tile, (it - 1) & 1, static_cast<uint32_t>(((it - 1) >> 1) & 1));
  if (tid == 0 && !w1_ready) mbar_wait(smem_u32(&s_bar[2]), 0);  // a CTA that got no tile still owns an in-flight W1 copy
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// One thread per ray, same rule as bucket_assign_one (kernels.cuh, the host-checkable statement of it); the list
// positions are handed out per CTA: the 256 rays of a CTA count themselves per bucket in shared memory, one thread per
// bucket reserves the CTA's range with a single global atomic, and the rays then take consecutive positions inside it.
// (One global atomic per ray on five addresses made this pre-pass 51 us per 131072 rays - 4 % of a frame.)  Rays of a
// CTA stay neighbours in the lists, so the tiles built from them keep the spatial locality of the chunk.
__global__ void __launch_bounds__(256) bucket_assign_kernel(const float* __restrict__ sam_w, float eps, int* counts,
                                                            int* lists, int64_t n, unsigned long long* totals) {
  __shared__ int s_count[kFeatBuckets], s_base[kFeatBuckets];
  if (threadIdx.x < kFeatBuckets) s_count[threadIdx.x] = 0;
  __syncthreads();
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  int b = -1, pos = 0;
  if (i < n) {
    const float4* row = reinterpret_cast<const float4*>(sam_w + i * 16);
    int k = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 w = __ldg(row + q);
      if (!(w.x < eps) && w.x != 0.f) k = 4 * q + 1;
      if (!(w.y < eps) && w.y != 0.f) k = 4 * q + 2;
      if (!(w.z < eps) && w.z != 0.f) k = 4 * q + 3;
      if (!(w.w < eps) && w.w != 0.f) k = 4 * q + 4;
    }
    b = k <= 1 ? 0 : k <= 2 ? 1 : k <= 4 ? 2 : k <= 8 ? 3 : 4;
    pos = atomicAdd(s_count + b, 1);
  }
  __syncthreads();
  if (threadIdx.x < kFeatBuckets) {
    const int c = s_count[threadIdx.x];
    s_base[threadIdx.x] = c ? atomicAdd(counts + threadIdx.x, c) : 0;
    if (totals && c) atomicAdd(totals + threadIdx.x, static_cast<unsigned long long>(c));
  }
  __syncthreads();
  if (b >= 0) lists[static_cast<int64_t>(b) * n + s_base[b] + pos] = static_cast<int>(i);
}

constexpr uint32_t kEnc0MaskStd = 0xE00u, kEnc1MaskStd = 0xFFFu;  // as in sam.cu

template <uint32_t M0, uint32_t M1>
cudaError_t launch_one(const SamBucketParams& P, int grid, cudaStream_t stream) {
  static bool configured_dev[64] = {false};
  int dev_id = 0;
  if (cudaGetDevice(&dev_id) != cudaSuccess || dev_id < 0 || dev_id >= 64) return cudaErrorInvalidDevice;
  auto* k = sam_bucket_kernel<M0, M1>;
  if (!configured_dev[dev_id]) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
    if (e != cudaSuccess) return e;
    configured_dev[dev_id] = true;
  }
  k<<<grid, kThreads, kSmemBytes, stream>>>(P);
  return cudaGetLastError();
}

}  // namespace

cudaError_t launch_bucket_assign(const float* sam_w, float eps, int* counts, int* lists, int64_t n,
                                 unsigned long long* totals, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  cudaError_t e = cudaMemsetAsync(counts, 0, (kFeatBuckets + 1) * sizeof(int), stream);  // + the tile counter of the main kernel
  if (e != cudaSuccess) return e;
  bucket_assign_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(sam_w, eps, counts, lists, n, totals);
  return cudaGetLastError();
}

// One persistent launch; the bucket sizes stay on the device (the kernel reads them), so the grid is sized for the worst
// case - every ray in the 16-slot bucket, n / 8 tiles - and capped at one CTA per SM.
cudaError_t launch_sam_bucketed(const SamBucketParams& P, int sm_count, cudaStream_t stream) {
  if (P.n_rays <= 0) return cudaSuccess;
  const bool std_cfg = hashed_mask(P.enc[0]) == kEnc0MaskStd && hashed_mask(P.enc[1]) == kEnc1MaskStd;
  const int64_t tiles = (P.n_rays + 7) / 8;
  const int grid = static_cast<int>(tiles < sm_count ? tiles : sm_count);
  return std_cfg ? launch_one<kEnc0MaskStd, kEnc1MaskStd>(P, grid, stream)
                 : launch_one<kRuntimeMask, kRuntimeMask>(P, grid, stream);
}

}  // namespace snrf
