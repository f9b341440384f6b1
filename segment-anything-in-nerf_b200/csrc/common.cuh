// Device-side building blocks shared by every kernel of the SAM-NeRF render path (sm_100a only).
//
// Semantics restated from the reference (paths relative to /root/reference) and from the published
// tiny-cuda-nn algorithm it calls:
//   * scene contraction        nerfstudio/field_components/spatial_distortions.py:66-88
//   * (p+2)/4 + (0,1) selector  nerfstudio/fields/density_fields.py:102-112, nerfacto_field.py:244-253
//   * piecewise spacing        nerfstudio/model_components/ray_samplers.py:242-243
//   * hash-grid indexing       tcnn grid.h (cross-check: nerfstudio/field_components/cuda/csrc/temporal_gridencoder.cu:46-88)
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

// Pure per-thread helpers are host+device so that tests/emu can run the bodies of the simple (one thread = one item)
// kernels on the CPU against the oracle; device code generation is unaffected.
#define SNRF_HD __host__ __device__ __forceinline__

namespace snrf {

// a*b and a*b+c WITHOUT fma contraction, so that device, host emulation and the torch oracle round alike
SNRF_HD float mul_rn(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fmul_rn(a, b);
#else
  volatile float m = a * b;
  return m;
#endif
}
SNRF_HD float mul_add_rn(float a, float b, float c) {
#ifdef __CUDA_ARCH__
  return __fadd_rn(__fmul_rn(a, b), c);
#else
  volatile float m = a * b;
  return m + c;
#endif
}
// Frustums.get_positions for a sample stored as tm2 = start + end: o + d * (start + end) / 2 (rays.py:48-57),
// op for op what march.cu / sam.cu compute
SNRF_HD float sample_coord(float o, float d, float tm2) {
#ifdef __CUDA_ARCH__
  return __fadd_rn(o, __fmul_rn(d, tm2) / 2.f);
#else
  volatile float m = d * tm2;
  volatile float h = m / 2.f;
  return o + h;
#endif
}

constexpr int kMaxLevels = 16;
constexpr uint32_t kPrimeY = 2654435761u;
constexpr uint32_t kPrimeZ = 805459861u;

struct GridLevel {
  float scale;      // exp2f(l*log2f(per_level_scale))*base - 1
  uint32_t res;     // ceil(scale)+1
  uint32_t size;    // entries in this level
  uint32_t offset;  // first entry of this level in the table
  uint32_t hashed;  // 1: coherent-prime hash, 0: dense x + y*res + z*res^2
};

// One multi-resolution table in HBM: fp16, level-major, F halfs per entry (tcnn layout).
struct GridDev {
  const __half* table;
  int n_levels;
  int n_features;
  GridLevel lv[kMaxLevels];
};

// ---------------------------------------------------------------------------------------------
// scalar helpers (op order mirrors the torch expressions the oracle evaluates)
// ---------------------------------------------------------------------------------------------
SNRF_HD float round_f16(float x) { return __half2float(__float2half_rn(x)); }

SNRF_HD float spacing_fn(float x) { return x < 1.f ? x / 2.f : 1.f - 1.f / (2.f * x); }
SNRF_HD float spacing_fn_inv(float x) { return x < 0.5f ? 2.f * x : 1.f / (2.f - 2.f * x); }

// Spacing-space bin edge j of n_bins + 1 under training-mode single jitter (ray_samplers.py:104-112):
// bins = linspace(0, 1, n_bins + 1); bin_lower + (bin_upper - bin_lower) * t_rand with lower / upper = the neighbouring
// bin centres (2j-1)/(2 n_bins), (2j+1)/(2 n_bins) and the ends 0, 1 at j = 0, n_bins.  n_bins is a power of two here,
// so every term but the final multiply-add is exact in fp32 and this is the torch expression bit for bit.
SNRF_HD float jittered_bin(int j, int n_bins, float t_rand) {
  const float lower = j == 0 ? 0.f : static_cast<float>(2 * j - 1) * (1.f / static_cast<float>(2 * n_bins));
  const float upper = j == n_bins ? 1.f : static_cast<float>(2 * j + 1) * (1.f / static_cast<float>(2 * n_bins));
  return mul_add_rn(upper - lower, t_rand, lower);
}

// torch.nan_to_num defaults: nan -> 0, +inf -> FLT_MAX, -inf -> -FLT_MAX
SNRF_HD float nan_to_num(float x) {
  if (x != x) return 0.f;
  if (x == INFINITY) return 3.402823466e+38f;
  if (x == -INFINITY) return -3.402823466e+38f;
  return x;
}

// contraction followed by (p+2)/4. `linf`: order=inf (density fields) else L2 (SAMField).
// `use_selector`: density fields zero positions outside (0,1) and gate the density with it.
SNRF_HD void contract_normalize(float px, float py, float pz, bool linf, bool use_selector,
                                                   float& x, float& y, float& z, float& sel) {
  float mag;
  if (linf) {
    mag = fmaxf(fabsf(px), fmaxf(fabsf(py), fabsf(pz)));
  } else {
    mag = sqrtf(px * px + py * py + pz * pz);
  }
  if (!(mag < 1.f)) {
    const float s = 2.f - (1.f / mag);
    px = s * (px / mag);
    py = s * (py / mag);
    pz = s * (pz / mag);
  }
  x = (px + 2.f) / 4.f;
  y = (py + 2.f) / 4.f;
  z = (pz + 2.f) / 4.f;
  sel = 1.f;
  if (use_selector) {
    const bool in = (x > 0.f) && (x < 1.f) && (y > 0.f) && (y < 1.f) && (z > 0.f) && (z < 1.f);
    sel = in ? 1.f : 0.f;
    x *= sel;
    y *= sel;
    z *= sel;
  }
}

// Hashed levels always have a power-of-two size (2^log2_hashmap_size; checked at upload), so `% size` is a mask.
// Dense levels: idx <= res + res^2 + res^3 < 2*size, and idx >= size only for the +1 corner of a coordinate on the
// upper boundary, so `% size` is one conditional subtract; the final min() keeps non-finite inputs in bounds.
SNRF_HD uint32_t wrap_dense(uint32_t idx, uint32_t size) {
  return min(idx >= size ? idx - size : idx, size - 1u);
}

SNRF_HD uint32_t grid_index(const GridLevel& L, uint32_t gx, uint32_t gy, uint32_t gz) {
  if (L.hashed) return ((gx ^ (gy * kPrimeY) ^ (gz * kPrimeZ)) & (L.size - 1u)) + L.offset;
  return wrap_dense(gx + gy * L.res + gz * L.res * L.res, L.size) + L.offset;
}

// Level-type masks: bit l set = level l is hashed.  Kernels are instantiated for the reference configs' masks so
// that the dense / hashed choice is resolved at compile time (no branches between the loads of different levels);
// any other configuration runs the kRuntimeMask instantiation, which reads GridLevel::hashed.
constexpr uint32_t kRuntimeMask = 0x80000000u;
template <uint32_t MASK>
SNRF_HD bool level_hashed(const GridLevel& L, int l) {
  return (MASK & kRuntimeMask) ? (L.hashed != 0u) : (((MASK >> l) & 1u) != 0u);
}

// entry indices of the four (y,z) corners c = dy + 2*dz of a voxel for a fixed x coordinate
SNRF_HD void corner_indices(const GridLevel& L, bool hashed, uint32_t gx, uint32_t gy, uint32_t gz,
                                               uint32_t (&idx)[4]) {
  if (hashed) {
    const uint32_t m = L.size - 1u;
    const uint32_t hy0 = gy * kPrimeY, hy1 = hy0 + kPrimeY;
    const uint32_t hz0 = gz * kPrimeZ, hz1 = hz0 + kPrimeZ;
    idx[0] = ((gx ^ hy0 ^ hz0) & m) + L.offset;
    idx[1] = ((gx ^ hy1 ^ hz0) & m) + L.offset;
    idx[2] = ((gx ^ hy0 ^ hz1) & m) + L.offset;
    idx[3] = ((gx ^ hy1 ^ hz1) & m) + L.offset;
  } else {
    const uint32_t r = L.res, r2 = r * r;
    const uint32_t b = gx + gy * r + gz * r2;
    idx[0] = wrap_dense(b, L.size) + L.offset;
    idx[1] = wrap_dense(b + r, L.size) + L.offset;
    idx[2] = wrap_dense(b + r2, L.size) + L.offset;
    idx[3] = wrap_dense(b + r + r2, L.size) + L.offset;
  }
}
inline uint32_t hashed_mask(const GridDev& g) {
  uint32_t m = 0;
  for (int l = 0; l < g.n_levels; ++l) m |= (g.lv[l].hashed ? 1u : 0u) << l;
  return m;
}

// ---------------------------------------------------------------------------------------------
// memory helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t ldg_u32(const void* p) { return __ldg(reinterpret_cast<const uint32_t*>(p)); }
__device__ __forceinline__ uint4 ldg_u128(const void* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }

__device__ __forceinline__ float2 h2_to_f2(uint32_t v) {
  return __half22float2(*reinterpret_cast<const __half2*>(&v));
}
__device__ __forceinline__ uint32_t f2_to_h2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---------------------------------------------------------------------------------------------
// warp-level tensor-core tile (legacy mma.sync path: HMMA.16816.F32)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mma_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(saddr));
}
__device__ __forceinline__ void ldmatrix_x2(uint32_t& r0, uint32_t& r1, uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];\n" : "=r"(r0), "=r"(r1) : "r"(saddr));
}

// ---------------------------------------------------------------------------------------------
// "core-matrix" shared-memory layout used for every tensor-core operand (A tiles and weights):
// element (row, k) of a [rows x K] fp16 K-major operand lives at byte
//     (row/8)*SBO + (k/8)*128 + (row%8)*16 + (k%8)*2,      SBO = (K/8)*128.
// One 8x8 core matrix is 128 contiguous bytes, so the same bytes serve as
//   * a tcgen05 shared-memory descriptor with SWIZZLE_NONE, K-major, LBO = 128 B, SBO = K*16 B, and
//   * an ldmatrix source (8 row addresses 16 B apart: conflict-free) for the mma.sync path.
// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t core_offset(uint32_t row, uint32_t k, uint32_t K) {
  return (row >> 3) * (K * 16u) + (k >> 3) * 128u + (row & 7u) * 16u + (k & 7u) * 2u;
}

// ---------------------------------------------------------------------------------------------
// tcgen05 (5th-gen tensor core) primitives
// ---------------------------------------------------------------------------------------------
// shared-memory matrix descriptor, no swizzle, K-major (cute::UMMA::SmemDescriptor bit layout):
// [0,14) addr>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 | [61,64) layout=0
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr >> 4) & 0x3FFFu);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;
  return d;
}
// instruction descriptor for kind::f16, A/B fp16 K-major, D fp32 (cute::UMMA::InstrDescriptor):
// [4,6) c_format=1(F32) | [7,10) a_format=0(F16) | [10,13) b_format=0 | [17,23) N>>3 | [24,29) M>>4
__host__ __device__ constexpr uint32_t umma_idesc_f16(uint32_t M, uint32_t N) {
  return (1u << 4) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_dst), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t mbar_saddr) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(mbar_saddr)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

__device__ __forceinline__ void mbar_init(uint32_t saddr, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(saddr), "r"(count) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t saddr, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r"(ok)
      : "r"(saddr), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t saddr, uint32_t parity) {
  while (!mbar_try_wait(saddr, parity)) {
  }
}
// transaction barriers + 1-D bulk copies (TMA unit, SASS UBLKCP): global -> shared without passing through registers
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread t of the warp receives row (lane base + t).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,"
      "%30,%31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// ---------------------------------------------------------------------------------------------
// warp reductions
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_incl_scan(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float n = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += n;
  }
  return v;
}

}  // namespace snrf
