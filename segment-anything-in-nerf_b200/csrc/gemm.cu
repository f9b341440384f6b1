// Kernel C/D - "tap GEMM":  out[M,N] = act( sum_t A_t[M,256] x W_t[N,256]^T + bias ),  fp16 operands, fp32 accumulate.
//
//   taps = 1 : the second (linear) layer of the SAM / ClipSeg MLP applied once per ray to hbar (sam.cu),
//              N = 256 or 192.               Reference: samnerf/sam_field.py:51-61,84-94 (CutlassMLP output layer).
//   taps = 9 : one 3x3 / pad-1 convolution of the patch head as an implicit GEMM over p x p = 4 x 4 ray patches
//              (16 consecutive rows = one patch, row = py*4+px); A_t is the input shifted by tap t with zero fill.
//              out_mode 1 writes fp16 rows after bias+ReLU (first conv), out_mode 2 writes the mean over each
//              patch's 16 rows (second conv + .mean(dim=[2,3])).  Reference: samnerf/sam_model.py:202-208,260-265.
//
// Persistent CTAs of 16 warps, one 128-row tile at a time, operands in shared memory in the core-matrix layout
// (common.cuh); accumulator in TMEM (tcgen05 engine, M=128 x N x K=16 x 16 per tap) or registers (legacy engine).
#include <stdlib.h>

#include "kernels.cuh"

namespace snrf {
namespace {

constexpr int kThreads = 512;
constexpr int kK = 256;
constexpr uint32_t kSBO = kK * 16;              // 4096 B between 8-row groups
constexpr uint32_t kATileBytes = 128 * kK * 2;  // 65536
constexpr uint32_t kWBytesMax = 256 * kK * 2;   // 131072
constexpr uint32_t kSmemBytes = kWBytesMax + kATileBytes + 1024 + 64;

// 16 lanes x 16 values -> lane keeps the sum of column `ret` in v[0].  15 shuffles.
__device__ __forceinline__ int halving_reduce16x16(float (&v)[16], int lane) {
  const unsigned FULL = 0xffffffffu;
  int base = 0;
#pragma unroll
  for (int step = 0; step < 4; ++step) {
    const int half = 8 >> step;
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

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- output stores of the fused tile all-gather -------------------------------------------------------------------
// One 16-byte store per destination: the local frame buffer, then either the NVSwitch multicast alias (the switch
// replicates the store into every rank's buffer) or each peer-mapped alias over NVLink.
__device__ __forceinline__ void st_mc_v4(float* p, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};\n" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void st_mc_v2(float* p, float2 v) {
  asm volatile("multimem.st.relaxed.sys.global.v2.f32 [%0], {%1,%2};\n" ::"l"(p), "f"(v.x), "f"(v.y) : "memory");
}
__device__ __forceinline__ void store_out4(const GemmParams& P, int64_t off, float4 v) {
  if (P.out_mc) {
    st_mc_v4(P.out_mc + off, v);
  } else {
    *reinterpret_cast<float4*>(P.out_f32 + off) = v;
    for (int p = 0; p < P.n_peers; ++p) *reinterpret_cast<float4*>(P.out_peer[p] + off) = v;
  }
}
__device__ __forceinline__ void store_out2(const GemmParams& P, int64_t off, float2 v) {
  if (P.out_mc) {
    st_mc_v2(P.out_mc + off, v);
  } else {
    *reinterpret_cast<float2*>(P.out_f32 + off) = v;
    for (int p = 0; p < P.n_peers; ++p) *reinterpret_cast<float2*>(P.out_peer[p] + off) = v;
  }
}

template <bool TC>
__global__ void __launch_bounds__(kThreads, 1) tapgemm_kernel(const GemmParams P) {
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* s_w = smem;
  unsigned char* s_a = smem + kWBytesMax;
  float* s_bias = reinterpret_cast<float*>(smem + kWBytesMax + kATileBytes);  // [256]
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_bias + 256);
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_bar + 1);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t w_bytes = static_cast<uint32_t>(P.n) * kK * 2;
  if (tid < 256) s_bias[tid] = (P.bias && tid < P.n) ? P.bias[tid] : 0.f;
  uint32_t tmem_base = 0;
  if (TC) {
    if (tid == 0) {
      mbar_init(smem_u32(s_bar), 1);
      asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 0) tmem_alloc(smem_u32(s_tmem), 256);
    tc_fence_before();
  }
  __syncthreads();
  if (TC) {
    tc_fence_after();
    tmem_base = *s_tmem;
  }

  const int64_t n_tiles = (P.m + 127) / 128;
  uint32_t n_commits = 0;
  int loaded_tap = -1;

  // legacy-engine accumulators: warp = (16-row group rw, column half nh)
  const int rw = warp & 7, nh = warp >> 3;
  const int nt_count = P.n / 16;  // n-tiles (of 8 columns) per column half: 16 (N=256) or 12 (N=192)
  float acc[16][4];

  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t row0 = tile * 128;
    if (!TC) {
#pragma unroll
      for (int nt = 0; nt < 16; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
    }
    for (int tap = 0; tap < P.taps; ++tap) {
      // ---- stage operands: W_tap (skipped when already resident) and the (shifted) A tile ----------
      if (loaded_tap != tap) {
        const uint4* src = reinterpret_cast<const uint4*>(P.w) + static_cast<size_t>(tap) * (w_bytes / 16);
        for (uint32_t i = tid; i < w_bytes / 16; i += kThreads) reinterpret_cast<uint4*>(s_w)[i] = ldg_u128(src + i);
        loaded_tap = tap;
      }
      const int dy = P.taps == 9 ? tap / 3 - 1 : 0, dx = P.taps == 9 ? tap % 3 - 1 : 0;
      for (int i = tid; i < 128 * 32; i += kThreads) {
        const int r = i >> 5, kc = i & 31;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        int64_t src_row = row0 + r;
        bool ok = src_row < P.m;
        if (P.taps == 9) {
          const int pix = r & 15, py = (pix >> 2) + dy, px = (pix & 3) + dx;
          ok = ok && py >= 0 && py < 4 && px >= 0 && px < 4;
          src_row = row0 + (r & ~15) + py * 4 + px;
        }
        if (ok) v = ldg_u128(P.a + src_row * kK + kc * 8);
        *reinterpret_cast<uint4*>(s_a + core_offset(r, kc * 8, kK)) = v;
      }
      if (TC) fence_async_smem();
      __syncthreads();

      if (TC) {
        if (tid == 0) {
          tc_fence_after();
          const uint32_t a_addr = smem_u32(s_a), b_addr = smem_u32(s_w);
          const uint32_t idesc = umma_idesc_f16(128, static_cast<uint32_t>(P.n));
#pragma unroll
          for (int ks = 0; ks < kK / 16; ++ks)
            umma_f16(tmem_base, umma_desc(a_addr + ks * 256, 128, kSBO), umma_desc(b_addr + ks * 256, 128, kSBO),
                     idesc, (tap > 0 || ks > 0) ? 1u : 0u);
          umma_commit(smem_u32(s_bar));
        }
        __syncwarp();
        // operands are single-buffered: wait for this tap's MMAs before anything is overwritten
        mbar_wait(smem_u32(s_bar), n_commits & 1u);
        ++n_commits;
        tc_fence_after();
      } else {
        const uint32_t a_addr = smem_u32(s_a), b_addr = smem_u32(s_w);
        const uint32_t a_row = rw * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll 1
        for (int ks = 0; ks < kK / 16; ++ks) {
          uint32_t af[4];
          ldmatrix_x4(af, a_addr + core_offset(a_row, (2 * ks + (lane >> 4)) * 8, kK));
#pragma unroll
          for (int np = 0; np < 8; ++np) {
            if (2 * np < nt_count) {
              uint32_t bf[4];
              const uint32_t n_row = nh * (P.n / 2) + (2 * np + (lane >> 4)) * 8 + (lane & 7);
              ldmatrix_x4(bf, b_addr + core_offset(n_row, (2 * ks + ((lane >> 3) & 1)) * 8, kK));
              mma_16816(acc[2 * np], af, bf[0], bf[1]);
              mma_16816(acc[2 * np + 1], af, bf[2], bf[3]);
            }
          }
        }
        __syncthreads();  // operands are reused by the next tap / tile
      }
    }

    // ---- epilogue ---------------------------------------------------------------------------------
    if (TC) {
      const int quarter = warp & 3, cg = warp >> 2;
      const int r = quarter * 32 + lane;
      const int64_t row = row0 + r;
      const int cols_per_warp = P.n / 4;  // 64 or 48
      if (P.out_mode == 0) {
        // fp32 rows leave through a shared-memory transpose (the A tile is free once the MMAs have committed):
        // TMEM hands each thread one row, but stores should be contiguous along a row - 16 B per lane, consecutive
        // lanes on consecutive 16-B pieces - so that local L2 writes are full sectors and the replicated stores of
        // the fused tile all-gather cross NVLink as 192-256 B bursts instead of 16 B packets.
        float* stage = reinterpret_cast<float*>(s_a);
        const int ld = cols_per_warp + 4;  // +4 floats: conflict-free float4 writes with one row per lane
        const int n4 = cols_per_warp / 4;
        for (int ph = 0; ph < 4; ++ph) {
          if (cg == ph) {
            for (int c = 0; c < cols_per_warp; c += 16) {
              float v[16];
              const int col0 = cg * cols_per_warp + c;
              tmem_ld16(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + col0, v);
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                v[i] += s_bias[col0 + i];
                if (P.relu) v[i] = fmaxf(v[i], 0.f);
              }
              float4* dst = reinterpret_cast<float4*>(stage + r * ld + c);
#pragma unroll
              for (int i = 0; i < 4; ++i) dst[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
            }
          }
          __syncthreads();
          for (int i = tid; i < 128 * n4; i += kThreads) {
            const int rr = i / n4, c4 = i - rr * n4;
            const int64_t grow = row0 + rr;
            if (grow < P.m)
              store_out4(P, grow * P.n + ph * cols_per_warp + c4 * 4, *reinterpret_cast<const float4*>(stage + rr * ld + c4 * 4));
          }
          __syncthreads();
        }
      } else
      for (int c = 0; c < cols_per_warp; c += 16) {
        float v[16];
        const int col0 = cg * cols_per_warp + c;
        tmem_ld16(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + col0, v);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          v[i] += s_bias[col0 + i];
          if (P.relu) v[i] = fmaxf(v[i], 0.f);
        }
        if (P.out_mode == 0) {
          if (row < P.m) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
              store_out4(P, row * P.n + col0 + 4 * i, make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]));
          }
        } else if (P.out_mode == 1) {
          if (row < P.m) {
            uint4* dst = reinterpret_cast<uint4*>(P.out_f16 + row * P.n + col0);
            dst[0] = make_uint4(f2_to_h2(v[0], v[1]), f2_to_h2(v[2], v[3]), f2_to_h2(v[4], v[5]), f2_to_h2(v[6], v[7]));
            dst[1] = make_uint4(f2_to_h2(v[8], v[9]), f2_to_h2(v[10], v[11]), f2_to_h2(v[12], v[13]),
                                f2_to_h2(v[14], v[15]));
          }
        } else {
          const int base = halving_reduce16x16(v, lane & 15);
          if (row < P.m) P.out_f32[(row >> 4) * P.n + col0 + base] = v[0] * (1.f / 16.f);
        }
      }
      tc_fence_before();
      __syncthreads();  // TMEM is overwritten by the next tile's first MMA
    } else {
      const int g = lane >> 2, q = lane & 3;
      const unsigned FULL = 0xffffffffu;
      const int64_t r_lo = row0 + rw * 16 + g, r_hi = r_lo + 8;
#pragma unroll
      for (int nt = 0; nt < 16; ++nt) {
        if (nt < nt_count) {
          const int col = nh * (P.n / 2) + nt * 8 + 2 * q;
          float c0 = acc[nt][0] + s_bias[col], c1 = acc[nt][1] + s_bias[col + 1];
          float c2 = acc[nt][2] + s_bias[col], c3 = acc[nt][3] + s_bias[col + 1];
          if (P.relu) {
            c0 = fmaxf(c0, 0.f); c1 = fmaxf(c1, 0.f); c2 = fmaxf(c2, 0.f); c3 = fmaxf(c3, 0.f);
          }
          if (P.out_mode == 0) {
            if (r_lo < P.m) store_out2(P, r_lo * P.n + col, make_float2(c0, c1));
            if (r_hi < P.m) store_out2(P, r_hi * P.n + col, make_float2(c2, c3));
          } else if (P.out_mode == 1) {
            if (r_lo < P.m) *reinterpret_cast<uint32_t*>(P.out_f16 + r_lo * P.n + col) = f2_to_h2(c0, c1);
            if (r_hi < P.m) *reinterpret_cast<uint32_t*>(P.out_f16 + r_hi * P.n + col) = f2_to_h2(c2, c3);
          } else {
            float s0 = c0 + c2, s1 = c1 + c3;
#pragma unroll
            for (int o = 4; o < 32; o <<= 1) {
              s0 += __shfl_xor_sync(FULL, s0, o);
              s1 += __shfl_xor_sync(FULL, s1, o);
            }
            if (g == 0 && r_lo < P.m)
              *reinterpret_cast<float2*>(P.out_f32 + (r_lo >> 4) * P.n + col) = make_float2(s0 / 16.f, s1 / 16.f);
          }
        }
      }
    }
  }
  if (TC) {
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 256);
  }
}

}  // namespace

cudaError_t launch_tapgemm(const GemmParams& P, bool tcgen05, int sm_count, cudaStream_t stream) {
  // the plain output layer runs on the TMA-fed kernel (gemm_tma.cu); SNRF_TAPGEMM=v1 keeps it on this file's kernel (A/B runs)
  static const bool force_v1 = [] { const char* e = getenv("SNRF_TAPGEMM"); return e && e[0] == 'v' && e[1] == '1'; }();
  if (tcgen05 && !force_v1) {
    const cudaError_t e = launch_tapgemm_tma(P, sm_count, stream);
    if (e != cudaErrorNotSupported) return e;
  }
  // function attributes are per device: remember which devices of this process have been configured
  static bool configured_dev[64] = {false};
  int dev_id = 0;
  if (cudaGetDevice(&dev_id) != cudaSuccess || dev_id < 0 || dev_id >= 64) return cudaErrorInvalidDevice;
  bool& configured = configured_dev[dev_id];
  if (!configured) {
    cudaError_t e =
        cudaFuncSetAttribute(tapgemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(tapgemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  if (P.m <= 0) return cudaSuccess;
  if ((P.n != 256 && P.n != 192) || (P.taps != 1 && P.taps != 9)) return cudaErrorInvalidValue;
  const int64_t n_tiles = (P.m + 127) / 128;
  const int grid = static_cast<int>(n_tiles < sm_count ? n_tiles : sm_count);
  if (tcgen05)
    tapgemm_kernel<true><<<grid, kThreads, kSmemBytes, stream>>>(P);
  else
    tapgemm_kernel<false><<<grid, kThreads, kSmemBytes, stream>>>(P);
  return cudaGetLastError();
}

}  // namespace snrf
