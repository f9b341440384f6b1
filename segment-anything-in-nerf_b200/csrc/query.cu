// Small stand-alone kernels behind the component-level shims (Field.density_fn, SAMField.get_outputs,
// RaySamples.get_weights and the renderers when they are called outside the fused render), plus the
// parameter packing kernels.  None of these is on the headline path; they are written for clarity.
//
// Reference semantics: nerfstudio/fields/base_field.py:38-56,99-118; nerfstudio/cameras/rays.py:141-163;
// nerfstudio/model_components/renderers.py:69-140,197-223,260-270; samnerf/sam_model.py:126-137.
#include "kernels.cuh"

namespace snrf {

// ---------------------------------------------------------------------------------------------
// hash-grid encoding of arbitrary points: one thread per (point, level)
// ---------------------------------------------------------------------------------------------
template <int F>
__global__ void encode_kernel(const QueryParams P, int gi, int col0, int width) {
  const GridDev& G = P.grid[gi];
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t n_items = P.n * G.n_levels;
  if (i >= n_items) return;
  const int64_t pt = i / G.n_levels;
  const int l = static_cast<int>(i % G.n_levels);
  float x, y, z, sel;
  contract_normalize(P.xyz[3 * pt], P.xyz[3 * pt + 1], P.xyz[3 * pt + 2], P.linf != 0, P.selector != 0, x, y, z, sel);
  if (P.sel && l == 0 && gi == 0) P.sel[pt] = sel;
  const GridLevel L = G.lv[l];
  const float qx = __fadd_rn(__fmul_rn(x, L.scale), 0.5f), qy = __fadd_rn(__fmul_rn(y, L.scale), 0.5f),
              qz = __fadd_rn(__fmul_rn(z, L.scale), 0.5f);
  const float fx = floorf(qx), fy = floorf(qy), fz = floorf(qz);
  const float rx = qx - fx, ry = qy - fy, rz = qz - fz;
  const uint32_t gx = static_cast<uint32_t>(static_cast<int>(fx)), gy = static_cast<uint32_t>(static_cast<int>(fy)),
                 gz = static_cast<uint32_t>(static_cast<int>(fz));
  float acc[F];
#pragma unroll
  for (int f = 0; f < F; ++f) acc[f] = 0.f;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    float w = (c & 1) ? rx : 1.f - rx;
    w *= (c & 2) ? ry : 1.f - ry;
    w *= (c & 4) ? rz : 1.f - rz;
    const uint32_t idx = grid_index(L, gx + (c & 1), gy + ((c >> 1) & 1), gz + (c >> 2));
    const __half* e = G.table + static_cast<size_t>(idx) * F;
#pragma unroll
    for (int f = 0; f < F; ++f) acc[f] += w * __half2float(e[f]);
  }
  __half* o = P.feat + pt * width + col0 + l * F;
#pragma unroll
  for (int f = 0; f < F; ++f) o[f] = __float2half_rn(acc[f]);
}

cudaError_t launch_encode(const QueryParams& P, cudaStream_t stream) {
  int width = 0;
  for (int g = 0; g < P.n_grids; ++g) width += P.grid[g].n_levels * P.grid[g].n_features;
  int col0 = 0;
  for (int g = 0; g < P.n_grids; ++g) {
    const int64_t items = P.n * P.grid[g].n_levels;
    if (items > 0) {
      const int blocks = static_cast<int>((items + 255) / 256);
      if (P.grid[g].n_features == 2)
        encode_kernel<2><<<blocks, 256, 0, stream>>>(P, g, col0, width);
      else if (P.grid[g].n_features == 8)
        encode_kernel<8><<<blocks, 256, 0, stream>>>(P, g, col0, width);
      else
        return cudaErrorInvalidValue;
    }
    col0 += P.grid[g].n_levels * P.grid[g].n_features;
  }
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// one dense layer, fp16 in / fp32 accumulate / fp16 out: out[n][j] = act(sum_i in[n][i] * W[j][i])
// `k_in` input columns are read from `in` (row stride ld_in); columns k_in..k_w-1 of W multiply `pad_value`
// (tcnn pads grid encodings with 0 and identity inputs with 1).  act: 0 none, 1 ReLU, 2 sigmoid.
// ---------------------------------------------------------------------------------------------
__global__ void dense_kernel(const __half* __restrict__ in, int ld_in, int k_in, const __half* __restrict__ w, int k_w,
                             float pad_value, __half* __restrict__ out, int ld_out, int n_out, int act, int64_t n) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n * n_out) return;
  const int64_t r = i / n_out;
  const int j = static_cast<int>(i % n_out);
  const __half* x = in + r * ld_in;
  const __half* wr = w + static_cast<size_t>(j) * k_w;
  float a = 0.f;
  for (int k = 0; k < k_in; ++k) a += __half2float(x[k]) * __half2float(wr[k]);
  for (int k = k_in; k < k_w; ++k) a += pad_value * __half2float(wr[k]);
  if (act == 1) a = fmaxf(a, 0.f);
  a = round_f16(a);
  if (act == 2) a = 1.f / (1.f + expf(-a));
  out[r * ld_out + j] = __float2half_rn(a);
}

cudaError_t launch_dense(const __half* in, int ld_in, int k_in, const __half* w, int k_w, float pad_value, __half* out,
                         int ld_out, int n_out, int act, int64_t n, cudaStream_t stream) {
  const int64_t items = n * n_out;
  if (items <= 0) return cudaSuccess;
  dense_kernel<<<static_cast<int>((items + 255) / 256), 256, 0, stream>>>(in, ld_in, k_in, w, k_w, pad_value, out,
                                                                          ld_out, n_out, act, n);
  return cudaGetLastError();
}

// density = exp(h[:,0]) * sel ; geo = h[:,1:1+n_geo]  (h fp16 [n, ld])
__global__ void density_finish_kernel(const __half* h, int ld, const float* sel, float* density, __half* geo, int n_geo,
                                      int64_t n) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  density[i] = expf(__half2float(h[i * ld])) * (sel ? sel[i] : 1.f);
  if (geo)
    for (int k = 0; k < n_geo; ++k) geo[i * n_geo + k] = h[i * ld + 1 + k];
}
cudaError_t launch_density_finish(const __half* h, int ld, const float* sel, float* density, __half* geo, int n_geo,
                                  int64_t n, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  density_finish_kernel<<<static_cast<int>((n + 255) / 256), 256, 0, stream>>>(h, ld, sel, density, geo, n_geo, n);
  return cudaGetLastError();
}

// colour-head input rows: [SH16(dir), geo(15)] fp16 (pad column handled by dense_kernel)
__global__ void head_input_kernel(const float* dirs, const __half* geo, __half* x, int64_t n) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float sx = ((dirs[3 * i] + 1.f) / 2.f) * 2.f - 1.f, sy = ((dirs[3 * i + 1] + 1.f) / 2.f) * 2.f - 1.f,
              sz = ((dirs[3 * i + 2] + 1.f) / 2.f) * 2.f - 1.f;
  const float xy = sx * sy, xz = sx * sz, yz = sy * sz, x2 = sx * sx, y2 = sy * sy, z2 = sz * sz;
  float sh[16];
  sh[0] = 0.28209479177387814f;
  sh[1] = -0.48860251190291987f * sy;
  sh[2] = 0.48860251190291987f * sz;
  sh[3] = -0.48860251190291987f * sx;
  sh[4] = 1.0925484305920792f * xy;
  sh[5] = -1.0925484305920792f * yz;
  sh[6] = 0.94617469575755997f * z2 - 0.31539156525251999f;
  sh[7] = -1.0925484305920792f * xz;
  sh[8] = 0.54627421529603959f * x2 - 0.54627421529603959f * y2;
  sh[9] = 0.59004358992664352f * sy * (-3.f * x2 + y2);
  sh[10] = 2.8906114426405538f * xy * sz;
  sh[11] = 0.45704579946446572f * sy * (1.f - 5.f * z2);
  sh[12] = 0.3731763325901154f * sz * (5.f * z2 - 3.f);
  sh[13] = 0.45704579946446572f * sx * (1.f - 5.f * z2);
  sh[14] = 1.4453057213202769f * sz * (x2 - y2);
  sh[15] = 0.59004358992664352f * sx * (-x2 + 3.f * y2);
  for (int k = 0; k < 16; ++k) x[i * 32 + k] = __float2half_rn(sh[k]);
  for (int k = 0; k < 15; ++k) x[i * 32 + 16 + k] = geo[i * 15 + k];
  x[i * 32 + 31] = __float2half_rn(1.f);
}
cudaError_t launch_head_input(const float* dirs, const __half* geo, __half* x, int64_t n, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  head_input_kernel<<<static_cast<int>((n + 255) / 256), 256, 0, stream>>>(dirs, geo, x, n);
  return cudaGetLastError();
}

__global__ void half_to_float_kernel(const __half* in, int ld_in, float* out, int ld_out, int cols, int64_t n) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n * cols) return;
  out[(i / cols) * ld_out + i % cols] = __half2float(in[(i / cols) * ld_in + i % cols]);
}
cudaError_t launch_half_to_float(const __half* in, int ld_in, float* out, int ld_out, int cols, int64_t n,
                                 cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  half_to_float_kernel<<<static_cast<int>((n * cols + 255) / 256), 256, 0, stream>>>(in, ld_in, out, ld_out, cols, n);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// renderer-level shims: one warp per ray, S samples strided over the lanes
// ---------------------------------------------------------------------------------------------
// mode 0: weights = get_weights(deltas, densities)            out[N,S]
// mode 1: accumulation = sum(weights)                         out[N]
// mode 2: median depth(weights, starts, ends)                 out[N]
// mode 3: rgb composite(rgb[N,S,3], weights, bg)              out[N,3]
// mode 4: mean renderer(embeds[N,S,C], weights)               out[N,C]
__global__ void ray_ops_kernel(int mode, const float* a, const float* b, const float* c, float* out, int64_t n, int S,
                               int C, int bg_mode, float bg0, float bg1, float bg2) {
  const int lane = threadIdx.x & 31;
  const int64_t ray = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (ray >= n) return;
  const unsigned FULL = 0xffffffffu;
  if (mode == 0) {
    float carry = 0.f;
    for (int s0 = 0; s0 < S; s0 += 32) {
      const int s = s0 + lane;
      const float ds = s < S ? a[ray * S + s] * b[ray * S + s] : 0.f;
      const float incl = warp_incl_scan(ds, lane);
      float excl = __shfl_up_sync(FULL, incl, 1);
      if (lane == 0) excl = 0.f;
      if (s < S) out[ray * S + s] = nan_to_num((1.f - expf(-ds)) * expf(-(carry + excl)));
      carry += __shfl_sync(FULL, incl, 31);
    }
  } else if (mode == 1) {
    float t = 0.f;
    for (int s = lane; s < S; s += 32) t += a[ray * S + s];
    t = warp_sum(t);
    if (lane == 0) out[ray] = t;
  } else if (mode == 2) {
    float carry = 0.f;
    int found = -1;
    for (int s0 = 0; s0 < S && found < 0; s0 += 32) {
      const int s = s0 + lane;
      const float w = s < S ? a[ray * S + s] : 0.f;
      const float cum = carry + warp_incl_scan(w, lane);
      const unsigned m = __ballot_sync(FULL, s < S && cum >= 0.5f);
      if (m) found = s0 + __ffs(m) - 1;
      carry = __shfl_sync(FULL, cum, 31);
    }
    if (found < 0) found = S - 1;
    if (lane == 0) out[ray] = (b[ray * S + found] + c[ray * S + found]) / 2.f;
  } else if (mode == 3) {
    float sr = 0.f, sg = 0.f, sb = 0.f, sw = 0.f;
    for (int s = lane; s < S; s += 32) {
      const float w = b[ray * S + s];
      sr += w * nan_to_num(a[(ray * S + s) * 3 + 0]);
      sg += w * nan_to_num(a[(ray * S + s) * 3 + 1]);
      sb += w * nan_to_num(a[(ray * S + s) * 3 + 2]);
      sw += w;
    }
    sr = warp_sum(sr); sg = warp_sum(sg); sb = warp_sum(sb); sw = warp_sum(sw);
    if (lane == 0) {
      if (bg_mode == kBgLastSample) {
        bg0 = nan_to_num(a[(ray * S + S - 1) * 3 + 0]);
        bg1 = nan_to_num(a[(ray * S + S - 1) * 3 + 1]);
        bg2 = nan_to_num(a[(ray * S + S - 1) * 3 + 2]);
      }
      out[ray * 3 + 0] = fminf(fmaxf(sr + bg0 * (1.f - sw), 0.f), 1.f);
      out[ray * 3 + 1] = fminf(fmaxf(sg + bg1 * (1.f - sw), 0.f), 1.f);
      out[ray * 3 + 2] = fminf(fmaxf(sb + bg2 * (1.f - sw), 0.f), 1.f);
    }
  } else if (mode == 4) {
    for (int ch = lane; ch < C; ch += 32) {
      float t = 0.f;
      for (int s = 0; s < S; ++s) t += b[ray * S + s] * a[(ray * S + s) * C + ch];
      out[ray * C + ch] = t;
    }
  }
}
cudaError_t launch_ray_ops(int mode, const float* a, const float* b, const float* c, float* out, int64_t n, int S, int C,
                           int bg_mode, const float* bg, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  const int64_t threads = n * 32;
  ray_ops_kernel<<<static_cast<int>((threads + 255) / 256), 256, 0, stream>>>(mode, a, b, c, out, n, S, C, bg_mode,
                                                                             bg ? bg[0] : 0.f, bg ? bg[1] : 0.f,
                                                                             bg ? bg[2] : 0.f);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// parameter packing (runs once per upload)
// ---------------------------------------------------------------------------------------------
__global__ void f32_to_f16_kernel(const float* in, __half* out, int64_t n) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __float2half_rn(in[i]);
}
cudaError_t launch_f32_to_f16(const float* in, __half* out, int64_t n, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  f32_to_f16_kernel<<<static_cast<int>((n + 255) / 256), 256, 0, stream>>>(in, out, n);
  return cudaGetLastError();
}

// row-major fp32 W[rows, cols] -> fp16 core-matrix layout (common.cuh::core_offset)
// (column k of the packed operand reads source column perm[k]; perm == null: identity)
__global__ void pack_core_kernel(const float* w, __half* out, int rows, int cols, const int* perm) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * cols) return;
  const int r = i / cols, k = i % cols;
  out[core_offset(r, k, cols) / 2] = __float2half_rn(w[r * cols + (perm ? perm[k] : k)]);
}
cudaError_t launch_pack_core(const float* w, __half* out, int rows, int cols, cudaStream_t stream, const int* perm) {
  pack_core_kernel<<<(rows * cols + 255) / 256, 256, 0, stream>>>(w, out, rows, cols, perm);
  return cudaGetLastError();
}

// row-major fp32 W[n_out, k_w] -> mma.sync B-fragment tiles.  Column c of the packed operand reads source
// column perm[c] (perm == null: identity).  tile (nt, ks), lane (g,q): {W[8nt+g][16ks+2q..+1], W[..][16ks+2q+8..+9]}
__global__ void pack_frag_kernel(const float* w, int k_w, const int* perm, uint2* out, int n_tiles, int k_steps) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_tiles * k_steps * 32) return;
  const int lane = i & 31, t = i >> 5;
  const int nt = t / k_steps, ks = t % k_steps;
  const int g = lane >> 2, q = lane & 3;
  const float* row = w + static_cast<size_t>(nt * 8 + g) * k_w;
  auto col = [&](int c) { return row[perm ? perm[c] : c]; };
  const int c0 = ks * 16 + 2 * q;
  out[i] = make_uint2(f2_to_h2(col(c0), col(c0 + 1)), f2_to_h2(col(c0 + 8), col(c0 + 9)));
}
cudaError_t launch_pack_frag(const float* w, int k_w, const int* perm, uint2* out, int n_tiles, int k_steps,
                             cudaStream_t stream) {
  const int n = n_tiles * k_steps * 32;
  pack_frag_kernel<<<(n + 255) / 256, 256, 0, stream>>>(w, k_w, perm, out, n_tiles, k_steps);
  return cudaGetLastError();
}

// fp32 (fp16-rounded) copy of the proposal MLP: w1 -> [16][17], density row of w2 -> [16]
__global__ void pack_prop_kernel(const float* w1, int k_w, const float* w2, float* o1, float* o2) {
  const int i = threadIdx.x;
  if (i < 256) o1[(i >> 4) * 17 + (i & 15)] = round_f16(w1[(i >> 4) * k_w + (i & 15)]);
  if (i < 16) o2[i] = round_f16(w2[i]);
}
cudaError_t launch_pack_prop(const float* w1, int k_w, const float* w2, float* o1, float* o2, cudaStream_t stream) {
  pack_prop_kernel<<<1, 256, 0, stream>>>(w1, k_w, w2, o1, o2);
  return cudaGetLastError();
}

}  // namespace snrf
