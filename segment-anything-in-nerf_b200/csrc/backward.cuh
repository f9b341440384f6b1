// Backward pass of the feature-field branch (first slice): gradients of
//     out[r] = W2 . sum_k w[r,k] * fp16(relu(W1 x[r,k])),     x[r,k] = concat(enc0, enc1)(contract_L2(pos[r,k]))
// with respect to W1, W2 and the two hash tables, given d_out.  Positions and weights carry no gradient:
// SAMField detaches the positions (samnerf/sam_field.py:116) and MeanRenderer is fed sam_weights.detach()
// (samnerf/sam_model.py:258-277), so this branch is the whole backward of the `sam_field` parameter group
// (samnerf/sam_model.py:330-335).  What tinycudann's autograd does for the reference - CutlassMLP backward
// (dL/dW = dY^T X, dL/dX = dY W) and the grid backward that scatters dL/dx * trilinear weight into the 8 corners
// with atomicAdd (in-tree statement of the same scatter: nerfstudio/field_components/cuda/csrc/
// temporal_gridencoder.cu:283-370) - restated here in fp32 (tcnn: fp16 with loss scaling; fp32 is the more exact
// reading, and the gradients land directly in the flat fp32 layout of `params.grad`).
//
// First version: every kernel is "one thread = one output item", no shared memory, no warp intrinsics, so that the
// bodies below run unchanged on the CPU in tests/emu against torch autograd through the oracle.  Rounding points of
// the forward pass (fp16 tables / encoder outputs / hidden activations) pass gradients straight through.
#pragma once
#include "common.cuh"

namespace snrf {

constexpr int kBwdK = 16;     // picked samples per ray
constexpr int kBwdIn = 192;   // encoder width (2 x 12 levels x 8 features)
constexpr int kBwdHid = 256;  // hidden width

SNRF_HD void atomic_add_f32(float* p, float v) {
#ifdef __CUDA_ARCH__
  atomicAdd(p, v);
#else
  *p += v;  // host emulation is single-threaded
#endif
}

struct FeatBwdParams {
  // inputs
  const float* origins;   // [N,3]
  const float* dirs;      // [N,3]
  const float* sam_t;     // [N,16] 2 x midpoint t of the picked samples
  const float* sam_w;     // [N,16] sharpened, renormalised weights
  const float* d_out;     // [N,n_out]
  const __half* x;        // [N,16,192] encoder outputs saved by the forward pass
  const __half* w1;       // [256,192] row-major fp16
  const __half* w2;       // [n_out,256] row-major fp16
  int64_t n_rays;
  int n_out;
  GridDev enc[2];         // geometry only (tables are not read)
  // scratch (fp32)
  float* d_hbar;          // [N,256]
  float* hbar;            // [N,256]
  float* dh;              // [N*16,256]
  float* dx;              // [N*16,192]
  // outputs, accumulated (+=)
  float* g_w1;            // [256,192]
  float* g_w2;            // [n_out,256]
  float* g_table[2];      // [entries*8] each
};

// K1: d_hbar[r,j] = sum_o d_out[r,o] * W2[o,j]                                     item = (ray, j)
SNRF_HD void bwd_dhbar_one(const FeatBwdParams& P, int64_t item) {
  const int64_t r = item / kBwdHid;
  const int j = static_cast<int>(item % kBwdHid);
  const float* g = P.d_out + r * P.n_out;
  float a = 0.f;
  for (int o = 0; o < P.n_out; ++o) a += g[o] * __half2float(P.w2[static_cast<size_t>(o) * kBwdHid + j]);
  P.d_hbar[item] = a;
}

// K2: recompute h[r,k,j] = W1[j,:] . x[r,k,:]; hbar[r,j] = fp16(sum_k w_k * fp16(relu(h)));
//     dh[r,k,j] = h > 0 ? w_k * d_hbar[r,j] : 0                                   item = (ray, j)
SNRF_HD void bwd_hidden_one(const FeatBwdParams& P, int64_t item) {
  const int64_t r = item / kBwdHid;
  const int j = static_cast<int>(item % kBwdHid);
  const __half* wr = P.w1 + static_cast<size_t>(j) * kBwdIn;
  const float g = P.d_hbar[item];
  float hb = 0.f;
  for (int k = 0; k < kBwdK; ++k) {
    const __half* xr = P.x + (r * kBwdK + k) * kBwdIn;
    float h = 0.f;
    for (int i = 0; i < kBwdIn; ++i) h += __half2float(xr[i]) * __half2float(wr[i]);
    const float w = P.sam_w[r * kBwdK + k];
    hb += w * round_f16(fmaxf(h, 0.f));
    P.dh[(r * kBwdK + k) * kBwdHid + j] = h > 0.f ? w * g : 0.f;
  }
  P.hbar[item] = round_f16(hb);
}

// K3: dx[row,i] = sum_j dh[row,j] * W1[j,i]                                        item = (row, i)
SNRF_HD void bwd_dx_one(const FeatBwdParams& P, int64_t item) {
  const int64_t row = item / kBwdIn;
  const int i = static_cast<int>(item % kBwdIn);
  const float* d = P.dh + row * kBwdHid;
  float a = 0.f;
  for (int j = 0; j < kBwdHid; ++j) a += d[j] * __half2float(P.w1[static_cast<size_t>(j) * kBwdIn + i]);
  P.dx[item] = a;
}

// K4: C[a,b] += sum_{row in slab} A[row,a] * B[row,b]   (weight gradients)          item = (slab, a, b)
//     W1: A = dh [R,256], B = x (fp16) [R,192];  W2: A = d_out [N,n_out], B = hbar [N,256]
template <typename TB>
SNRF_HD float ld_f32(const TB* p);
template <>
SNRF_HD float ld_f32<float>(const float* p) { return *p; }
template <>
SNRF_HD float ld_f32<__half>(const __half* p) { return __half2float(*p); }

template <typename TB>
SNRF_HD void bwd_wgrad_one(const float* A, int na, const TB* B, int nb, int64_t rows, int slab_rows, float* C,
                           int64_t item) {
  const int b = static_cast<int>(item % nb);
  const int a = static_cast<int>((item / nb) % na);
  const int64_t slab = item / (static_cast<int64_t>(na) * nb);
  const int64_t r0 = slab * slab_rows;
  const int64_t r1 = r0 + slab_rows < rows ? r0 + slab_rows : rows;
  float acc = 0.f;
  for (int64_t r = r0; r < r1; ++r) acc += A[r * na + a] * ld_f32<TB>(B + r * nb + b);
  // non-finite rows (rays whose top-k weights are 0/0 = NaN, sam_model.py:248) poison the sum exactly as they do in
  // the reference's dense autograd; nothing is filtered here
  atomic_add_f32(C + static_cast<size_t>(a) * nb + b, acc);
}

// K5: scatter dx into the hash tables: grad[idx(corner), f] += w(corner) * dx[row, 96*e + 8*l + f]
//                                                                               item = (row, encoding e, level l)
SNRF_HD void bwd_scatter_one(const FeatBwdParams& P, int64_t item) {
  const int l = static_cast<int>(item % 12);
  const int e = static_cast<int>((item / 12) % 2);
  const int64_t row = item / 24;
  const int64_t r = row / kBwdK;
  const float tm2 = P.sam_t[row];
  // positions exactly as the forward kernel builds them (sam.cu): pos = o + d * (ts + te) / 2
  const float px = sample_coord(P.origins[3 * r + 0], P.dirs[3 * r + 0], tm2);
  const float py = sample_coord(P.origins[3 * r + 1], P.dirs[3 * r + 1], tm2);
  const float pz = sample_coord(P.origins[3 * r + 2], P.dirs[3 * r + 2], tm2);
  float x, y, z, sel;
  contract_normalize(px, py, pz, false, false, x, y, z, sel);
  const GridLevel& L = P.enc[e].lv[l];
  const float qx = mul_add_rn(x, L.scale, 0.5f), qy = mul_add_rn(y, L.scale, 0.5f), qz = mul_add_rn(z, L.scale, 0.5f);
  const float fx = floorf(qx), fy = floorf(qy), fz = floorf(qz);
  const float rx = qx - fx, ry = qy - fy, rz = qz - fz;
  const uint32_t gx = static_cast<uint32_t>(static_cast<int>(fx)), gy = static_cast<uint32_t>(static_cast<int>(fy)),
                 gz = static_cast<uint32_t>(static_cast<int>(fz));
  const float* d = P.dx + row * kBwdIn + e * 96 + l * 8;
  float g[8];
  bool any = false;
  for (int f = 0; f < 8; ++f) {
    g[f] = d[f];
    any = any || g[f] != 0.f;
  }
  if (!any) return;  // dead hidden units / zero-weight samples contribute nothing
  float* table = P.g_table[e];
  for (int c = 0; c < 8; ++c) {
    float w = (c & 1) ? rx : 1.f - rx;
    w *= (c & 2) ? ry : 1.f - ry;
    w *= (c & 4) ? rz : 1.f - rz;
    const uint32_t idx = grid_index(L, gx + (c & 1), gy + ((c >> 1) & 1), gz + (c >> 2));
    float* t = table + static_cast<size_t>(idx) * 8;
    for (int f = 0; f < 8; ++f) atomic_add_f32(t + f, w * g[f]);
  }
}

// rays per internal block of the backward pass (bounds the fp32 scratch: 28 KB per ray)
constexpr int64_t kBwdBlockRays = 8192;
size_t feat_bwd_scratch_floats(int64_t n_rays);
cudaError_t launch_feat_backward(const FeatBwdParams& P, cudaStream_t stream, int64_t* launches);

}  // namespace snrf
