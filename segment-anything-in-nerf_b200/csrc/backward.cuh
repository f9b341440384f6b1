// Backward passes of the field evaluations on the hot path.
//
// The reference trains through tinycudann's autograd: every tcnn module (sam_field.py:51,63,84,99;
// nerfacto_field.py:157-175,228-240; density_fields.py:92-100) is a torch.autograd.Function whose backward is
// CutlassMLP / FullyFusedMLP backward (dL/dW = dY^T X, dL/dX = dY W through the ReLU masks) followed by the grid
// backward that scatters dL/dx * trilinear weight into the 8 corners with atomicAdd (in-tree statement of that
// scatter: nerfstudio/field_components/cuda/csrc/temporal_gridencoder.cu:283-370), while torch autograd handles the
// compositing and the losses on [N,S] tensors.  Restated here in fp32 (tcnn: fp16 with loss scaling; fp32 is the more
// exact reading, and the gradients land directly in the flat fp32 layout of `params.grad`):
//
//   feature branch   out[r] = W2 . sum_k w[r,k] fp16(relu(W1 x[r,k])), x = concat(enc0, enc1)(contract_L2(pos))
//                    positions detached (sam_field.py:116), weights detached (sam_model.py:258-277)
//   nerfacto field   density = trunc_exp(o[0]) * sel, geo = o[1:16], o = W2 relu(W1 enc(contract_inf(pos)))
//                    rgb = sigmoid(Wh3 relu(Wh2 relu(Wh1 [SH(dir), geo, 1])))        (nerfacto_field.py:242-351)
//   proposal field   density = trunc_exp((W2 relu(W1 [enc, 0pad]))[0]) * sel          (density_fields.py:102-125)
//   trunc_exp backward: g * exp(clamp(x, -15, 15))                                    (activations.py:33-37)
//
// First version: every kernel is "one thread = one output item", no shared memory, no warp intrinsics, so that the
// bodies below run unchanged on the CPU in tests/emu against torch autograd through the oracle.  Rounding points of
// the forward pass (fp16 tables / encoder outputs / hidden activations) pass gradients straight through.
#pragma once
#include "common.cuh"

namespace snrf {

constexpr int kBwdK = 16;     // picked samples per ray (feature branch)
constexpr int kBwdIn = 192;   // feature encoder width (2 x 12 levels x 8 features)
constexpr int kBwdHid = 256;  // feature hidden width
constexpr int kSlabRows = 256;  // rows reduced per thread before one atomicAdd (weight gradients)

SNRF_HD void atomic_add_f32(float* p, float v) {
#ifdef __CUDA_ARCH__
  atomicAdd(p, v);
#else
  *p += v;  // host emulation is single-threaded
#endif
}

// F consecutive floats at a 4F-byte aligned address: one vector reduction (REDG.E.ADD.F32x4 / F32x2, sm_90+)
// instead of F scalar ones
template <int F>
SNRF_HD void atomic_add_vec(float* p, const float (&v)[F]) {
#ifdef __CUDA_ARCH__
  if (F == 8) {
    atomicAdd(reinterpret_cast<float4*>(p), make_float4(v[0], v[1], v[2], v[3]));
    atomicAdd(reinterpret_cast<float4*>(p) + 1, make_float4(v[4 % F], v[5 % F], v[6 % F], v[7 % F]));
  } else if (F == 2) {
    atomicAdd(reinterpret_cast<float2*>(p), make_float2(v[0], v[1]));
  } else {
    for (int f = 0; f < F; ++f) atomicAdd(p + f, v[f]);
  }
#else
  for (int f = 0; f < F; ++f) p[f] += v[f];
#endif
}

template <typename T>
SNRF_HD float ld_f32(const T* p);
template <>
SNRF_HD float ld_f32<float>(const float* p) { return *p; }
template <>
SNRF_HD float ld_f32<__half>(const __half* p) { return __half2float(*p); }

// ---------------------------------------------------------------------------------------------
// generic MLP pieces
// ---------------------------------------------------------------------------------------------
// dX[row,i] = (sum_{j<ny} dY[row,j] * W[j,i]) * (X ? X[row,i] > 0 : 1)                 item = (row, i), i < nx
// (X = the layer's saved post-ReLU input: relu'(pre) = [X > 0])
// `rows_dev` (optional, device memory): number of valid rows when it is only known on the device (compacted rows of the
// feature branch); items beyond it return at once.
SNRF_HD void mlp_dgrad_one(const float* dY, int ldy, int ny, const __half* W, int ldw, const __half* X, int ldx,
                           float* dX, int lddx, int nx, const int* rows_dev, int64_t item) {
  const int64_t row = item / nx;
  if (rows_dev && row >= *rows_dev) return;
  const int i = static_cast<int>(item % nx);
  const float* g = dY + row * ldy;
  float a = 0.f;
  for (int j = 0; j < ny; ++j) a += g[j] * __half2float(W[static_cast<size_t>(j) * ldw + i]);
  if (X && !(__half2float(X[row * ldx + i]) > 0.f)) a = 0.f;
  dX[row * lddx + i] = a;
}

// C[a,b] += sum_{row in slab} A[row,a] * B[row,b]                                       item = (slab, a, b)
// Non-finite rows (e.g. rays whose top-k weights are 0/0 = NaN, sam_model.py:248) poison the sum exactly as they do
// in the reference's dense autograd; nothing is filtered here.
// `bmap` (optional): row r of A pairs with row bmap[r] of B (compacted rows of the feature branch).
template <typename TB>
SNRF_HD void mlp_wgrad_one(const float* A, int lda, int na, const TB* B, int ldb, int nb, int64_t rows, float* C, int ldc,
                           const int* rows_dev, const int* bmap, int64_t item) {
  if (rows_dev) rows = *rows_dev;
  const int b = static_cast<int>(item % nb);
  const int a = static_cast<int>((item / nb) % na);
  const int64_t slab = item / (static_cast<int64_t>(na) * nb);
  const int64_t r0 = slab * kSlabRows;
  if (r0 >= rows) return;
  const int64_t r1 = r0 + kSlabRows < rows ? r0 + kSlabRows : rows;
  float acc = 0.f;
  for (int64_t r = r0; r < r1; ++r) acc += A[r * lda + a] * ld_f32<TB>(B + (bmap ? bmap[r] : r) * ldb + b);
  atomic_add_f32(C + static_cast<size_t>(a) * ldc + b, acc);
}
SNRF_HD int64_t mlp_wgrad_items(int64_t rows, int na, int nb) {
  return ((rows + kSlabRows - 1) / kSlabRows) * na * nb;
}

// ---------------------------------------------------------------------------------------------
// hash-grid scatter: grad[idx(corner), f] += w(corner) * dX[pt, col0 + F*l + f]          item = (point, level)
// positions are world-space; contraction / (p+2)/4 / selector exactly as the forward pass
// ---------------------------------------------------------------------------------------------
template <int F>
SNRF_HD void grid_scatter_one(const GridDev& G, bool linf, bool selector, const float* xyz, const float* dX, int lddx,
                              int col0, float* g_table, const int* rows_dev, int64_t item) {
  const int l = static_cast<int>(item % G.n_levels);
  const int64_t pt = item / G.n_levels;
  if (rows_dev && pt >= *rows_dev) return;
  const float* d = dX + pt * lddx + col0 + l * F;
  float g[F];
  bool any = false;
  for (int f = 0; f < F; ++f) {
    g[f] = d[f];
    any = any || g[f] != 0.f;
  }
  if (!any) return;  // dead units / zero-weight samples contribute nothing
  float x, y, z, sel;
  contract_normalize(xyz[3 * pt], xyz[3 * pt + 1], xyz[3 * pt + 2], linf, selector, x, y, z, sel);
  const GridLevel& L = G.lv[l];
  const float qx = mul_add_rn(x, L.scale, 0.5f), qy = mul_add_rn(y, L.scale, 0.5f), qz = mul_add_rn(z, L.scale, 0.5f);
  const float fx = floorf(qx), fy = floorf(qy), fz = floorf(qz);
  const float rx = qx - fx, ry = qy - fy, rz = qz - fz;
  const uint32_t gx = static_cast<uint32_t>(static_cast<int>(fx)), gy = static_cast<uint32_t>(static_cast<int>(fy)),
                 gz = static_cast<uint32_t>(static_cast<int>(fz));
  for (int c = 0; c < 8; ++c) {
    float w = (c & 1) ? rx : 1.f - rx;
    w *= (c & 2) ? ry : 1.f - ry;
    w *= (c & 4) ? rz : 1.f - rz;
    const uint32_t idx = grid_index(L, gx + (c & 1), gy + ((c >> 1) & 1), gz + (c >> 2));
    float wg[F];
    for (int f = 0; f < F; ++f) wg[f] = w * g[f];
    atomic_add_vec<F>(g_table + static_cast<size_t>(idx) * F, wg);
  }
}

// ---------------------------------------------------------------------------------------------
// feature branch
// ---------------------------------------------------------------------------------------------
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
  float* xyz;             // [N*16,3]
  // Row compaction: only the leading slots of a ray whose weight is not below `cutoff` get a row (the same rule as
  // the bucketed forward kernel, sam_bucket.cu); cutoff < 0 keeps all 16.  dh / dx / xyz are indexed by compact row.
  float cutoff;
  int* n_rows;            // [1] number of compact rows (device counter, zeroed by the launcher)
  int* row_map;           // [N*16] compact row -> ray*16 + slot
  int* row_start;         // [N] first compact row of the ray
  int* row_k;             // [N] rows of the ray
  // outputs, accumulated (+=)
  float* g_w1;            // [256,192]
  float* g_w2;            // [n_out,256]
  float* g_table[2];      // [entries*8] each
};

// F0: rows of each ray                                                                        item = ray
SNRF_HD void feat_rows_assign_one(const FeatBwdParams& P, int64_t ray) {
  int k = kBwdK;
  if (!(P.cutoff < 0.f)) {
    k = 0;
    for (int s = 0; s < kBwdK; ++s) {
      const float w = P.sam_w[ray * kBwdK + s];
      if (!(w < P.cutoff) && w != 0.f) k = s + 1;  // NaN weights (0/0 rays) count as significant
    }
  }
#ifdef __CUDA_ARCH__
  const int pos = atomicAdd(P.n_rows, k);
#else
  const int pos = *P.n_rows;
  *P.n_rows += k;
#endif
  P.row_start[ray] = pos;
  P.row_k[ray] = k;
  for (int s = 0; s < k; ++s) P.row_map[pos + s] = static_cast<int>(ray * kBwdK + s);
}

// F1: recompute h[r,k,j] = W1[j,:] . x[r,k,:]; hbar[r,j] = fp16(sum_k w_k * fp16(relu(h)));
//     dh[r,k,j] = h > 0 ? w_k * d_hbar[r,j] : 0   (d_hbar from mlp_dgrad_one)          item = (ray, j)
SNRF_HD void feat_hidden_one(const FeatBwdParams& P, int64_t item) {
  const int64_t r = item / kBwdHid;
  const int j = static_cast<int>(item % kBwdHid);
  const __half* wr = P.w1 + static_cast<size_t>(j) * kBwdIn;
  const float g = P.d_hbar[item];
  const int64_t row0 = P.row_start[r];
  const int nk = P.row_k[r];
  float hb = 0.f;
  for (int k = 0; k < nk; ++k) {
    const __half* xr = P.x + (r * kBwdK + k) * kBwdIn;
    float h = 0.f;
    for (int i = 0; i < kBwdIn; ++i) h += __half2float(xr[i]) * __half2float(wr[i]);
    const float w = P.sam_w[r * kBwdK + k];
    hb += w * round_f16(fmaxf(h, 0.f));
    P.dh[(row0 + k) * kBwdHid + j] = h > 0.f ? w * g : 0.f;
  }
  P.hbar[item] = round_f16(hb);
}

// F2: world positions of the picked samples, op for op as the forward kernel builds them          item = row
SNRF_HD void feat_positions_one(const FeatBwdParams& P, int64_t row) {
  if (row >= *P.n_rows) return;
  const int64_t orig = P.row_map[row];
  const int64_t r = orig / kBwdK;
  const float tm2 = P.sam_t[orig];
  for (int c = 0; c < 3; ++c) P.xyz[3 * row + c] = sample_coord(P.origins[3 * r + c], P.dirs[3 * r + c], tm2);
}

// The chain of items for one block of rays, written once against an executor so that the CUDA launcher (backward.cu)
// and the host emulation (tests/emu) run the same wiring of buffers, strides and offsets.
template <class Exec>
inline void feat_backward_chain(const FeatBwdParams& P, Exec& ex) {
  const int64_t n = P.n_rays, rows = n * kBwdK;  // `rows` is the worst case; the real count lives in *P.n_rows
  ex.feat_rows_assign(P, n);
  // d_hbar = d_out . W2 ; hidden recompute, hbar, dh ; dx = dh . W1
  ex.dgrad(P.d_out, P.n_out, P.n_out, P.w2, kBwdHid, nullptr, 0, P.d_hbar, kBwdHid, kBwdHid, n, nullptr);
  ex.feat_hidden(P, n * kBwdHid);
  ex.dgrad(P.dh, kBwdHid, kBwdHid, P.w1, kBwdIn, nullptr, 0, P.dx, kBwdIn, kBwdIn, rows, P.n_rows);
  // weight gradients
  ex.wgrad_h(P.dh, kBwdHid, kBwdHid, P.x, kBwdIn, kBwdIn, rows, P.g_w1, kBwdIn, P.n_rows, P.row_map);
  ex.wgrad_f(P.d_out, P.n_out, P.n_out, P.hbar, kBwdHid, kBwdHid, n, P.g_w2, kBwdHid);
  // table scatter (L2 contraction, no selector: sam_field.py:32,116-118)
  ex.feat_positions(P, rows);
  for (int e = 0; e < 2; ++e)
    ex.scatter8(P.enc[e], false, false, P.xyz, P.dx, kBwdIn, e * 96, P.g_table[e], rows, P.n_rows);
}
constexpr int kFeatBwdLaunches = 9;

// rays per internal block of the feature backward (bounds the fp32 scratch: 30 KB per ray)
constexpr int64_t kBwdBlockRays = 8192;
size_t feat_bwd_scratch_floats(int64_t n_rays);
cudaError_t launch_feat_backward(const FeatBwdParams& P, cudaStream_t stream, int64_t* launches);

// ---------------------------------------------------------------------------------------------
// density fields (nerfacto base + colour head, proposal)
// ---------------------------------------------------------------------------------------------
// d_pre[row, c] = d_rgb[row, c] * s (1 - s), s = sigmoid(pre[row, c])                    item = (row, c), c < 3
SNRF_HD void sigmoid_bwd_one(const float* d_rgb, const __half* pre, int ldp, float* d_pre, int64_t item) {
  const int64_t row = item / 3;
  const int c = static_cast<int>(item % 3);
  const float s = 1.f / (1.f + expf(-__half2float(pre[row * ldp + c])));
  d_pre[item] = d_rgb[item] * s * (1.f - s);
}

// d_o[row, 0] = d_density[row] * exp(clamp(o[row,0], -15, 15)) * sel[row];  d_o[row, 1 + k] = d_geo[row, k]
//                                                                                      item = (row, c), c < n_o
SNRF_HD void density_bwd_one(const float* d_density, const __half* o, int ldo, const float* sel, const float* d_geo,
                             int ldg, int col_geo, float* d_o, int n_o, int64_t item) {
  const int64_t row = item / n_o;
  const int c = static_cast<int>(item % n_o);
  float v;
  if (c == 0) {
    const float pre = fminf(fmaxf(__half2float(o[row * ldo]), -15.f), 15.f);
    v = d_density ? d_density[row] * expf(pre) * sel[row] : 0.f;
  } else {
    v = d_geo ? d_geo[row * ldg + col_geo + c - 1] : 0.f;
  }
  d_o[item] = v;
}

struct FieldBwdParams {
  const float* xyz;        // [n,3] world positions
  const float* dirs;       // [n,3] (nerfacto colour head) or null
  const float* d_density;  // [n] or null
  const float* d_rgb;      // [n,3] or null
  int64_t n;
  int which;               // 0 proposal, 1 nerfacto
  GridDev grid;
  // forward weights (row-major fp16, tcnn [out,in])
  const __half *w1, *w2, *wh1, *wh2, *wh3;
  // saved / recomputed forward activations (fp16) and scratch (fp32): carved out of one buffer by the launcher
  __half *x, *h1, *o, *hx, *g1, *g2, *pre3;
  float *sel, *d_pre3, *d_g2, *d_g1, *d_hx, *d_o, *d_h1, *d_x;
  // outputs, accumulated (+=): flat tcnn layouts
  float* g_base;           // [W1, W2, table] = field.mlp_base.params / proposal_networks.0.mlp_base.params
  float* g_head;           // [Wh1, Wh2, Wh3] = field.mlp_head.params (nerfacto only)
};
// Backward chain for one block of samples whose forward activations are already in P (see feat_backward_chain).
template <class Exec>
inline void field_backward_chain(const FieldBwdParams& P, Exec& ex) {
  const bool nerfacto = P.which == 1;
  const int width = nerfacto ? 32 : 10, hidden = nerfacto ? 64 : 16, k_w = nerfacto ? 32 : 16;
  const int n_net = nerfacto ? 64 * 32 + 16 * 64 : 16 * 16 + 16 * 16;
  const int64_t n = P.n;
  const bool colour = nerfacto && P.d_rgb != nullptr;
  if (colour) {  // rgb = sigmoid(Wh3 relu(Wh2 relu(Wh1 hx))), only outputs 0..2 of the 16 padded ones are used
    ex.sigmoid_bwd(P.d_rgb, P.pre3, 16, P.d_pre3, n * 3);
    ex.dgrad(P.d_pre3, 3, 3, P.wh3, 64, P.g2, 64, P.d_g2, 64, 64, n, nullptr);
    ex.wgrad_h(P.d_pre3, 3, 3, P.g2, 64, 64, n, P.g_head + 64 * 32 + 64 * 64, 64, nullptr, nullptr);
    ex.dgrad(P.d_g2, 64, 64, P.wh2, 64, P.g1, 64, P.d_g1, 64, 64, n, nullptr);
    ex.wgrad_h(P.d_g2, 64, 64, P.g1, 64, 64, n, P.g_head + 64 * 32, 64, nullptr, nullptr);
    ex.dgrad(P.d_g1, 64, 64, P.wh1, 32, nullptr, 0, P.d_hx, 32, 32, n, nullptr);
    ex.wgrad_h(P.d_g1, 64, 64, P.hx, 32, 32, n, P.g_head, 32, nullptr, nullptr);
  }
  // density (output 0) and, for nerfacto, the 15 geo features (outputs 1..15 = head inputs 16..30)
  const int n_o = nerfacto ? 16 : 1;
  ex.density_bwd(P.d_density, P.o, 16, P.sel, colour ? P.d_hx : nullptr, 32, 16, P.d_o, n_o, n * n_o);
  ex.dgrad(P.d_o, n_o, n_o, P.w2, hidden, P.h1, hidden, P.d_h1, hidden, hidden, n, nullptr);
  ex.wgrad_h(P.d_o, n_o, n_o, P.h1, hidden, hidden, n, P.g_base + hidden * k_w, hidden, nullptr, nullptr);
  ex.dgrad(P.d_h1, hidden, hidden, P.w1, k_w, nullptr, 0, P.d_x, width, width, n, nullptr);
  ex.wgrad_h(P.d_h1, hidden, hidden, P.x, width, width, n, P.g_base, k_w, nullptr, nullptr);
  ex.scatter2(P.grid, true, true, P.xyz, P.d_x, width, 0, P.g_base + n_net, n, nullptr);
}

// ---------------------------------------------------------------------------------------------
// ray-wise renderer ops (what torch autograd does for the reference on [N,S] tensors)
// ---------------------------------------------------------------------------------------------
constexpr int kMaxRaySamples = 64;

// RaySamples.get_weights backward (rays.py:141-163; deltas carry no gradient - the sampler detaches its bins,
// ray_samplers.py:357):  w_i = T_i - T_{i+1},  T_{i+1} = T_i exp(-delta_i sigma_i)
//   d sigma_k = delta_k * ( g_k * T_{k+1} - sum_{i>k} g_i * w_i )                                  item = ray
SNRF_HD void weights_bwd_one(const float* deltas, const float* dens, const float* g_w, float* d_dens, int S, int64_t ray) {
  float w[kMaxRaySamples], t_next[kMaxRaySamples];
  const float* dl = deltas + ray * S;
  const float* sg = dens + ray * S;
  const float* g = g_w + ray * S;
  float cum = 0.f;
  for (int i = 0; i < S; ++i) {
    const float ds = dl[i] * sg[i];
    const float T = expf(-cum);
    w[i] = (1.f - expf(-ds)) * T;
    cum += ds;
    t_next[i] = expf(-cum);
  }
  float suffix = 0.f;  // sum_{i>k} g_i w_i
  for (int k = S - 1; k >= 0; --k) {
    const bool finite = w[k] == w[k] && fabsf(w[k]) <= 3.402823466e+38f;  // nan_to_num blocks the gradient of a non-finite weight
    const float gk = finite ? g[k] : 0.f;
    d_dens[ray * S + k] = dl[k] * (gk * t_next[k] - suffix);
    suffix += gk * (finite ? w[k] : 0.f);
  }
}

// RGBRenderer.combine_rgb backward (renderers.py:69-112): C = sum_i w_i c_i + bg (1 - sum_i w_i), bg = c_{S-1}
// ("last_sample") or a fixed colour:  d c_i = w_i G (+ (1 - acc) G for the last sample), d w_i = G . (c_i - bg)
//                                                                                              item = (ray, i)
SNRF_HD void rgb_bwd_one(const float* rgb, const float* w, const float* g_out, int bg_fixed, float bg0, float bg1, float bg2,
                         float* d_rgb, float* d_w, int S, int64_t item) {
  const int64_t ray = item / S;
  const int i = static_cast<int>(item % S);
  const float* c = rgb + (ray * S + i) * 3;
  const float* G = g_out + ray * 3;
  const float* last = rgb + (ray * S + S - 1) * 3;
  const float b0 = bg_fixed ? bg0 : last[0], b1 = bg_fixed ? bg1 : last[1], b2 = bg_fixed ? bg2 : last[2];
  const float wi = w[ray * S + i];
  float extra = 0.f;
  if (!bg_fixed && i == S - 1) {
    float acc = 0.f;
    for (int j = 0; j < S; ++j) acc += w[ray * S + j];
    extra = 1.f - acc;
  }
  for (int ch = 0; ch < 3; ++ch) d_rgb[(ray * S + i) * 3 + ch] = (wi + extra) * G[ch];
  d_w[ray * S + i] = G[0] * (c[0] - b0) + G[1] * (c[1] - b1) + G[2] * (c[2] - b2);
}
// Top-k pick + sharpening of the feature samples (samnerf/sam_model.py:244-255) as a stand-alone ray op, so that the
// training path can take its picks from the weights it differentiates instead of from a second fused render:
// rank by weight (ties by sample index, like march.cu), keep rank < k, w^sharpen, renormalise; slot = rank, i.e.
// descending weight.  sam_t = start + end of the picked sample (2 x midpoint).                     item = ray
SNRF_HD void pick_samples_one(const float* w, const float* starts, const float* ends, int S, int k, float sharpen,
                              float* sam_t, float* sam_w, int64_t ray) {
  const float* wr = w + ray * S;
  float tot = 0.f;
  for (int i = 0; i < S; ++i) {
    int rank = 0;
    for (int j = 0; j < S; ++j) rank += (wr[j] > wr[i]) || (wr[j] == wr[i] && j < i);
    if (rank < k) tot += powf(wr[i], sharpen);
  }
  for (int i = 0; i < S; ++i) {
    int rank = 0;
    for (int j = 0; j < S; ++j) rank += (wr[j] > wr[i]) || (wr[j] == wr[i] && j < i);
    if (rank < k) {
      sam_t[ray * k + rank] = starts[ray * S + i] + ends[ray * S + i];
      sam_w[ray * k + rank] = powf(wr[i], sharpen) / tot;
    }
  }
}
cudaError_t launch_pick_samples(const float* w, const float* starts, const float* ends, int64_t n, int S, int k,
                                float sharpen, float* sam_t, float* sam_w, cudaStream_t stream);
cudaError_t launch_weights_bwd(const float* deltas, const float* dens, const float* g_w, float* d_dens, int64_t n, int S,
                               cudaStream_t stream);
cudaError_t launch_rgb_bwd(const float* rgb, const float* w, const float* g_out, int bg_fixed, const float* bg,
                           float* d_rgb, float* d_w, int64_t n, int S, cudaStream_t stream);

// ---------------------------------------------------------------------------------------------
// conv head (patch aggregation, sam_model.py:202-208,260-265): Conv3x3(pad 1) -> ReLU -> Conv3x3(pad 1) -> mean over
// the 4 x 4 patch.  Rows are patch-major with row-major positions inside a patch; torch's Conv2d weight
// [out][in][ky][kx] read as a row-major [256 x 2304] matrix with k = in*9 + ky*3 + kx is the GEMM operand of the
// im2col form, so weight gradients come out directly in the layout of `conv_head.{0,2}.weight.grad`.
// Operands are fp16 with fp32 accumulation like the forward kernel (deviation 3).
// ---------------------------------------------------------------------------------------------
constexpr int kConvC = 256, kConvP = 4, kConvPos = kConvP * kConvP, kConvK = kConvC * 9;

// row of the position shifted by tap (ky, kx) inside the same patch, or -1 outside the patch
SNRF_HD int64_t conv_shifted_row(int64_t row, int dy, int dx) {
  const int pos = static_cast<int>(row % kConvPos);
  const int y = pos / kConvP + dy, x = pos % kConvP + dx;
  if (y < 0 || y >= kConvP || x < 0 || x >= kConvP) return -1;
  return row - pos + y * kConvP + x;
}
// Xcol[row, i*9 + ky*3 + kx] = X[shifted(row, ky-1, kx-1), i] (0 outside)                 item = (row, k)
template <typename TX>
SNRF_HD void conv_im2col_one(const TX* X, __half* Xcol, int64_t item) {
  const int64_t row = item / kConvK;
  const int k = static_cast<int>(item % kConvK);
  const int i = k / 9, t = k % 9;
  const int64_t src = conv_shifted_row(row, t / 3 - 1, t % 3 - 1);
  Xcol[item] = __float2half_rn(src < 0 ? 0.f : ld_f32<TX>(X + src * kConvC + i));
}
// Y[row,o] = fp16(act(sum_k Xcol[row,k] W[o,k] + b[o]))                                   item = (row, o)
SNRF_HD void conv_fwd_one(const __half* Xcol, const __half* W, const float* b, __half* Y, int relu, int64_t item) {
  const int64_t row = item / kConvC;
  const int o = static_cast<int>(item % kConvC);
  const __half* x = Xcol + row * kConvK;
  const __half* w = W + static_cast<size_t>(o) * kConvK;
  float a = 0.f;
  for (int k = 0; k < kConvK; ++k) a += __half2float(x[k]) * __half2float(w[k]);
  a += b[o];
  if (relu) a = fmaxf(a, 0.f);
  Y[item] = __float2half_rn(a);
}
// d_y[row,o] = d_out[patch(row), o] / 16 (backward of the mean over the patch)            item = (row, o)
SNRF_HD void conv_mean_bwd_one(const float* d_out, float* d_y, int64_t item) {
  const int64_t row = item / kConvC;
  d_y[item] = d_out[(row / kConvPos) * kConvC + item % kConvC] / static_cast<float>(kConvPos);
}
// adjoint of im2col: dX[row,i] = sum_t dXcol[shifted(row, 1-ky, 1-kx), i*9 + t], times [mask > 0] if given
//                                                                                         item = (row, i)
SNRF_HD void conv_col2im_one(const float* dXcol, const __half* mask, float* dX, int64_t item) {
  const int64_t row = item / kConvC;
  const int i = static_cast<int>(item % kConvC);
  float a = 0.f;
  for (int t = 0; t < 9; ++t) {
    const int64_t src = conv_shifted_row(row, 1 - t / 3, 1 - t % 3);
    if (src >= 0) a += dXcol[src * kConvK + i * 9 + t];
  }
  if (mask && !(__half2float(mask[item]) > 0.f)) a = 0.f;
  dX[item] = a;
}
// g_b[o] += sum_{row in slab} dY[row,o]                                                   item = (slab, o)
SNRF_HD void conv_bias_grad_one(const float* dY, int64_t rows, float* g_b, int64_t item) {
  const int o = static_cast<int>(item % kConvC);
  const int64_t r0 = (item / kConvC) * kSlabRows;
  const int64_t r1 = r0 + kSlabRows < rows ? r0 + kSlabRows : rows;
  float a = 0.f;
  for (int64_t r = r0; r < r1; ++r) a += dY[r * kConvC + o];
  atomic_add_f32(g_b + o, a);
}

struct ConvBwdParams {
  const float* feat_in;   // [rows,256] fp32 rows as handed to snrf_patch_aggregate
  const float* d_out;     // [rows/16,256]
  int64_t rows;
  const __half *w1, *w2;  // [256 x 2304] row-major fp16 (torch layout)
  const float *b1, *b2;   // [256]
  // scratch
  __half *xcol1, *xcol2, *hid;  // [rows,2304] x2, [rows,256]
  float *d_y, *d_xcol, *d_hid;  // [rows,256], [rows,2304], [rows,256]
  // outputs: parameter gradients accumulated (+=), d_feat written
  float *g_w1, *g_b1, *g_w2, *g_b2;
  float* d_feat;          // [rows,256] or null
};
template <class Exec>
inline void conv_backward_chain(const ConvBwdParams& P, Exec& ex) {
  const int64_t n = P.rows;
  // forward recomputation
  ex.im2col_f32(P.feat_in, P.xcol1, n * kConvK);
  ex.conv_fwd(P.xcol1, P.w1, P.b1, P.hid, 1, n * kConvC);
  ex.im2col_f16(P.hid, P.xcol2, n * kConvK);
  // second conv
  ex.conv_mean_bwd(P.d_out, P.d_y, n * kConvC);
  ex.bias_grad(P.d_y, n, P.g_b2);
  ex.wgrad_h(P.d_y, kConvC, kConvC, P.xcol2, kConvK, kConvK, n, P.g_w2, kConvK, nullptr, nullptr);
  ex.dgrad(P.d_y, kConvC, kConvC, P.w2, kConvK, nullptr, 0, P.d_xcol, kConvK, kConvK, n, nullptr);
  ex.col2im(P.d_xcol, P.hid, P.d_hid, n * kConvC);
  // first conv
  ex.bias_grad(P.d_hid, n, P.g_b1);
  ex.wgrad_h(P.d_hid, kConvC, kConvC, P.xcol1, kConvK, kConvK, n, P.g_w1, kConvK, nullptr, nullptr);
  if (P.d_feat) {
    ex.dgrad(P.d_hid, kConvC, kConvC, P.w1, kConvK, nullptr, 0, P.d_xcol, kConvK, kConvK, n, nullptr);
    ex.col2im(P.d_xcol, nullptr, P.d_feat, n * kConvC);
  }
}
constexpr int64_t kConvBwdBlockPatches = 256;  // 4096 rows, 21 KB of scratch per row
size_t conv_bwd_scratch_bytes(int64_t n_patches);
cudaError_t launch_conv_backward(const ConvBwdParams& P, void* scratch, cudaStream_t stream, int64_t* launches);

constexpr int64_t kFieldBwdBlock = 1 << 16;  // samples per internal block (1.8 KB of scratch per sample)
size_t field_bwd_scratch_bytes(int64_t n);
// `fwd` runs the forward recomputation (encode + dense layers) into the activation buffers: supplied by api.cu so that
// the GPU-verified query kernels are the ones that run
cudaError_t launch_field_backward(const FieldBwdParams& P, void* scratch, cudaStream_t stream, int64_t* launches);

}  // namespace snrf
