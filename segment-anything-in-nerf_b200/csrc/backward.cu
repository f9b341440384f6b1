// Kernels F* (feature branch) and D* (density fields) - the backward passes of backward.cuh.  First, simple version:
// one thread per output item, fp32 throughout, atomics for the reductions; forward activations are recomputed with
// the component-query kernels of query.cu.  The dominant cost of the feature branch is the table scatter
// (16 x 24 x 8 corners x 8 features = 24 576 fp32 atomics per ray, ~98 KB of read-modify-write traffic per ray against
// 49 KB gathered by the forward pass); the fused tensor-core version that mirrors sam.cu is the next step once this
// one is parity-green on hardware.
#include <string.h>

#include "backward.cuh"
#include "kernels.cuh"

namespace snrf {
namespace {

constexpr int kThreads = 256;
inline unsigned blocks_for(int64_t items) { return static_cast<unsigned>((items + kThreads - 1) / kThreads); }

__global__ void mlp_dgrad_kernel(const float* dY, int ldy, int ny, const __half* W, int ldw, const __half* X, int ldx,
                                 float* dX, int lddx, int nx, const int* rows_dev, int64_t items) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < items) mlp_dgrad_one(dY, ldy, ny, W, ldw, X, ldx, dX, lddx, nx, rows_dev, i);
}
template <typename TB>
__global__ void mlp_wgrad_kernel(const float* A, int lda, int na, const TB* B, int ldb, int nb, int64_t rows, float* C,
                                 int ldc, const int* rows_dev, const int* bmap, int64_t items) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < items) mlp_wgrad_one<TB>(A, lda, na, B, ldb, nb, rows, C, ldc, rows_dev, bmap, i);
}
template <int F>
__global__ void grid_scatter_kernel(const GridDev G, bool linf, bool selector, const float* xyz, const float* dX, int lddx,
                                    int col0, float* g_table, const int* rows_dev, int64_t items) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < items) grid_scatter_one<F>(G, linf, selector, xyz, dX, lddx, col0, g_table, rows_dev, i);
}
__global__ void feat_rows_assign_kernel(const FeatBwdParams P, int64_t items) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < items) feat_rows_assign_one(P, i);
}
__global__ void feat_hidden_kernel(const FeatBwdParams P, int64_t items) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < items) feat_hidden_one(P, i);
}
__global__ void feat_positions_kernel(const FeatBwdParams P, int64_t items) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < items) feat_positions_one(P, i);
}
__global__ void sigmoid_bwd_kernel(const float* d_rgb, const __half* pre, int ldp, float* d_pre, int64_t items) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < items) sigmoid_bwd_one(d_rgb, pre, ldp, d_pre, i);
}
__global__ void density_bwd_kernel(const float* d_density, const __half* o, int ldo, const float* sel, const float* d_geo,
                                   int ldg, int col_geo, float* d_o, int n_o, int64_t items) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < items) density_bwd_one(d_density, o, ldo, sel, d_geo, ldg, col_geo, d_o, n_o, i);
}

__global__ void weights_bwd_kernel(const float* deltas, const float* dens, const float* g_w, float* d_dens, int S, int64_t n) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) weights_bwd_one(deltas, dens, g_w, d_dens, S, i);
}
__global__ void rgb_bwd_kernel(const float* rgb, const float* w, const float* g_out, int bg_fixed, float bg0, float bg1,
                               float bg2, float* d_rgb, float* d_w, int S, int64_t items) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < items) rgb_bwd_one(rgb, w, g_out, bg_fixed, bg0, bg1, bg2, d_rgb, d_w, S, i);
}

template <typename TX>
__global__ void conv_im2col_kernel(const TX* X, __half* Xcol, int64_t items) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < items) conv_im2col_one<TX>(X, Xcol, i);
}
__global__ void conv_fwd_kernel(const __half* Xcol, const __half* W, const float* b, __half* Y, int relu, int64_t items) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < items) conv_fwd_one(Xcol, W, b, Y, relu, i);
}
__global__ void conv_mean_bwd_kernel(const float* d_out, float* d_y, int64_t items) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < items) conv_mean_bwd_one(d_out, d_y, i);
}
__global__ void conv_col2im_kernel(const float* dXcol, const __half* mask, float* dX, int64_t items) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < items) conv_col2im_one(dXcol, mask, dX, i);
}
__global__ void conv_bias_grad_kernel(const float* dY, int64_t rows, float* g_b, int64_t items) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < items) conv_bias_grad_one(dY, rows, g_b, i);
}

__global__ void pick_samples_kernel(const float* w, const float* starts, const float* ends, int S, int k, float sharpen,
                                    float* sam_t, float* sam_w, int64_t n) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) pick_samples_one(w, starts, ends, S, k, sharpen, sam_t, sam_w, i);
}

// executor of the backward chains of backward.cuh: every step is one kernel launch on `s`
struct DeviceExec {
  cudaStream_t s;
  void im2col_f32(const float* X, __half* Xcol, int64_t items) { conv_im2col_kernel<float><<<blocks_for(items), kThreads, 0, s>>>(X, Xcol, items); }
  void im2col_f16(const __half* X, __half* Xcol, int64_t items) { conv_im2col_kernel<__half><<<blocks_for(items), kThreads, 0, s>>>(X, Xcol, items); }
  void conv_fwd(const __half* Xcol, const __half* W, const float* b, __half* Y, int relu, int64_t items) {
    conv_fwd_kernel<<<blocks_for(items), kThreads, 0, s>>>(Xcol, W, b, Y, relu, items);
  }
  void conv_mean_bwd(const float* d_out, float* d_y, int64_t items) { conv_mean_bwd_kernel<<<blocks_for(items), kThreads, 0, s>>>(d_out, d_y, items); }
  void col2im(const float* dXcol, const __half* mask, float* dX, int64_t items) {
    conv_col2im_kernel<<<blocks_for(items), kThreads, 0, s>>>(dXcol, mask, dX, items);
  }
  void bias_grad(const float* dY, int64_t rows, float* g_b) {
    const int64_t items = ((rows + kSlabRows - 1) / kSlabRows) * kConvC;
    conv_bias_grad_kernel<<<blocks_for(items), kThreads, 0, s>>>(dY, rows, g_b, items);
  }
  void dgrad(const float* dY, int ldy, int ny, const __half* W, int ldw, const __half* X, int ldx, float* dX, int lddx,
             int nx, int64_t rows, const int* rows_dev) {
    const int64_t items = rows * nx;
    mlp_dgrad_kernel<<<blocks_for(items), kThreads, 0, s>>>(dY, ldy, ny, W, ldw, X, ldx, dX, lddx, nx, rows_dev, items);
  }
  void wgrad_h(const float* A, int lda, int na, const __half* B, int ldb, int nb, int64_t rows, float* C, int ldc,
               const int* rows_dev, const int* bmap) {
    const int64_t items = mlp_wgrad_items(rows, na, nb);
    mlp_wgrad_kernel<__half><<<blocks_for(items), kThreads, 0, s>>>(A, lda, na, B, ldb, nb, rows, C, ldc, rows_dev, bmap, items);
  }
  void wgrad_f(const float* A, int lda, int na, const float* B, int ldb, int nb, int64_t rows, float* C, int ldc) {
    const int64_t items = mlp_wgrad_items(rows, na, nb);
    mlp_wgrad_kernel<float><<<blocks_for(items), kThreads, 0, s>>>(A, lda, na, B, ldb, nb, rows, C, ldc, nullptr, nullptr, items);
  }
  void scatter2(const GridDev& G, bool linf, bool sel, const float* xyz, const float* dX, int lddx, int col0, float* g,
                int64_t points, const int* rows_dev) {
    const int64_t items = points * G.n_levels;
    grid_scatter_kernel<2><<<blocks_for(items), kThreads, 0, s>>>(G, linf, sel, xyz, dX, lddx, col0, g, rows_dev, items);
  }
  void scatter8(const GridDev& G, bool linf, bool sel, const float* xyz, const float* dX, int lddx, int col0, float* g,
                int64_t points, const int* rows_dev) {
    const int64_t items = points * G.n_levels;
    grid_scatter_kernel<8><<<blocks_for(items), kThreads, 0, s>>>(G, linf, sel, xyz, dX, lddx, col0, g, rows_dev, items);
  }
  void feat_rows_assign(const FeatBwdParams& P, int64_t items) { feat_rows_assign_kernel<<<blocks_for(items), kThreads, 0, s>>>(P, items); }
  void feat_hidden(const FeatBwdParams& P, int64_t items) { feat_hidden_kernel<<<blocks_for(items), kThreads, 0, s>>>(P, items); }
  void feat_positions(const FeatBwdParams& P, int64_t items) { feat_positions_kernel<<<blocks_for(items), kThreads, 0, s>>>(P, items); }
  void sigmoid_bwd(const float* d_rgb, const __half* pre, int ldp, float* d_pre, int64_t items) {
    sigmoid_bwd_kernel<<<blocks_for(items), kThreads, 0, s>>>(d_rgb, pre, ldp, d_pre, items);
  }
  void density_bwd(const float* d_density, const __half* o, int ldo, const float* sel, const float* d_geo, int ldg,
                   int col_geo, float* d_o, int n_o, int64_t items) {
    density_bwd_kernel<<<blocks_for(items), kThreads, 0, s>>>(d_density, o, ldo, sel, d_geo, ldg, col_geo, d_o, n_o, items);
  }
};

}  // namespace

cudaError_t launch_weights_bwd(const float* deltas, const float* dens, const float* g_w, float* d_dens, int64_t n, int S,
                               cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  if (S < 1 || S > kMaxRaySamples) return cudaErrorInvalidValue;
  weights_bwd_kernel<<<blocks_for(n), kThreads, 0, stream>>>(deltas, dens, g_w, d_dens, S, n);
  return cudaGetLastError();
}
cudaError_t launch_pick_samples(const float* w, const float* starts, const float* ends, int64_t n, int S, int k,
                                float sharpen, float* sam_t, float* sam_w, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  if (S < 1 || k < 1 || k > S) return cudaErrorInvalidValue;
  pick_samples_kernel<<<blocks_for(n), kThreads, 0, stream>>>(w, starts, ends, S, k, sharpen, sam_t, sam_w, n);
  return cudaGetLastError();
}
cudaError_t launch_rgb_bwd(const float* rgb, const float* w, const float* g_out, int bg_fixed, const float* bg,
                           float* d_rgb, float* d_w, int64_t n, int S, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  if (S < 1) return cudaErrorInvalidValue;
  const int64_t items = n * S;
  rgb_bwd_kernel<<<blocks_for(items), kThreads, 0, stream>>>(rgb, w, g_out, bg_fixed, bg ? bg[0] : 0.f, bg ? bg[1] : 0.f,
                                                            bg ? bg[2] : 0.f, d_rgb, d_w, S, items);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// feature branch
// ---------------------------------------------------------------------------------------------
size_t feat_bwd_scratch_floats(int64_t n_rays) {
  const int64_t b = n_rays < kBwdBlockRays ? n_rays : kBwdBlockRays;
  // fp32: d_hbar, hbar [b,256]; dh [b*16,256]; dx [b*16,192]; xyz [b*16,3]; int32: n_rows (4), row_map [b*16], row_start, row_k [b]
  return static_cast<size_t>(b) * (2 * kBwdHid + kBwdK * kBwdHid + kBwdK * kBwdIn + kBwdK * 3 + kBwdK + 2) + 4;
}

// P.d_hbar must point at feat_bwd_scratch_floats(P.n_rays) floats; hbar / dh / dx / xyz are carved out of it here.
cudaError_t launch_feat_backward(const FeatBwdParams& P0, cudaStream_t stream, int64_t* launches) {
  float* scratch = P0.d_hbar;
  for (int64_t r0 = 0; r0 < P0.n_rays; r0 += kBwdBlockRays) {
    const int64_t n = P0.n_rays - r0 < kBwdBlockRays ? P0.n_rays - r0 : kBwdBlockRays;
    FeatBwdParams P = P0;
    P.origins += 3 * r0;
    P.dirs += 3 * r0;
    P.sam_t += kBwdK * r0;
    P.sam_w += kBwdK * r0;
    P.d_out += static_cast<int64_t>(P.n_out) * r0;
    P.x += kBwdK * kBwdIn * r0;
    P.n_rays = n;
    P.d_hbar = scratch;
    P.hbar = P.d_hbar + n * kBwdHid;
    P.dh = P.hbar + n * kBwdHid;
    P.dx = P.dh + n * kBwdK * kBwdHid;
    P.xyz = P.dx + n * kBwdK * kBwdIn;
    int* ints = reinterpret_cast<int*>(P.xyz + n * kBwdK * 3);
    P.n_rows = ints;
    P.row_map = ints + 4;
    P.row_start = P.row_map + n * kBwdK;
    P.row_k = P.row_start + n;
    cudaError_t em = cudaMemsetAsync(P.n_rows, 0, sizeof(int), stream);
    if (em != cudaSuccess) return em;
    DeviceExec ex{stream};
    feat_backward_chain(P, ex);
    if (launches) *launches += kFeatBwdLaunches;
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

// ---------------------------------------------------------------------------------------------
// density fields
// ---------------------------------------------------------------------------------------------
namespace {
// per-sample scratch: fp16 x[32] h1[64] o[16] hx[32] g1[64] g2[64] pre3[16] geo[16] = 304 halfs,
//                     fp32 sel[1] dens[1] d_pre3[4] d_g2[64] d_g1[64] d_hx[32] d_o[16] d_h1[64] d_x[32] = 278 floats
constexpr size_t kFieldHalfs = 304, kFieldFloats = 278;
}  // namespace

size_t field_bwd_scratch_bytes(int64_t n) {
  const int64_t b = n < kFieldBwdBlock ? n : kFieldBwdBlock;
  return static_cast<size_t>(b) * (kFieldHalfs * 2 + kFieldFloats * 4) + 256;
}

cudaError_t launch_field_backward(const FieldBwdParams& P0, void* scratch, cudaStream_t s, int64_t* launches) {
  const bool nerfacto = P0.which == 1;
  const int width = nerfacto ? 32 : 10, hidden = nerfacto ? 64 : 16, k_w = nerfacto ? 32 : 16;
  for (int64_t r0 = 0; r0 < P0.n; r0 += kFieldBwdBlock) {
    const int64_t n = P0.n - r0 < kFieldBwdBlock ? P0.n - r0 : kFieldBwdBlock;
    FieldBwdParams P = P0;
    P.xyz += 3 * r0;
    if (P.dirs) P.dirs += 3 * r0;
    if (P.d_density) P.d_density += r0;
    if (P.d_rgb) P.d_rgb += 3 * r0;
    P.n = n;
    // carve: floats first (alignment), then halfs
    float* f = reinterpret_cast<float*>(scratch);
    P.sel = f;            f += n;
    float* dens = f;      f += n;
    P.d_pre3 = f;         f += n * 4;
    P.d_g2 = f;           f += n * 64;
    P.d_g1 = f;           f += n * 64;
    P.d_hx = f;           f += n * 32;
    P.d_o = f;            f += n * 16;
    P.d_h1 = f;           f += n * 64;
    P.d_x = f;            f += n * 32;
    __half* h = reinterpret_cast<__half*>(f);
    P.x = h;              h += n * 32;
    P.h1 = h;             h += n * 64;
    P.o = h;              h += n * 16;
    P.hx = h;             h += n * 32;
    P.g1 = h;             h += n * 64;
    P.g2 = h;             h += n * 64;
    P.pre3 = h;           h += n * 16;
    __half* geo = h;
    // ---- forward recomputation with the component-query kernels (query.cu) -----------------------
    QueryParams Q;
    memset(&Q, 0, sizeof(Q));
    Q.xyz = P.xyz;
    Q.n = n;
    Q.grid[0] = P.grid;
    Q.n_grids = 1;
    Q.linf = 1;
    Q.selector = 1;
    Q.feat = P.x;
    Q.sel = P.sel;
    cudaError_t e = launch_encode(Q, s);
    if (e != cudaSuccess) return e;
    e = launch_dense(P.x, width, width, P.w1, k_w, 0.f, P.h1, hidden, hidden, 1, n, s);
    if (e != cudaSuccess) return e;
    e = launch_dense(P.h1, hidden, hidden, P.w2, hidden, 0.f, P.o, 16, 16, 0, n, s);
    if (e != cudaSuccess) return e;
    if (launches) *launches += 3;
    const bool colour = nerfacto && P.d_rgb != nullptr;
    if (colour) {
      e = launch_density_finish(P.o, 16, P.sel, dens, geo, 15, n, s);
      if (e != cudaSuccess) return e;
      e = launch_head_input(P.dirs, geo, P.hx, n, s);
      if (e != cudaSuccess) return e;
      e = launch_dense(P.hx, 32, 32, P.wh1, 32, 1.f, P.g1, 64, 64, 1, n, s);
      if (e != cudaSuccess) return e;
      e = launch_dense(P.g1, 64, 64, P.wh2, 64, 0.f, P.g2, 64, 64, 1, n, s);
      if (e != cudaSuccess) return e;
      e = launch_dense(P.g2, 64, 64, P.wh3, 64, 0.f, P.pre3, 16, 16, 0, n, s);
      if (e != cudaSuccess) return e;
      if (launches) *launches += 5;
    }
    DeviceExec ex{s};
    field_backward_chain(P, ex);
    if (launches) *launches += colour ? 13 : 6;
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

// ---------------------------------------------------------------------------------------------
// conv head
// ---------------------------------------------------------------------------------------------
size_t conv_bwd_scratch_bytes(int64_t n_patches) {
  const int64_t rows = (n_patches < kConvBwdBlockPatches ? n_patches : kConvBwdBlockPatches) * kConvPos;
  return static_cast<size_t>(rows) * (2 * kConvK * 2 + kConvC * 2 + kConvC * 4 + kConvK * 4 + kConvC * 4) + 256;
}

cudaError_t launch_conv_backward(const ConvBwdParams& P0, void* scratch, cudaStream_t s, int64_t* launches) {
  const int64_t block_rows = kConvBwdBlockPatches * kConvPos;
  for (int64_t r0 = 0; r0 < P0.rows; r0 += block_rows) {
    const int64_t n = P0.rows - r0 < block_rows ? P0.rows - r0 : block_rows;
    ConvBwdParams P = P0;
    P.feat_in += r0 * kConvC;
    P.d_out += (r0 / kConvPos) * kConvC;
    if (P.d_feat) P.d_feat += r0 * kConvC;
    P.rows = n;
    float* f = reinterpret_cast<float*>(scratch);
    P.d_y = f;      f += n * kConvC;
    P.d_xcol = f;   f += n * kConvK;
    P.d_hid = f;    f += n * kConvC;
    __half* h = reinterpret_cast<__half*>(f);
    P.xcol1 = h;    h += n * kConvK;
    P.xcol2 = h;    h += n * kConvK;
    P.hid = h;
    DeviceExec ex{s};
    conv_backward_chain(P, ex);
    if (launches) *launches += P.d_feat ? 12 : 10;
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

}  // namespace snrf
