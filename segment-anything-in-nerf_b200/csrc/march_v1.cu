// Kernel A, ROUND-1 VERSION (kept for A/B measurements against march.cu; selected with SNRF_MARCH=v1) - the per-ray march: proposal sampling -> proposal density -> weights -> PDF resample ->
// nerfacto field (hash grid + base MLP + SH + colour head) -> weights -> RGB / median depth / accumulation
// -> top-k + sharpening of the feature samples.  One warp owns one ray from start to finish; every
// per-ray intermediate lives in registers or in a 2.6 KB per-warp shared-memory scratch.
//
// Reference path restated (paths relative to /root/reference):
//   NearFarCollider                 nerfstudio/model_components/scene_colliders.py:183-188
//   UniformLinDispPiecewiseSampler  nerfstudio/model_components/ray_samplers.py:79-126,223-246
//   HashMLPDensityField.get_density nerfstudio/fields/density_fields.py:102-125
//   RaySamples.get_weights          nerfstudio/cameras/rays.py:141-163
//   PDFSampler (eval)               nerfstudio/model_components/ray_samplers.py:274-369
//   TCNNNerfactoField               nerfstudio/fields/nerfacto_field.py:242-351
//   RGB/Depth/Accumulation          nerfstudio/model_components/renderers.py:69-140,197-223,260-270
//   top-k + sharpen                 samnerf/sam_model.py:243-255
//
// Thread mapping: lane = (s16, xb).  s16 = lane>>1 picks one of 16 samples of the current tile, xb = lane&1
// picks the x-neighbour.  The two x-corners of a voxel are adjacent in memory for dense levels and, for hashed
// levels, whenever gx is even (the x prime is 1), so the lane pair's two 4-byte loads fall into one 32-byte
// sector and one L1 wavefront.  The MLPs run as mma.sync m16n8k16 tiles with register-resident activations.
#include "kernels.cuh"

namespace snrf {

namespace {

#ifndef SNRF_MARCH_MIN_CTAS
#define SNRF_MARCH_MIN_CTAS 3
#endif
constexpr int kWarpsPerCta = 8;
constexpr int kSP = 64;  // proposal samples per ray
constexpr int kSN = 32;  // nerf samples per ray

// per-warp scratch (bytes)
struct alignas(16) WarpScratch {
  uint4 a_tile[80];     // 16 rows x 80 B (64 B of data + 16 B pad: conflict-free ldmatrix)
  float cdf[68];        // 65 used
  float w0[64];         // proposal weights
  float t1[36];         // 33 nerf bin edges (euclidean)
  float dens[32];       // density pre-activation (fp16-rounded) per nerf sample
  float sel[32];        // (0,1) selector per nerf sample
  float rgb[96];        // per-sample rgb
};

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// F = 2 gather for the sample this lane pair owns.  p[l][f] = this lane's x-half of the trilinear sum.
template <int NL, uint32_t MASK>
__device__ __forceinline__ void gather_f2(const GridDev& G, float x, float y, float z, int xb, float (&p)[NL][2]) {
#pragma unroll
  for (int l = 0; l < NL; ++l) {
    const float scale = G.lv[l].scale;
    const float px = __fadd_rn(__fmul_rn(x, scale), 0.5f);
    const float py = __fadd_rn(__fmul_rn(y, scale), 0.5f);
    const float pz = __fadd_rn(__fmul_rn(z, scale), 0.5f);
    const float fx = floorf(px), fy = floorf(py), fz = floorf(pz);
    const float rx = px - fx, ry = py - fy, rz = pz - fz;
    const uint32_t gx = static_cast<uint32_t>(static_cast<int>(fx)) + xb;
    const uint32_t gy = static_cast<uint32_t>(static_cast<int>(fy));
    const uint32_t gz = static_cast<uint32_t>(static_cast<int>(fz));
    const float wx = xb ? rx : 1.f - rx;
    uint32_t idx[4], v[4];
    corner_indices(G.lv[l], level_hashed<MASK>(G.lv[l], l), gx, gy, gz, idx);
#pragma unroll
    for (int c = 0; c < 4; ++c) v[c] = ldg_u32(G.table + 2 * static_cast<size_t>(idx[c]));
    float a0 = 0.f, a1 = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float w = wx * ((c & 1) ? ry : 1.f - ry);
      w *= ((c >> 1) ? rz : 1.f - rz);
      const float2 f = h2_to_f2(v[c]);
      a0 += w * f.x;
      a1 += w * f.y;
    }
    p[l][0] = a0;
    p[l][1] = a1;
  }
}

__device__ __forceinline__ void relu_pack(const float (&acc)[8][4], uint32_t (&a)[4][4]) {
#pragma unroll
  for (int kb = 0; kb < 4; ++kb) {
    a[kb][0] = f2_to_h2(fmaxf(acc[2 * kb][0], 0.f), fmaxf(acc[2 * kb][1], 0.f));
    a[kb][1] = f2_to_h2(fmaxf(acc[2 * kb][2], 0.f), fmaxf(acc[2 * kb][3], 0.f));
    a[kb][2] = f2_to_h2(fmaxf(acc[2 * kb + 1][0], 0.f), fmaxf(acc[2 * kb + 1][1], 0.f));
    a[kb][3] = f2_to_h2(fmaxf(acc[2 * kb + 1][2], 0.f), fmaxf(acc[2 * kb + 1][3], 0.f));
  }
}

template <int NT, int KS>
__device__ __forceinline__ void mlp_layer(float (&acc)[NT][4], const uint32_t (*a)[4], const uint2* __restrict__ wf,
                                          int lane) {
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      const uint2 b = wf[(nt * KS + ks) * 32 + lane];
      mma_16816(acc[nt], a[ks], b.x, b.y);
    }
  }
}

}  // namespace

// PM / FM: hashed-level masks of the proposal / nerfacto grids (kRuntimeMask = read them from the descriptor)
// ET: early termination (opt-in, snrf_set_early_termination): when the transmittance left after the first 16 nerf
// samples is below P.et_eps, the second tile's field evaluation (half of the nerfacto gathers and MLP work) is
// skipped and its samples get weight 0 - they could have moved rgb / accumulation by at most et_eps, cannot hold the
// median (cumulative weight >= 1 - et_eps > 0.5 is reached inside the first tile) and, after the w^10 sharpening,
// cannot carry feature weight.  ET = false is the exact path and compiles to the same code as before.
// JIT: training-mode stratified sampling with one random number per ray and level (single jitter,
// ray_samplers.py:104-112,314-322; nerfacto.py:113,211): P.jitter[ray] = {t_rand of the initial sampler, rand of the
// PDF sampler}, drawn by the caller (torch.rand in the reference).  JIT = false is the eval path, unchanged.
template <uint32_t PM, uint32_t FM, bool ET, bool JIT>
__global__ void __launch_bounds__(kWarpsPerCta * 32, SNRF_MARCH_MIN_CTAS) march_v1_kernel(const MarchParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // layout: [wfrag kMarchFragTiles*256 B][WarpScratch x warps]
  uint2* s_wf = reinterpret_cast<uint2*>(smem_raw);
  WarpScratch* s_ws = reinterpret_cast<WarpScratch*>(smem_raw + kMarchFragTiles * 256);

  for (int i = threadIdx.x; i < kMarchFragTiles * 32; i += blockDim.x) s_wf[i] = P.wfrag[i];
  __syncthreads();

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int s16 = lane >> 1, xb = lane & 1;
  const int g = lane >> 2, q = lane & 3;
  WarpScratch& ws = s_ws[warp];
  const unsigned FULL = 0xffffffffu;

  const int64_t warps_total = static_cast<int64_t>(gridDim.x) * kWarpsPerCta;
  for (int64_t ray = static_cast<int64_t>(blockIdx.x) * kWarpsPerCta + warp; ray < P.n_rays; ray += warps_total) {
    const float ox = P.origins[3 * ray + 0], oy = P.origins[3 * ray + 1], oz = P.origins[3 * ray + 2];
    const float dx = P.dirs[3 * ray + 0], dy = P.dirs[3 * ray + 1], dz = P.dirs[3 * ray + 2];
    const float near = P.nears ? P.nears[ray] : P.near_default;
    const float far = P.fars ? P.fars[ray] : P.far_default;
    const float s_near = spacing_fn(near), s_far = spacing_fn(far);
    float jit0 = 0.f, jit1 = 0.f;
    if (JIT) {
      jit0 = P.jitter[2 * ray];
      jit1 = P.jitter[2 * ray + 1];
    }
    // spacing-space bin edge j of 65: linspace(0,1,65)[j] = j/64 exactly; with jitter see jittered_bin (common.cuh)
    auto bin0 = [&](int j) -> float {
      return JIT ? jittered_bin(j, kSP, jit0) : static_cast<float>(j) * (1.f / kSP);
    };
    auto edge0 = [&](int j) -> float {  // proposal bin edge j of 65 in euclidean t
      const float b = bin0(j);
      return spacing_fn_inv(__fadd_rn(__fmul_rn(b, s_far), __fmul_rn(1.f - b, s_near)));
    };

    // ---------------- proposal density + weights: 4 tiles of 16 samples --------------------------
    float carry = 0.f;  // running sum of delta*sigma over earlier tiles
#pragma unroll 1
    for (int tile = 0; tile < kSP / 16; ++tile) {
      const int j = tile * 16 + s16;
      const float ts = edge0(j), te = edge0(j + 1);
      const float tm2 = ts + te;
      const float px = __fadd_rn(ox, __fmul_rn(dx, tm2) / 2.f);
      const float py = __fadd_rn(oy, __fmul_rn(dy, tm2) / 2.f);
      const float pz = __fadd_rn(oz, __fmul_rn(dz, tm2) / 2.f);
      float x, y, z, sel;
      contract_normalize(px, py, pz, true, true, x, y, z, sel);
      float p[5][2];
      gather_f2<5, PM>(P.prop, x, y, z, xb, p);
      // finish the x-pair sums: lane xb=0 takes levels 0-3 (columns 0-7), lane xb=1 level 4 plus the zero padding
      // tcnn appends to reach width 16; one 16-byte store each into the warp's activation tile
      {
        uint32_t pk[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float s0 = xb ? p[i][0] : p[4][0], s1 = xb ? p[i][1] : p[4][1];
          const float m0 = xb ? p[4][0] : p[i][0], m1 = xb ? p[4][1] : p[i][1];
          const float r0 = __shfl_xor_sync(FULL, s0, 1), r1 = __shfl_xor_sync(FULL, s1, 1);
          pk[i] = f2_to_h2(m0 + r0, m1 + r1);
        }
        if (xb) pk[1] = pk[2] = pk[3] = 0u;
        ws.a_tile[s16 * 5 + xb] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
      __syncwarp();
      // 16 -> 16 -> 1 MLP as three mma.sync tiles (fp16 operands, fp32 accumulate, fp16 hidden like tcnn)
      uint32_t a_p[1][4];
      ldmatrix_x4(a_p[0], smem_u32(ws.a_tile) + ((lane & 7) + ((lane >> 3) & 1) * 8) * 80 + (lane >> 4) * 16);
      float acc_h[2][4];
      mlp_layer<2, 1>(acc_h, a_p, s_wf + kFragProp1 * 32, lane);
      uint32_t a_q[1][4];
      a_q[0][0] = f2_to_h2(fmaxf(acc_h[0][0], 0.f), fmaxf(acc_h[0][1], 0.f));
      a_q[0][1] = f2_to_h2(fmaxf(acc_h[0][2], 0.f), fmaxf(acc_h[0][3], 0.f));
      a_q[0][2] = f2_to_h2(fmaxf(acc_h[1][0], 0.f), fmaxf(acc_h[1][1], 0.f));
      a_q[0][3] = f2_to_h2(fmaxf(acc_h[1][2], 0.f), fmaxf(acc_h[1][3], 0.f));
      float acc_o[1][4];
      mlp_layer<1, 1>(acc_o, a_q, s_wf + kFragProp2 * 32, lane);
      // output column 0 (the density) of row r sits in lane 4*(r&7): register 0 for rows 0-7, register 2 for rows 8-15
      const float v_lo = __shfl_sync(FULL, acc_o[0][0], 4 * (s16 & 7));
      const float v_hi = __shfl_sync(FULL, acc_o[0][2], 4 * (s16 & 7));
      const float h = round_f16(s16 < 8 ? v_lo : v_hi);
      const float sigma = expf(h) * sel;
      const float ds = (te - ts) * sigma;
      // inclusive scan over the 16 samples (values are duplicated in each lane pair)
      float incl = ds;
#pragma unroll
      for (int o = 2; o < 32; o <<= 1) {
        const float nb = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl += nb;
      }
      float excl = __shfl_up_sync(FULL, incl, 2);
      if (lane < 2) excl = 0.f;
      const float trans = expf(-(carry + excl));
      const float w = nan_to_num((1.f - expf(-ds)) * trans);
      if (xb == 0) ws.w0[j] = w;
      carry += __shfl_sync(FULL, incl, 31);
    }
    __syncwarp();

    // ---------------- proposal median depth + PDF resample (lane i owns bins i and i+32) ----------
    {
      const float wa = ws.w0[lane], wb = ws.w0[lane + 32];
      if (P.dbg_w0) {
        P.dbg_w0[ray * kSP + lane] = wa;
        P.dbg_w0[ray * kSP + lane + 32] = wb;
      }
      const float ca = warp_incl_scan(wa, lane);
      const float cb = warp_incl_scan(wb, lane) + __shfl_sync(FULL, ca, 31);
      if (P.prop_depth) {
        const unsigned ba = __ballot_sync(FULL, ca >= 0.5f), bb = __ballot_sync(FULL, cb >= 0.5f);
        const int idx = ba ? (__ffs(ba) - 1) : (bb ? 32 + __ffs(bb) - 1 : kSP - 1);
        if (lane == 0) store_rep(P.prop_depth, P.rep[3], ray, (edge0(idx) + edge0(idx + 1)) / 2.f);
      }
      // proposal-weight annealing before the PDF resample (ray_samplers.py:583); training instantiation only -
      // the reported weights and the proposal depth above use the raw weights
      const float za = (JIT && P.anneal != 1.f) ? powf(wa, P.anneal) : wa;
      const float zb = (JIT && P.anneal != 1.f) ? powf(wb, P.anneal) : wb;
      float pa = za + P.hist_padding, pb = zb + P.hist_padding;
      float sum = warp_sum(pa + pb);
      const float padding = fmaxf(1e-5f - sum, 0.f);
      pa += padding / kSP;
      pb += padding / kSP;
      sum += padding;
      pa /= sum;
      pb /= sum;
      const float ia = warp_incl_scan(pa, lane);
      const float ib = warp_incl_scan(pb, lane) + __shfl_sync(FULL, ia, 31);
      if (lane == 0) ws.cdf[0] = 0.f;
      ws.cdf[lane + 1] = fminf(1.f, ia);
      ws.cdf[lane + 33] = fminf(1.f, ib);
      __syncwarp();
      for (int j = lane; j < kSN + 1; j += 32) {
        const float u = JIT ? P.pdf_u_base[j] + jit1 / static_cast<float>(kSN + 1) : P.pdf_u[j];
        int lo = 0, hi = kSP + 1;  // searchsorted(cdf, u, side="right"): number of entries <= u
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (ws.cdf[mid] <= u) lo = mid + 1; else hi = mid;
        }
        const int below = min(max(lo - 1, 0), kSP), above = min(max(lo, 0), kSP);
        const float c0 = ws.cdf[below], c1 = ws.cdf[above];
        float t = nan_to_num((u - c0) / (c1 - c0));
        t = fminf(fmaxf(t, 0.f), 1.f);
        const float b0 = JIT ? bin0(below) : below * (1.f / kSP), b1 = JIT ? bin0(above) : above * (1.f / kSP);
        const float bin = __fadd_rn(b0, __fmul_rn(t, b1 - b0));
        const float e = spacing_fn_inv(__fadd_rn(__fmul_rn(bin, s_far), __fmul_rn(1.f - bin, s_near)));
        ws.t1[j] = e;
        if (P.dbg_edges) P.dbg_edges[ray * (kSN + 1) + j] = e;
      }
      __syncwarp();
    }
    if (P.flags & kFlagSamplesOnly) continue;

    // ---------------- nerfacto field: 2 tiles of 16 samples -------------------------------------
    // SH(4) of the ray direction, packed as the A fragment of the colour head's second k-block
    uint32_t sh_lo, sh_hi;
    {
      const float sx = ((dx + 1.f) / 2.f) * 2.f - 1.f, sy = ((dy + 1.f) / 2.f) * 2.f - 1.f,
                  sz = ((dz + 1.f) / 2.f) * 2.f - 1.f;
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
      uint32_t pk[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) pk[i] = f2_to_h2(sh[2 * i], sh[2 * i + 1]);
      sh_lo = q == 0 ? pk[0] : q == 1 ? pk[1] : q == 2 ? pk[2] : pk[3];
      sh_hi = q == 0 ? pk[4] : q == 1 ? pk[5] : q == 2 ? pk[6] : pk[7];
    }

#pragma unroll 1
    for (int tile = 0; tile < kSN / 16; ++tile) {
      const int j = tile * 16 + s16;
      const float ts = ws.t1[j], te = ws.t1[j + 1];
      const float tm2 = ts + te;
      const float px = __fadd_rn(ox, __fmul_rn(dx, tm2) / 2.f);
      const float py = __fadd_rn(oy, __fmul_rn(dy, tm2) / 2.f);
      const float pz = __fadd_rn(oz, __fmul_rn(dz, tm2) / 2.f);
      float x, y, z, sel;
      contract_normalize(px, py, pz, true, true, x, y, z, sel);
      if (xb == 0) ws.sel[j] = sel;
      {
        float p[16][2];
        gather_f2<16, FM>(P.field, x, y, z, xb, p);
        // lane xb=0 finishes levels 0-7 (k-block 0), lane xb=1 levels 8-15 (k-block 1)
        uint32_t pk[8];
#pragma unroll
        for (int l = 0; l < 8; ++l) {
          const float s0 = xb ? p[l][0] : p[l + 8][0], s1 = xb ? p[l][1] : p[l + 8][1];
          const float m0 = xb ? p[l + 8][0] : p[l][0], m1 = xb ? p[l + 8][1] : p[l][1];
          const float r0 = __shfl_xor_sync(FULL, s0, 1), r1 = __shfl_xor_sync(FULL, s1, 1);
          pk[l] = f2_to_h2(m0 + r0, m1 + r1);
        }
        uint4* row = ws.a_tile + s16 * 5 + xb * 2;
        row[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        row[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
      }
      __syncwarp();
      uint32_t a_in[2][4];
      {
        const uint32_t base = smem_u32(ws.a_tile) + ((lane & 7) + ((lane >> 3) & 1) * 8) * 80 + (lane >> 4) * 16;
        ldmatrix_x4(a_in[0], base);
        ldmatrix_x4(a_in[1], base + 32);
      }
      float acc[8][4];
      uint32_t a_h[4][4];
      // base MLP 32 -> 64 -> 16
      mlp_layer<8, 2>(acc, a_in, s_wf + kFragBase1 * 32, lane);
      relu_pack(acc, a_h);
      float acc2[2][4];
      mlp_layer<2, 4>(acc2, a_h, s_wf + kFragBase2 * 32, lane);
      if (q == 0) {
        ws.dens[tile * 16 + g] = round_f16(acc2[0][0]);
        ws.dens[tile * 16 + g + 8] = round_f16(acc2[0][2]);
      }
      // colour head input: k-block 0 = [pad(=1), geo 0..14] (weights permuted at pack time), k-block 1 = SH
      uint32_t a_c[2][4];
      a_c[0][0] = f2_to_h2(q == 0 ? 1.f : acc2[0][0], acc2[0][1]);
      a_c[0][1] = f2_to_h2(q == 0 ? 1.f : acc2[0][2], acc2[0][3]);
      a_c[0][2] = f2_to_h2(acc2[1][0], acc2[1][1]);
      a_c[0][3] = f2_to_h2(acc2[1][2], acc2[1][3]);
      a_c[1][0] = sh_lo;
      a_c[1][1] = sh_lo;
      a_c[1][2] = sh_hi;
      a_c[1][3] = sh_hi;
      mlp_layer<8, 2>(acc, a_c, s_wf + kFragHead1 * 32, lane);
      relu_pack(acc, a_h);
      mlp_layer<8, 4>(acc, a_h, s_wf + kFragHead2 * 32, lane);
      relu_pack(acc, a_h);
      float acc3[1][4];
      mlp_layer<1, 4>(acc3, a_h, s_wf + kFragHead3 * 32, lane);
      if (q < 2) {
        const float v0 = round_f16(sigmoidf_(round_f16(acc3[0][0]))), v1 = round_f16(sigmoidf_(round_f16(acc3[0][1])));
        const float v2 = round_f16(sigmoidf_(round_f16(acc3[0][2]))), v3 = round_f16(sigmoidf_(round_f16(acc3[0][3])));
        float* r0 = ws.rgb + (tile * 16 + g) * 3;
        float* r1 = ws.rgb + (tile * 16 + g + 8) * 3;
        if (q == 0) {
          r0[0] = v0; r0[1] = v1; r1[0] = v2; r1[1] = v3;
        } else {
          r0[2] = v0; r1[2] = v2;
        }
      }
      __syncwarp();
      if (ET && tile == 0) {
        // transmittance after the first tile: exp(-sum_{j<16} delta_j * sigma_j)
        float ds = 0.f;
        if (lane < 16) ds = (ws.t1[lane + 1] - ws.t1[lane]) * (expf(ws.dens[lane]) * ws.sel[lane]);
        const float od = warp_sum(ds);
        if (expf(-od) < P.et_eps) {
          // samples 16..31: density 0 (-inf pre-activation), colour 0 -> weight exactly 0
          if (lane < 16) {
            ws.dens[16 + lane] = -INFINITY;
            ws.sel[16 + lane] = 0.f;
          }
          for (int i = lane; i < 48; i += 32) ws.rgb[48 + i] = 0.f;
          __syncwarp();
          break;
        }
      }
    }

    // ---------------- compositing: lane = sample ------------------------------------------------
    {
      const float ts = ws.t1[lane], te = ws.t1[lane + 1];
      const float sigma = expf(ws.dens[lane]) * ws.sel[lane];
      const float ds = (te - ts) * sigma;
      const float incl = warp_incl_scan(ds, lane);
      float excl = __shfl_up_sync(FULL, incl, 1);
      if (lane == 0) excl = 0.f;
      const float w = nan_to_num((1.f - expf(-ds)) * expf(-excl));
      if (P.dbg_weights) P.dbg_weights[ray * kSN + lane] = w;
      if (P.dbg_density) P.dbg_density[ray * kSN + lane] = sigma;
      const float accw = warp_sum(w);
      float cr = nan_to_num(ws.rgb[lane * 3 + 0]), cg = nan_to_num(ws.rgb[lane * 3 + 1]),
            cb = nan_to_num(ws.rgb[lane * 3 + 2]);
      if (P.dbg_rgb) {
        P.dbg_rgb[(ray * kSN + lane) * 3 + 0] = cr;
        P.dbg_rgb[(ray * kSN + lane) * 3 + 1] = cg;
        P.dbg_rgb[(ray * kSN + lane) * 3 + 2] = cb;
      }
      float sr = warp_sum(w * cr), sg = warp_sum(w * cg), sb = warp_sum(w * cb);
      float bgr, bgg, bgb;
      if (P.bg_mode == kBgLastSample) {
        bgr = __shfl_sync(FULL, cr, 31); bgg = __shfl_sync(FULL, cg, 31); bgb = __shfl_sync(FULL, cb, 31);
      } else {
        bgr = P.bg[0]; bgg = P.bg[1]; bgb = P.bg[2];
      }
      const float cw = __shfl_sync(FULL, warp_incl_scan(w, lane), lane);  // inclusive cumsum of weights
      const unsigned bm = __ballot_sync(FULL, cw >= 0.5f);
      const int mi = bm ? (__ffs(bm) - 1) : kSN - 1;
      if (lane == 0) {
        const float om = 1.f - accw;
        store_rep(P.rgb, P.rep[0], 3 * ray + 0, fminf(fmaxf(sr + bgr * om, 0.f), 1.f));
        store_rep(P.rgb, P.rep[0], 3 * ray + 1, fminf(fmaxf(sg + bgg * om, 0.f), 1.f));
        store_rep(P.rgb, P.rep[0], 3 * ray + 2, fminf(fmaxf(sb + bgb * om, 0.f), 1.f));
        store_rep(P.depth, P.rep[1], ray, (ws.t1[mi] + ws.t1[mi + 1]) / 2.f);
        if (P.acc) store_rep(P.acc, P.rep[2], ray, accw);
      }
      // top-k by weight (ties broken by sample index), sharpen, renormalise
      if (P.sam_t) {
        int rank = 0;
#pragma unroll
        for (int o = 0; o < 32; ++o) {
          const float wo = __shfl_sync(FULL, w, o);
          rank += (wo > w) || (wo == w && o < lane);
        }
        const bool pick = rank < P.k_sam;
        const float sw = pick ? powf(w, P.sharpen) : 0.f;
        const float tot = warp_sum(sw);
        if (pick) {
          P.sam_t[ray * P.k_sam + rank] = ts + te;  // 2 x midpoint: kernel B rebuilds pos = o + d*(ts+te)/2
          P.sam_w[ray * P.k_sam + rank] = sw / tot;
        }
      }
    }
    __syncwarp();
  }
}

static size_t march_smem_bytes() {
  return kMarchFragTiles * 256 + kWarpsPerCta * sizeof(WarpScratch);
}

// hashed-level masks of the shipped configs (samconfigs.py): proposal 16..128 over 5 levels at T=2^17 -> levels 3-4;
// nerfacto 16..2048 over 16 levels at T=2^19 -> levels 5-15
constexpr uint32_t kPropMaskStd = 0x18u, kFieldMaskStd = 0xFFE0u;

cudaError_t launch_march_v1(const MarchParams& P, int sm_count, cudaStream_t stream) {
  // function attributes are per device: remember which devices of this process have been configured
  static bool configured_dev[64] = {false};
  int dev_id = 0;
  if (cudaGetDevice(&dev_id) != cudaSuccess || dev_id < 0 || dev_id >= 64) return cudaErrorInvalidDevice;
  bool& configured = configured_dev[dev_id];
  const size_t smem = march_smem_bytes();
  auto* k_std = march_v1_kernel<kPropMaskStd, kFieldMaskStd, false, false>;
  auto* k_any = march_v1_kernel<kRuntimeMask, kRuntimeMask, false, false>;
  auto* k_std_et = march_v1_kernel<kPropMaskStd, kFieldMaskStd, true, false>;
  auto* k_any_et = march_v1_kernel<kRuntimeMask, kRuntimeMask, true, false>;
  auto* k_std_jit = march_v1_kernel<kPropMaskStd, kFieldMaskStd, false, true>;
  auto* k_any_jit = march_v1_kernel<kRuntimeMask, kRuntimeMask, false, true>;
  if (!configured) {
    for (auto* k : {k_std, k_any, k_std_et, k_any_et, k_std_jit, k_any_jit}) {
      cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
    }
    configured = true;
  }
  if (P.n_rays <= 0) return cudaSuccess;
  const int64_t ctas_needed = (P.n_rays + kWarpsPerCta - 1) / kWarpsPerCta;
  // persistent grid: a multiple of the SM count (resident CTAs per SM x 4 waves of work-striding)
  const int64_t cap = static_cast<int64_t>(sm_count) * SNRF_MARCH_MIN_CTAS * 4;
  const int grid = static_cast<int>(ctas_needed < cap ? ctas_needed : cap);
  const bool std_cfg = hashed_mask(P.prop) == kPropMaskStd && (hashed_mask(P.field) == kFieldMaskStd || (P.flags & kFlagSamplesOnly));
  const bool jit = P.jitter != nullptr;  // training-mode sampling: exact arithmetic only (no early termination)
  const bool et = !jit && P.et_eps > 0.f && !(P.flags & kFlagSamplesOnly);
  auto* k = std_cfg ? (jit ? k_std_jit : et ? k_std_et : k_std) : (jit ? k_any_jit : et ? k_any_et : k_any);
  k<<<grid, kWarpsPerCta * 32, smem, stream>>>(P);
  return cudaGetLastError();
}

}  // namespace snrf
