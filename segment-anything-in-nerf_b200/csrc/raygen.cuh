// Camera ray generation fused in front of the render: one thread = one ray.
//
// Reference restated (paths relative to /root/reference):
//   Cameras.get_image_coords            nerfstudio/cameras/cameras.py:284-310  (pixel centre +0.5, (y, x) order)
//   Cameras._generate_rays_from_coords  nerfstudio/cameras/cameras.py:575-726  (pinhole / fisheye / equirectangular,
//                                       -z forward, y flipped, rotation by c2w, normalisation, pixel_area)
//   radial_and_tangential_undistort     nerfstudio/cameras/camera_utils.py:298-401 (10 Newton steps, eps 1e-3)
//   LOOP B sub-grid and ray order       samnerf/sam_model.py:368-379 (patch-major, row-major inside a patch)
//
// The body is host+device so that tests/emu can run it on the CPU against tests/golden/raygen.npz.
#pragma once
#include "common.cuh"

namespace snrf {

constexpr int kCamPerspective = 1, kCamFisheye = 2, kCamEquirect = 3;  // CameraType, cameras.py:42-47

struct CameraDev {
  float fx, fy, cx, cy;
  int type;      // kCam*
  int has_dist;  // distortion_params present (cameras.py:617-635)
  float dist[6]; // k1 k2 k3 k4 p1 p2
  float c2w[12]; // row-major 3x4 camera-to-world
};

// Ray / axis-aligned box intersection for the viewer's crop box (Cameras.generate_rays(aabb_box=...),
// cameras.py:463-482 -> nerfstudio/utils/math.py:201-238): slab test, both distances clamped to [0, 1e10], a miss
// reports 1e10 for both.  0/0 lanes (a ray inside a slab's plane with zero direction component) give NaN, which
// torch.min / max propagate; fminf / fmaxf would drop it, hence the explicit tests.
SNRF_HD void aabb_near_far(const float (&o)[3], const float (&d)[3], const float (&box)[6], float& t_near, float& t_far) {
  float lo = -INFINITY, hi = INFINITY;
  bool nan = false;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float t0 = (box[a] - o[a]) / d[a], t1 = (box[3 + a] - o[a]) / d[a];
    nan = nan || t0 != t0 || t1 != t1;
    lo = fmaxf(lo, fminf(t0, t1));
    hi = fminf(hi, fmaxf(t0, t1));
  }
  if (nan) lo = hi = NAN;
  lo = lo != lo ? lo : fminf(fmaxf(lo, 0.f), 1e10f);
  hi = hi != hi ? hi : fminf(fmaxf(hi, 0.f), 1e10f);
  const bool miss = hi <= lo;
  t_near = miss ? 1e10f : lo;
  t_far = miss ? 1e10f : hi;
}

struct RayGenParams {
  CameraDev cam;
  const int* rows;   // [n_rows] pixel rows, or null -> 0..n_rows-1
  const int* cols;   // [n_cols] pixel columns, or null -> 0..n_cols-1
  int n_rows, n_cols;
  int patch;         // p: rays are emitted patch-major over p x p blocks of the (rows x cols) grid; 1 = row-major
  float* origins;    // [n_rows*n_cols, 3]
  float* dirs;       // [n_rows*n_cols, 3]
  float* pixel_area; // [n_rows*n_cols] or null
  int has_aabb;      // crop box: also write nears / fars (cameras.py:463-482)
  float aabb[6];     // x_min y_min z_min x_max y_max z_max
  float* nears;      // [n_rows*n_cols] or null
  float* fars;       // [n_rows*n_cols] or null
};

SNRF_HD void undistort_rt(float& x, float& y, const float (&k)[6]) {
  const float xd = x, yd = y;
  const float k1 = k[0], k2 = k[1], k3 = k[2], k4 = k[3], p1 = k[4], p2 = k[5];
#pragma unroll 1
  for (int it = 0; it < 10; ++it) {
    const float r = x * x + y * y;
    const float d = 1.0f + r * (k1 + r * (k2 + r * (k3 + r * k4)));
    const float fx = d * x + 2.f * p1 * x * y + p2 * (r + 2.f * x * x) - xd;
    const float fy = d * y + 2.f * p2 * x * y + p1 * (r + 2.f * y * y) - yd;
    const float d_r = k1 + r * (2.0f * k2 + r * (3.0f * k3 + r * 4.0f * k4));
    const float d_x = 2.0f * x * d_r, d_y = 2.0f * y * d_r;
    const float fx_x = d + d_x * x + 2.0f * p1 * y + 6.0f * p2 * x;
    const float fx_y = d_y * x + 2.0f * p1 * x + 2.0f * p2 * y;
    const float fy_x = d_x * y + 2.0f * p2 * y + 2.0f * p1 * x;
    const float fy_y = d + d_y * y + 2.0f * p2 * x + 6.0f * p1 * y;
    const float den = fy_x * fx_y - fx_x * fy_y;
    const bool ok = fabsf(den) > 1e-3f;
    x += ok ? (fx * fy_y - fy * fx_y) / den : 0.f;
    y += ok ? (fy * fx_x - fx * fy_x) / den : 0.f;
  }
}

// world-space unit direction for normalised image coordinates (u, v)
SNRF_HD void camera_direction(const CameraDev& C, float u, float v, float (&out)[3]) {
  if (C.has_dist && C.type != kCamEquirect) undistort_rt(u, v, C.dist);
  float d0, d1, d2;
  if (C.type == kCamFisheye) {
    float theta = sqrtf(u * u + v * v);
    theta = fminf(fmaxf(theta, 0.f), 3.14159265358979323846f);
    const float s = sinf(theta);
    d0 = u * s / theta;
    d1 = v * s / theta;
    d2 = -cosf(theta);
  } else if (C.type == kCamEquirect) {
    const float theta = -3.14159265358979323846f * u;
    const float phi = 3.14159265358979323846f * (0.5f - v);
    const float sp = sinf(phi);
    d0 = -sinf(theta) * sp;
    d1 = cosf(phi);
    d2 = -cosf(theta) * sp;
  } else {
    d0 = u;
    d1 = v;
    d2 = -1.f;
  }
  float w[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) w[j] = d0 * C.c2w[4 * j + 0] + d1 * C.c2w[4 * j + 1] + d2 * C.c2w[4 * j + 2];
  const float norm = fmaxf(sqrtf(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]), 8.8817841970012523e-16f);
#pragma unroll
  for (int j = 0; j < 3; ++j) out[j] = w[j] / norm;
}

// pixel (row, col) indices of ray i
SNRF_HD void raygen_pixel(const RayGenParams& P, int64_t i, int& py, int& px) {
  int r, c;
  if (P.patch > 1) {
    const int p = P.patch, fw = P.n_cols / p;
    const int v = static_cast<int>(i % p);
    const int u = static_cast<int>((i / p) % p);
    const int b = static_cast<int>((i / (p * p)) % fw);
    const int a = static_cast<int>(i / (static_cast<int64_t>(p) * p * fw));
    r = a * p + u;
    c = b * p + v;
  } else {
    r = static_cast<int>(i / P.n_cols);
    c = static_cast<int>(i % P.n_cols);
  }
  py = P.rows ? P.rows[r] : r;
  px = P.cols ? P.cols[c] : c;
}

SNRF_HD void raygen_one(const RayGenParams& P, int64_t i) {
  const CameraDev& C = P.cam;
  int py, px;
  raygen_pixel(P, i, py, px);
  const float y = static_cast<float>(py) + 0.5f, x = static_cast<float>(px) + 0.5f;
  const float u = (x - C.cx) / C.fx, v = -(y - C.cy) / C.fy;
  float d[3];
  camera_direction(C, u, v, d);
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    P.origins[3 * i + j] = C.c2w[4 * j + 3];
    P.dirs[3 * i + j] = d[j];
  }
  if (P.has_aabb && P.nears && P.fars) {
    const float o[3] = {C.c2w[3], C.c2w[7], C.c2w[11]};
    aabb_near_far(o, d, P.aabb, P.nears[i], P.fars[i]);
  }
  if (P.pixel_area) {
    float dxo[3], dyo[3];
    camera_direction(C, (x - C.cx + 1.f) / C.fx, v, dxo);
    camera_direction(C, u, -(y - C.cy + 1.f) / C.fy, dyo);
    float sx = 0.f, sy = 0.f;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      sx += (d[j] - dxo[j]) * (d[j] - dxo[j]);
      sy += (d[j] - dyo[j]) * (d[j] - dyo[j]);
    }
    P.pixel_area[i] = sqrtf(sx) * sqrtf(sy);
  }
}

cudaError_t launch_raygen(const RayGenParams& P, cudaStream_t stream);

}  // namespace snrf
