"""CPU restatement of the reference's camera ray generation (torch fp32).

TEST INFRASTRUCTURE ONLY - the checker, never the product (see ``oracle/samnerf_oracle.py``).
Pinned by ``tests/golden/raygen.npz``, which ``oracle/make_raygen_golden.py`` produced with the reference's own
``Cameras.generate_rays``; ``tests/test_raygen.py`` holds this file to it.

Follows (paths relative to /root/reference):
  * ``Cameras.get_image_coords``            nerfstudio/cameras/cameras.py:284-310   (pixel centre + 0.5, stored (y, x))
  * ``Cameras._generate_rays_from_coords``  nerfstudio/cameras/cameras.py:490-726
  * ``radial_and_tangential_undistort``     nerfstudio/cameras/camera_utils.py:298-401 (10 Newton steps, eps 1e-3)
  * ``normalize_with_norm``                 nerfstudio/cameras/camera_utils.py:240-252
  * LOOP B pixel sub-grid and ray order     samnerf/sam_model.py:368-379
"""
from __future__ import annotations

import math
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

PERSPECTIVE, FISHEYE, EQUIRECTANGULAR = 1, 2, 3  # cameras.py:42-47 (Enum auto())
_EPS = float(np.finfo(float).eps * 4.0)  # camera_utils.py:28


def undistort(coords: torch.Tensor, dist: torch.Tensor, eps: float = 1e-3, max_iterations: int = 10) -> torch.Tensor:
    """camera_utils.py:364-401 with the residual / Jacobian of :298-360.  ``coords[...,2]``, ``dist[6]``."""
    k1, k2, k3, k4, p1, p2 = [dist[i] for i in range(6)]
    xd, yd = coords[..., 0], coords[..., 1]
    x, y = xd, yd
    for _ in range(max_iterations):
        r = x * x + y * y
        d = 1.0 + r * (k1 + r * (k2 + r * (k3 + r * k4)))
        fx = d * x + 2 * p1 * x * y + p2 * (r + 2 * x * x) - xd
        fy = d * y + 2 * p2 * x * y + p1 * (r + 2 * y * y) - yd
        d_r = k1 + r * (2.0 * k2 + r * (3.0 * k3 + r * 4.0 * k4))
        d_x, d_y = 2.0 * x * d_r, 2.0 * y * d_r
        fx_x = d + d_x * x + 2.0 * p1 * y + 6.0 * p2 * x
        fx_y = d_y * x + 2.0 * p1 * x + 2.0 * p2 * y
        fy_x = d_x * y + 2.0 * p2 * y + 2.0 * p1 * x
        fy_y = d + d_y * y + 2.0 * p2 * x + 6.0 * p1 * y
        den = fy_x * fx_y - fx_x * fy_y
        ok = den.abs() > eps
        x = x + torch.where(ok, (fx * fy_y - fy * fx_y) / den, torch.zeros_like(den))
        y = y + torch.where(ok, (fy * fx_x - fx * fy_x) / den, torch.zeros_like(den))
    return torch.stack([x, y], dim=-1)


def generate_rays(
    fx: float, fy: float, cx: float, cy: float, c2w: torch.Tensor, ys: torch.Tensor, xs: torch.Tensor,
    camera_type: int = PERSPECTIVE, dist: Optional[Sequence[float]] = None,
) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Rays through pixels ``(ys, xs)`` (integer pixel indices, any common shape).  Returns origins ``[...,3]``,
    unit directions ``[...,3]`` and ``pixel_area[...,1]`` (cameras.py:575-726)."""
    c2w = c2w.to(torch.float32)
    y = ys.to(torch.float32) + 0.5
    x = xs.to(torch.float32) + 0.5
    fx, fy, cx, cy = (torch.tensor(v, dtype=torch.float32) for v in (fx, fy, cx, cy))
    coord = torch.stack([(x - cx) / fx, -(y - cy) / fy], -1)
    coord_x = torch.stack([(x - cx + 1) / fx, -(y - cy) / fy], -1)
    coord_y = torch.stack([(x - cx) / fx, -(y - cy + 1) / fy], -1)
    stack = torch.stack([coord, coord_x, coord_y], dim=0)
    if dist is not None and len(dist) and camera_type != EQUIRECTANGULAR:
        stack = undistort(stack, torch.tensor(list(dist), dtype=torch.float32))
    if camera_type == PERSPECTIVE:
        d = torch.stack([stack[..., 0], stack[..., 1], -torch.ones_like(stack[..., 0])], -1)
    elif camera_type == FISHEYE:
        theta = torch.clip(torch.sqrt(torch.sum(stack**2, dim=-1)), 0.0, math.pi)
        s = torch.sin(theta)
        d = torch.stack([stack[..., 0] * s / theta, stack[..., 1] * s / theta, -torch.cos(theta)], -1)
    elif camera_type == EQUIRECTANGULAR:
        theta = -torch.pi * stack[..., 0]
        phi = torch.pi * (0.5 - stack[..., 1])
        d = torch.stack([-torch.sin(theta) * torch.sin(phi), torch.cos(phi), -torch.cos(theta) * torch.sin(phi)], -1)
    else:
        raise ValueError(f"Camera type {camera_type} not supported.")
    rot = c2w[:3, :3]
    d = torch.sum(d[..., None, :] * rot, dim=-1)
    norm = torch.maximum(torch.linalg.vector_norm(d, dim=-1, keepdims=True), torch.tensor([_EPS]).to(d))
    d = d / norm
    dirs = d[0]
    dx = torch.sqrt(torch.sum((dirs - d[1]) ** 2, dim=-1))
    dy = torch.sqrt(torch.sum((dirs - d[2]) ** 2, dim=-1))
    origins = c2w[:3, 3].expand(dirs.shape)
    return origins, dirs, (dx * dy)[..., None]


def full_image_pixels(h: int, w: int) -> Tuple[torch.Tensor, torch.Tensor]:
    ys, xs = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
    return ys, xs


def feature_grid_pixels(h: int, w: int, fh: int, fw: int, p: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """sam_model.py:372-379: ``linspace(0, H-1, fh*p).long()`` x ``linspace(0, W-1, fw*p).long()``, reshaped
    ``(fh, p, fw, p)``, transposed to ``(fh, fw, p, p)`` and flattened: patch-major, row-major inside a patch."""
    hi = torch.linspace(0, h - 1, fh * p, dtype=torch.long)
    wi = torch.linspace(0, w - 1, fw * p, dtype=torch.long)
    hind, wind = torch.meshgrid(hi, wi, indexing="ij")
    hind = hind.reshape(fh, p, fw, p).transpose(1, 2).flatten()
    wind = wind.reshape(fh, p, fw, p).transpose(1, 2).flatten()
    return hind, wind


def intersect_aabb(origins: torch.Tensor, directions: torch.Tensor, aabb: torch.Tensor, max_bound: float = 1e10,
                   invalid_value: float = 1e10) -> Tuple[torch.Tensor, torch.Tensor]:
    """``_intersect_aabb`` - nerfstudio/utils/math.py:201-238 (the path taken when nerfacc is absent, :260-270):
    slab test, distances clamped to ``[0, max_bound]``, a miss reports ``invalid_value`` twice.  ``aabb[6]`` =
    ``[x_min y_min z_min x_max y_max z_max]``; used by ``Cameras.generate_rays(aabb_box=...)`` (cameras.py:463-482)."""
    tx_min = (aabb[:3] - origins) / directions
    tx_max = (aabb[3:] - origins) / directions
    t_min = torch.max(torch.min(tx_min, tx_max), dim=-1).values
    t_max = torch.min(torch.max(tx_min, tx_max), dim=-1).values
    t_min = torch.clamp(t_min, min=0, max=max_bound)
    t_max = torch.clamp(t_max, min=0, max=max_bound)
    cond = t_max <= t_min
    return torch.where(cond, invalid_value, t_min), torch.where(cond, invalid_value, t_max)
