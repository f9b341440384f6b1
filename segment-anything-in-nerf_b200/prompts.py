"""Host-side glue between the rendered maps and the 2-D decoders that consume them (first
slice): prompt lifting / projection and the packing of the rendered feature maps.  A handful of points and two small
reshapes per frame - host logic, no kernels; the SAM mask decoder and the ClipSeg decoder themselves (2-D transformer
networks with their own checkpoints) stay out of scope.

Restated from (paths relative to /root/reference):
  * ``project``                          samnerf/sam_model.py:95-123
  * click -> 3-D prompt                  samnerf/sam_model.py:441-462   (depth minus ``TOR`` along the pixel's ray)
  * prompts inside the image             samnerf/sam_model.py:468-475
  * visibility test of ``show_prompts``  samnerf/sam_model.py:48-81     (``EPS`` slack against the rendered depth)
  * ``SamPredictor.set_feature`` padding samnerf/segment_anything/predictor.py:100-127
  * ClipSeg activations from the map     samnerf/sam_model.py:487-497
"""
from __future__ import annotations

import math
from typing import List, Tuple

import torch

EPS = 1e-4  # sam_model.py:36
TOR = 1e-2  # sam_model.py:37


def project(intrin: torch.Tensor, c2w: torch.Tensor, points: torch.Tensor) -> torch.Tensor:
    """World points ``[n,3]`` -> integer pixel coordinates ``[n,2]`` (x, y) of a pinhole camera (-z forward, y up)."""
    fx, fy, cx, cy = intrin[0, 0], intrin[1, 1], intrin[0, 2], intrin[1, 2]
    if c2w.shape[0] == 3:
        c2w = torch.cat([c2w, torch.tensor([[0.0, 0.0, 0.0, 1.0]]).to(c2w)], dim=0)
    if points.shape[-1] == 3:
        points = torch.cat([points, torch.tensor([[1.0]] * points.shape[0]).to(points)], dim=-1)
    w2c = torch.inverse(c2w)[:3]
    img = torch.einsum("ij,bj->bi", w2c, points)
    img = -img / img[..., -1:]  # img_z == -1
    img = img[..., :2]
    x = img[..., 0] * fx + cx
    y = img[..., 1] * -fy + cy
    return torch.stack([x, y], dim=-1).to(torch.int32)


def pixel_directions(intrin: torch.Tensor, c2w: torch.Tensor, pixels: torch.Tensor) -> torch.Tensor:
    """Unit world directions through integer pixels ``[n,2]`` (x, y) - NOT through the pixel centres: the prompt code
    uses the raw click coordinates (sam_model.py:452-460), unlike ``Cameras.generate_rays``."""
    fx, fy, cx, cy = intrin[0, 0], intrin[1, 1], intrin[0, 2], intrin[1, 2]
    x = (pixels[..., 0] - cx) / fx
    y = -(pixels[..., 1] - cy) / fy
    coords = torch.stack([x, y, -torch.ones_like(x)], dim=-1)[..., None, :]
    rotation = c2w[:3, :3].unsqueeze(0).repeat(coords.shape[0], 1, 1)
    direction = torch.sum(coords * rotation, dim=-1)
    return direction / torch.norm(direction, dim=-1, keepdim=True)


def lift_points(pixels: torch.Tensor, depth: torch.Tensor, intrin: torch.Tensor, c2w: torch.Tensor) -> torch.Tensor:
    """Clicked pixels ``[n,2]`` (x, y; long) -> 3-D prompts ``[n,3]``: the point at the rendered depth minus ``TOR``
    along the pixel's ray (sam_model.py:447-462).  ``depth``: the rendered ``[H,W,1]`` median-depth map."""
    pixels = pixels.to(torch.long)
    t = depth[pixels[..., 1], pixels[..., 0]] - TOR
    direction = pixel_directions(intrin, c2w, pixels)
    return c2w[:3, 3] + t.to(direction) * direction


def prompts_in_image(prompts_3d: torch.Tensor, intrin: torch.Tensor, c2w: torch.Tensor, width: int, height: int) -> torch.Tensor:
    """3-D prompts -> the pixel prompts that fall inside the ``width x height`` image (sam_model.py:468-475)."""
    px = project(intrin, c2w, prompts_3d)
    bounds = torch.tensor([[width, height]]).to(px)
    legal = torch.logical_and(px >= 0, px < bounds).all(dim=-1)
    return px[legal]


def visible(pixels: torch.Tensor, prompts_3d: torch.Tensor, depth: torch.Tensor, intrin: torch.Tensor, c2w: torch.Tensor,
            t_reduce: str = "min") -> torch.Tensor:
    """Which projected prompts are not occluded in this view: ray parameter of the 3-D prompt against the rendered
    depth with ``EPS`` slack (sam_model.py:60-81)."""
    pixels = pixels.to(torch.long)
    fx, fy, cx, cy = intrin[0, 0], intrin[1, 1], intrin[0, 2], intrin[1, 2]
    coords = (pixels - torch.tensor([[cx, cy]])) / torch.tensor([[fx, -fy]])
    coords = torch.cat([coords, -torch.ones_like(coords[..., :1])], dim=-1)[..., None, :]
    rays_d = torch.sum(coords * c2w[:3, :3], dim=-1)
    rays_d = rays_d / rays_d.norm(dim=-1, keepdim=True)
    rays_o = c2w[:3, 3].unsqueeze(0).repeat(coords.shape[0], 1)
    ts = (prompts_3d - rays_o) / rays_d
    ts = ts.min(dim=-1)[0] if t_reduce == "min" else ts.mean(dim=-1)
    return ts < (depth[pixels[..., 1], pixels[..., 0]].to(ts.device).squeeze() + EPS)


def predictor_input_size(original_image_size: Tuple[int, int], img_size: int = 1024) -> Tuple[int, int]:
    """``SamPredictor.set_feature``'s ``input_size`` (predictor.py:104-116)."""
    h, w = original_image_size
    if h <= w:
        return int(math.ceil(h / w * img_size)), img_size
    return img_size, int(math.ceil(w / h * img_size))


def pad_feature_map(feature: torch.Tensor) -> torch.Tensor:
    """Rendered ``sam[fh,fw,256]`` -> the ``[1,256,S,S]`` embedding SAM's mask decoder expects, zero-padded on the
    short side (predictor.py:117-125).  Deliberate deviation: for portrait images (h > w) the reference concatenates
    along the wrong axis (``dim=2`` with a ``[1,c,h,h-w]`` block) and raises; the evident intent - pad the width -
    is what this does."""
    f = feature.permute(2, 0, 1).unsqueeze(dim=0)
    c, h, w = f.shape[-3:]
    if h < w:
        f = torch.cat([f, torch.zeros(1, c, w - h, w).to(f)], dim=2)
    if h > w:
        f = torch.cat([f, torch.zeros(1, c, h, h - w).to(f)], dim=3)
    return f


def clipseg_activations(clipseg_map: torch.Tensor) -> List[torch.Tensor]:
    """Rendered ``clipseg[32,32,192]`` -> the three ``[1 + 1024, 1, 64]`` activation tensors the ClipSeg decoder takes,
    the CLS slot filled with the mean token (sam_model.py:487-492)."""
    acts = []
    for i in range(3):
        a = clipseg_map[..., 64 * i: 64 * (i + 1)].reshape(-1, 64).unsqueeze(dim=1)
        acts.append(torch.cat([a.mean(dim=0, keepdim=True), a], dim=0))
    return acts
