"""Synthetic (random-init) parameters and rays for benchmarks and tests.

There is no network for checkpoints or datasets, so every measured configuration uses
random parameters of the reference architecture, generated on the CPU with a fixed seed
and shared byte-for-byte between the CUDA path and the CPU oracle (SURVEY.md section 8 d).

Keys are the reference's ``state_dict`` names; every tcnn module is one flat fp32
``params`` tensor in tcnn order (network weights first, then the grid).
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import torch

from .config import SAMNeRFConfig

REGIMES = ("init", "scene")


def _xavier(gen: torch.Generator, out_f: int, in_f: int, gain: float) -> torch.Tensor:
    a = gain * math.sqrt(6.0 / (in_f + out_f))
    return (torch.rand(out_f * in_f, generator=gen, dtype=torch.float32) * 2 - 1) * a


def _grid(gen: torch.Generator, n: int, amp: float) -> torch.Tensor:
    return (torch.rand(n, generator=gen, dtype=torch.float32) * 2 - 1) * amp


def make_synthetic_params(cfg: SAMNeRFConfig, regime: str = "scene", seed: int = 0) -> Dict[str, torch.Tensor]:
    """Random parameters in the reference's checkpoint layout.

    ``init``  - tcnn defaults: grids U(-1e-4, 1e-4), Xavier-uniform MLPs (sigma ~ 1 everywhere).
    ``scene`` - grids U(-0.5, 0.5), Xavier x 2 feature/colour MLPs, Xavier x 6 density MLPs whose density
                row is shifted negative (proposal -0.5, field -0.1): ``exp(MLP)`` then spans ~7 orders of
                magnitude, space is mostly empty, rays terminate at ray-dependent depths from 0.1 to beyond the
                contraction radius, accumulation stays ~1 and the sharpened top-k weights concentrate on 1-3
                samples - the statistics of a trained scene (checked in tests/test_host.py::test_scene_regime_is_scene_like).
    """
    assert regime in REGIMES, regime
    gen = torch.Generator(device="cpu")
    gen.manual_seed(seed)
    amp = 1e-4 if regime == "init" else 0.5
    gain = 1.0 if regime == "init" else 2.0
    dgain = 1.0 if regime == "init" else 6.0
    shift_p, shift_f = (0.0, 0.0) if regime == "init" else (0.5, 0.1)
    p: Dict[str, torch.Tensor] = {}

    def density_out(hidden: int, shift: float) -> torch.Tensor:
        w = _xavier(gen, 16, hidden, dgain).view(16, hidden)
        w[0, :] -= shift  # row 0 is the density pre-activation
        return w.reshape(-1)

    pg = cfg.proposal_grid
    p["proposal_networks.0.mlp_base.params"] = torch.cat(
        [
            _xavier(gen, cfg.proposal_hidden, cfg.proposal_in, dgain),
            density_out(cfg.proposal_hidden, shift_p),
            _grid(gen, pg.n_params, amp),
        ]
    )
    fg = cfg.field_grid
    p["field.mlp_base.params"] = torch.cat(
        [
            _xavier(gen, cfg.field_hidden, fg.n_output_dims, dgain),
            density_out(cfg.field_hidden, shift_f),
            _grid(gen, fg.n_params, amp),
        ]
    )
    p["field.mlp_head.params"] = torch.cat(
        [
            _xavier(gen, cfg.head_hidden, cfg.head_in, gain),
            _xavier(gen, cfg.head_hidden, cfg.head_hidden, gain),
            _xavier(gen, 16, cfg.head_hidden, gain),
        ]
    )
    if cfg.distill_sam:
        for i, g in enumerate(cfg.sam_grids):
            p[f"sam_field.clip_encs.{i}.params"] = _grid(gen, g.n_params, amp)
        p["sam_field.sam_net.params"] = torch.cat(
            [_xavier(gen, cfg.sam_hidden, cfg.sam_in, gain), _xavier(gen, cfg.sam_out, cfg.sam_hidden, gain)]
        )
        if cfg.use_clipseg_feature:
            for i, g in enumerate(cfg.sam_grids):
                p[f"sam_field.clipseg_encs.{i}.params"] = _grid(gen, g.n_params, amp)
            p["sam_field.clipseg_net.params"] = torch.cat(
                [_xavier(gen, cfg.sam_hidden, cfg.sam_in, gain), _xavier(gen, cfg.clipseg_out, cfg.sam_hidden, gain)]
            )
        # conv head: PyTorch Conv2d default init (kaiming_uniform a=sqrt(5) -> U(-1/sqrt(fan_in), 1/sqrt(fan_in)))
        k = cfg.kernel_size
        bound = 1.0 / math.sqrt(cfg.sam_out * k * k)
        for idx in (0, 2):
            p[f"conv_head.{idx}.weight"] = _grid(gen, cfg.sam_out * cfg.sam_out * k * k, bound).view(
                cfg.sam_out, cfg.sam_out, k, k
            )
            p[f"conv_head.{idx}.bias"] = _grid(gen, cfg.sam_out, bound)
    return p


def pinhole_rays(
    height: int, width: int, fx: float, fy: float, c2w: torch.Tensor
) -> Tuple[torch.Tensor, torch.Tensor]:
    """Pinhole rays the way ``nerfstudio/cameras/cameras.py:607-714`` builds them: pixel centre
    +0.5, -z forward, y flipped, directions normalised.  Returns ``origins[H,W,3], dirs[H,W,3]`` fp32."""
    cx, cy = width / 2.0, height / 2.0
    ys, xs = torch.meshgrid(
        torch.arange(height, dtype=torch.float32) + 0.5, torch.arange(width, dtype=torch.float32) + 0.5, indexing="ij"
    )
    d = torch.stack([(xs - cx) / fx, -(ys - cy) / fy, -torch.ones_like(xs)], dim=-1)
    d = d @ c2w[:3, :3].T.to(torch.float32)
    d = d / torch.linalg.norm(d, dim=-1, keepdim=True)
    o = c2w[:3, 3].to(torch.float32).expand_as(d).contiguous()
    return o, d.contiguous()


def look_at(eye: Tuple[float, float, float], target: Tuple[float, float, float] = (0.0, 0.0, 0.0)) -> torch.Tensor:
    """Camera-to-world with -z forward, +y up (z-up world)."""
    e = torch.tensor(eye, dtype=torch.float64)
    t = torch.tensor(target, dtype=torch.float64)
    fwd = t - e
    fwd = fwd / fwd.norm()
    up = torch.tensor([0.0, 0.0, 1.0], dtype=torch.float64)
    right = torch.linalg.cross(fwd, up)
    right = right / right.norm()
    true_up = torch.linalg.cross(right, fwd)
    c2w = torch.eye(4, dtype=torch.float64)
    c2w[:3, 0], c2w[:3, 1], c2w[:3, 2], c2w[:3, 3] = right, true_up, -fwd, e
    return c2w.to(torch.float32)


def orbit_rays(height: int = 800, width: int = 800, focal: float = 800.0, radius: float = 1.2, z: float = 0.4,
               angle: float = 0.0) -> Tuple[torch.Tensor, torch.Tensor]:
    """SURVEY 8 d config 2/3 camera: radius-1.2 orbit at height 0.4 looking at the origin."""
    eye = (radius * math.cos(angle), radius * math.sin(angle), z)
    return pinhole_rays(height, width, focal, focal, look_at(eye))


def plumbing_rays() -> Tuple[torch.Tensor, torch.Tensor]:
    """SURVEY 8 d config 1: two 64x64 cameras at (+-1.5, 0, 0.5), fx = fy = 64 -> 8192 rays."""
    os_, ds_ = [], []
    for x in (1.5, -1.5):
        o, d = pinhole_rays(64, 64, 64.0, 64.0, look_at((x, 0.0, 0.5)))
        os_.append(o.reshape(-1, 3))
        ds_.append(d.reshape(-1, 3))
    return torch.cat(os_), torch.cat(ds_)
