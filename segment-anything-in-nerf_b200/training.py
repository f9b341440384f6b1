"""Loss terms the training loop builds on the path's outputs ("weights_list / ray_samples_list
outputs for interlevel_loss / distortion_loss").  Plain torch on ``[N,S]`` tensors, exactly where the reference computes
them (torch autograd handles these; the kernels' backward entry points take over below ``weights`` / ``rgb``).

Restated from (paths relative to /root/reference):
  * proposal ("interlevel") loss of mip-NeRF 360   nerfstudio/model_components/losses.py:33-120
  * distortion loss                                nerfstudio/model_components/losses.py:124-143
  * loss / metric dictionaries                     nerfstudio/models/nerfacto.py:316-344, samnerf/sam_model.py:316-328
  * proposal-weight annealing schedule             nerfstudio/models/nerfacto.py:242-256
"""
from __future__ import annotations

from enum import Enum, auto
from inspect import signature
from typing import Callable, Dict, List, Optional, Tuple

import numpy as np
import torch

EPS = 1.0e-7  # losses.py:22


def s_edges(ray_samples) -> torch.Tensor:
    """``ray_samples_to_sdist`` (losses.py:96-101): the S+1 bin edges of a sample list in spacing space."""
    return torch.cat([ray_samples.spacing_starts[..., 0], ray_samples.spacing_ends[..., -1:, 0]], dim=-1)


def envelope_mass(t: torch.Tensor, t_env: torch.Tensor, w_env: torch.Tensor) -> torch.Tensor:
    """Upper bound on the mass the enveloping histogram ``(t_env, w_env)`` assigns to every bin of ``t``
    (``outer``, losses.py:33-64): cumulative envelope weight between the envelope bins that contain the bin's ends."""
    lo_edge, hi_edge = t[..., :-1].contiguous(), t[..., 1:].contiguous()
    cum = torch.cat([torch.zeros_like(w_env[..., :1]), torch.cumsum(w_env, dim=-1)], dim=-1)
    last = w_env.shape[-1] - 1
    first_bin = (torch.searchsorted(t_env[..., :-1].contiguous(), lo_edge, side="right") - 1).clamp(0, last)
    last_bin = torch.searchsorted(t_env[..., 1:].contiguous(), hi_edge, side="right").clamp(0, last)
    return torch.take_along_dim(cum[..., 1:], last_bin, dim=-1) - torch.take_along_dim(cum[..., :-1], first_bin, dim=-1)


def interlevel_loss(weights_list: List[torch.Tensor], ray_samples_list: List) -> torch.Tensor:
    """losses.py:104-118: the final level's histogram (detached) must stay under every proposal level's envelope."""
    t = s_edges(ray_samples_list[-1]).detach()
    w = weights_list[-1][..., 0].detach()
    total = 0.0
    for samples, weights in zip(ray_samples_list[:-1], weights_list[:-1]):
        excess = torch.clip(w - envelope_mass(t, s_edges(samples), weights[..., 0]), min=0)
        total = total + torch.mean(excess**2 / (w + EPS))
    return total


def distortion_loss(weights_list: List[torch.Tensor], ray_samples_list: List) -> torch.Tensor:
    """losses.py:122-143 on the final level: pairwise midpoint distances weighted by both weights, plus the
    within-bin term."""
    t = s_edges(ray_samples_list[-1])
    w = weights_list[-1][..., 0]
    mid = (t[..., 1:] + t[..., :-1]) / 2
    between = torch.sum(w * torch.sum(w[..., None, :] * torch.abs(mid[..., :, None] - mid[..., None, :]), dim=-1), dim=-1)
    within = torch.sum(w**2 * (t[..., 1:] - t[..., :-1]), dim=-1) / 3
    return torch.mean(between + within)


def psnr(pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """torchmetrics' PeakSignalNoiseRatio(data_range=1.0) as nerfacto.py:231,319 uses it."""
    return -10.0 * torch.log10(torch.mean((pred - target) ** 2))


def metrics_dict(outputs: Dict, batch: Dict, training: bool) -> Dict[str, torch.Tensor]:
    """nerfacto.py:316-322."""
    image = batch["image"].to(outputs["rgb"].device)
    m = {"psnr": psnr(outputs["rgb"].detach(), image)}
    if training:
        m["distortion"] = distortion_loss(outputs["weights_list"], outputs["ray_samples_list"])
    return m


def loss_dict(outputs: Dict, batch: Dict, metrics: Dict, cfg, training: bool) -> Dict[str, torch.Tensor]:
    """nerfacto.py:324-344 + sam_model.py:316-328 (feature losses: per-ray mean over channels, ``nanmean`` over rays,
    because rays whose top-k weights are all zero render NaN features, sam_model.py:248)."""
    dev = outputs["rgb"].device
    losses = {"rgb_loss": torch.nn.functional.mse_loss(batch["image"].to(dev), outputs["rgb"])}
    if training:
        losses["interlevel_loss"] = cfg.interlevel_loss_mult * interlevel_loss(outputs["weights_list"], outputs["ray_samples_list"])
        losses["distortion_loss"] = cfg.distortion_loss_mult * metrics["distortion"]
        for name, weight in (("sam", cfg.sam_loss_weight), ("clipseg", cfg.clipseg_loss_weight)):
            if name in outputs and name in batch:
                per_elem = torch.nn.functional.mse_loss(outputs[name], batch[name].to(dev), reduction="none")
                losses[f"{name}_loss"] = weight * per_elem.mean(dim=-1).nanmean()
    return losses


def proposal_anneal(step: int, max_num_iters: int = 1000, slope: float = 10.0) -> float:
    """nerfacto.py:248-255 (mip-NeRF 360 eq. 18): the exponent applied to the proposal weights before PDF sampling."""
    frac = float(np.clip(step / max_num_iters, 0, 1))
    return (slope * frac) / ((slope - 1) * frac + 1)


def proposal_update_schedule(warmup: int = 5000, update_every: int = 5):
    """nerfacto.py:196-200: steps between proposal-network updates, ramping up over the warm-up."""
    return lambda step: float(np.clip(np.interp(step, [0, warmup], [0, update_every]), 1, update_every))


class TrainingCallbackLocation(Enum):
    """nerfstudio/engine/callbacks.py:42-46."""

    BEFORE_TRAIN_ITERATION = auto()
    AFTER_TRAIN_ITERATION = auto()


class TrainingCallback:
    """nerfstudio/engine/callbacks.py:49-110: ``func(step=...)`` every ``update_every_num_iters`` iterations (or at the
    listed ``iters``), at the places named in ``where_to_run``.  What ``Model.get_training_callbacks`` hands the trainer."""

    def __init__(self, where_to_run: List[TrainingCallbackLocation], func: Callable, update_every_num_iters: Optional[int] = None,
                 iters: Optional[Tuple[int, ...]] = None, args: Optional[List] = None, kwargs: Optional[Dict] = None):
        assert "step" in signature(func).parameters, f"'step: int' must be an argument of the callback {func}"
        self.where_to_run, self.func = where_to_run, func
        self.update_every_num_iters, self.iters = update_every_num_iters, iters
        self.args, self.kwargs = args or [], kwargs or {}

    def run_callback(self, step: int) -> None:
        if self.update_every_num_iters is not None:
            if step % self.update_every_num_iters == 0:
                self.func(*self.args, **self.kwargs, step=step)
        elif self.iters is not None:
            if step in self.iters:
                self.func(*self.args, **self.kwargs, step=step)

    def run_callback_at_location(self, step: int, location: TrainingCallbackLocation) -> None:
        if location in self.where_to_run:
            self.run_callback(step=step)
