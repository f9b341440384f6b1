"""Hyper-parameters of the SAM-NeRF rendering hot path.

Values mirror the two shipped reference configs (``samnerf/samconfigs.py:51-161``)
on top of the nerfacto defaults (``nerfstudio/models/nerfacto.py:69-137``) and the
``SAMModelConfig`` defaults (``samnerf/sam_model.py:141-161``).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Tuple

import numpy as np


@dataclass(frozen=True)
class GridConfig:
    """One tcnn ``HashGrid`` encoding (``samnerf/sam_field.py:96-110``,
    ``nerfstudio/fields/nerfacto_field.py:157-166``, ``density_fields.py:73-81``)."""

    n_levels: int
    n_features: int
    log2_hashmap_size: int
    base_resolution: int
    max_resolution: int

    @property
    def per_level_scale(self) -> float:
        # reference: np.exp((np.log(max_res) - np.log(base_res)) / (num_levels - 1))
        return float(np.exp((np.log(self.max_resolution) - np.log(self.base_resolution)) / (self.n_levels - 1)))

    def levels(self) -> List[Tuple[float, int, int, int, bool]]:
        """Per level ``(scale, resolution, offset_entries, size_entries, hashed)``.

        Restates tiny-cuda-nn's ``grid_scale`` / ``grid_resolution`` / level-size rule
        (SURVEY.md section 8 a-17): ``scale = exp2f(l*log2f(pls))*base - 1``, ``res = ceil(scale)+1``,
        ``size = min(round_up(res^3, 8), 2^log2T)``; a level is hashed iff ``res^3 > size``.
        All arithmetic in float32 like the CUDA code it restates.
        """
        log2_pls = np.float32(np.log2(np.float32(self.per_level_scale)))
        out = []
        offset = 0
        for l in range(self.n_levels):
            scale = np.float32(np.exp2(np.float32(l) * log2_pls)) * np.float32(self.base_resolution) - np.float32(1.0)
            res = int(math.ceil(float(scale))) + 1
            dense = res**3
            size = min((dense + 7) // 8 * 8, 1 << self.log2_hashmap_size)
            out.append((float(scale), res, offset, size, dense > size))
            offset += size
        return out

    @property
    def n_entries(self) -> int:
        lv = self.levels()
        return lv[-1][2] + lv[-1][3]

    @property
    def n_params(self) -> int:
        return self.n_entries * self.n_features

    @property
    def n_output_dims(self) -> int:
        return self.n_levels * self.n_features


def _pad16(n: int) -> int:
    return (n + 15) // 16 * 16


@dataclass(frozen=True)
class SAMNeRFConfig:
    """Everything the renderer needs to know about the model (defaults = ``samnerf_distill``)."""

    # samplers (samconfigs.py:84-87,138-141)
    num_proposal_samples: int = 64
    num_nerf_samples: int = 32
    num_sam_samples: int = 16
    sharpening_temperature: float = 10.0  # sam_model.py:161
    patch_size: int = 4  # samconfigs.py:136
    kernel_size: int = 3
    distill_sam: bool = True
    use_clipseg_feature: bool = True
    near_plane_eval: float = 0.0  # scene_colliders.py:185 (eval mode)
    far_plane: float = 1000.0  # nerfacto.py:73
    histogram_padding: float = 0.01  # ray_samplers.py:268
    eval_num_rays_per_chunk: int = 1 << 15  # samconfigs.py:79,133
    # training-side knobs (nerfacto.py:110-125, sam_model.py:143-147)
    interlevel_loss_mult: float = 1.0
    distortion_loss_mult: float = 0.002
    sam_loss_weight: float = 1.0
    clipseg_loss_weight: float = 1.0
    use_proposal_weight_anneal: bool = True
    proposal_weights_anneal_slope: float = 10.0
    proposal_weights_anneal_max_num_iters: int = 1000
    proposal_warmup: int = 5000
    proposal_update_every: int = 5
    # proposal density field (nerfacto.py:105, density_fields.py:50-100)
    proposal_grid: GridConfig = field(default_factory=lambda: GridConfig(5, 2, 17, 16, 128))
    proposal_hidden: int = 16
    # main nerfacto field (nerfacto_field.py:93-175,228-240)
    field_grid: GridConfig = field(default_factory=lambda: GridConfig(16, 2, 19, 16, 2048))
    field_hidden: int = 64
    geo_feat_dim: int = 15
    head_hidden: int = 64
    sh_degree: int = 4
    # SAM / ClipSeg feature field (sam_model.py:153-155, sam_field.py:26-94)
    sam_grids: Tuple[GridConfig, ...] = field(
        default_factory=lambda: (GridConfig(12, 8, 19, 16, 128), GridConfig(12, 8, 19, 128, 512))
    )
    sam_hidden: int = 256
    sam_out: int = 256
    clipseg_out: int = 192

    # ---- derived sizes (tcnn layout: network weights then grid) -------------
    @property
    def proposal_in(self) -> int:
        return _pad16(self.proposal_grid.n_output_dims)

    @property
    def proposal_mlp_params(self) -> int:
        return self.proposal_hidden * self.proposal_in + 16 * self.proposal_hidden

    @property
    def field_mlp_params(self) -> int:
        return self.field_hidden * self.field_grid.n_output_dims + _pad16(1 + self.geo_feat_dim) * self.field_hidden

    @property
    def head_in(self) -> int:
        return _pad16(self.sh_degree**2 + self.geo_feat_dim)

    @property
    def head_mlp_params(self) -> int:
        return self.head_hidden * self.head_in + self.head_hidden * self.head_hidden + 16 * self.head_hidden

    @property
    def sam_in(self) -> int:
        return sum(g.n_output_dims for g in self.sam_grids)

    @property
    def sam_mlp_params(self) -> int:
        return self.sam_hidden * self.sam_in + self.sam_out * self.sam_hidden

    @property
    def clipseg_mlp_params(self) -> int:
        return self.sam_hidden * self.sam_in + self.clipseg_out * self.sam_hidden

    @staticmethod
    def no_distill() -> "SAMNeRFConfig":
        """``samnerf_no_distill`` (samconfigs.py:51-102): RGB only, patch 1."""
        return SAMNeRFConfig(distill_sam=False, use_clipseg_feature=False, patch_size=1, num_sam_samples=3)

    @staticmethod
    def distill(clipseg: bool = True, patch_size: int = 4) -> "SAMNeRFConfig":
        return SAMNeRFConfig(distill_sam=True, use_clipseg_feature=clipseg, patch_size=patch_size)

    @staticmethod
    def tiny(clipseg: bool = True, patch_size: int = 4) -> "SAMNeRFConfig":
        """Same topology with small tables: used by fixtures that carry their own parameters."""
        return SAMNeRFConfig(
            patch_size=patch_size,
            use_clipseg_feature=clipseg,
            proposal_grid=GridConfig(5, 2, 11, 16, 128),
            field_grid=GridConfig(16, 2, 12, 16, 2048),
            sam_grids=(GridConfig(12, 8, 11, 16, 128), GridConfig(12, 8, 11, 128, 512)),
        )


def get_feature_size(h: int, w: int, largesize: int = 64) -> Tuple[int, int]:
    """``samnerf/sam_utils.py:7-14`` with the square case defined.

    The reference has no branch for ``h == w`` and raises ``UnboundLocalError``; this build
    defines it as ``(largesize, largesize)`` (deliberate, documented deviation - SURVEY section 0 item 9).
    """
    if h < w:
        return int(math.ceil((h / w) * largesize)), largesize
    if h > w:
        return largesize, int(math.ceil((w / h) * largesize))
    return largesize, largesize
