"""CPU oracle for the SAM-NeRF rendering hot path (plain PyTorch fp32 on the CPU).

TEST INFRASTRUCTURE ONLY - this file is the checker, never the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs may import it.
It does not import the product package and the product never imports it.

Pinning: ``oracle/make_golden.py`` runs the reference's OWN Python for this path
(``ProposalNetworkSampler``, ``PDFSampler``, ``UniformLinDispPiecewiseSampler``, ``RaySamples.get_weights``,
the three renderers, ``SceneContraction``, ``NearFarCollider``, ``HashMLPDensityField``,
``TCNNNerfactoField``, ``SAMField`` and the bodies of ``SAMModel.get_outputs`` / ``_get_outputs_nerfacto`` /
``get_outputs_for_camera_ray_bundle``) unmodified from /root/reference with ``oracle/fake_tinycudann.py``
injected, stores its outputs under ``tests/golden/`` and ``tests/test_oracle_golden.py`` holds this restatement
to them.  The tiny-cuda-nn boundary itself is PARITY UNPINNED (see ``oracle/tcnn_spec.py``).

Every function cites the reference lines it follows (paths relative to /root/reference).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

from . import tcnn_spec as T


# --------------------------------------------------------------------------------------------
# small pieces
# --------------------------------------------------------------------------------------------
def contract(x: torch.Tensor, order: Optional[float]) -> torch.Tensor:
    """``SceneContraction.forward`` - nerfstudio/field_components/spatial_distortions.py:66-88."""
    mag = torch.linalg.norm(x, ord=order, dim=-1)[..., None]
    return torch.where(mag < 1, x, (2 - (1 / mag)) * (x / mag))


class _TruncExp(torch.autograd.Function):
    """``trunc_exp`` - nerfstudio/field_components/activations.py:24-40: forward ``exp(x)``, backward
    ``g * exp(clamp(x, -15, 15))``."""

    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return torch.exp(x)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        return g * torch.exp(x.clamp(-15, 15))


def trunc_exp(x: torch.Tensor) -> torch.Tensor:
    return _TruncExp.apply(x) if x.requires_grad else torch.exp(x)


def spacing_fn(x: torch.Tensor) -> torch.Tensor:
    """UniformLinDispPiecewiseSampler spacing - nerfstudio/model_components/ray_samplers.py:242."""
    return torch.where(x < 1, x / 2, 1 - 1 / (2 * x))


def spacing_fn_inv(x: torch.Tensor) -> torch.Tensor:
    """ray_samplers.py:243."""
    return torch.where(x < 0.5, 2 * x, 1 / (2 - 2 * x))


def get_weights(deltas: torch.Tensor, densities: torch.Tensor) -> torch.Tensor:
    """``RaySamples.get_weights`` - nerfstudio/cameras/rays.py:141-163.  ``[N,S]`` in, ``[N,S]`` out."""
    delta_density = deltas * densities
    alphas = 1 - torch.exp(-delta_density)
    transmittance = torch.cumsum(delta_density[..., :-1], dim=-1)
    transmittance = torch.cat([torch.zeros_like(transmittance[..., :1]), transmittance], dim=-1)
    transmittance = torch.exp(-transmittance)
    return torch.nan_to_num(alphas * transmittance)


def median_depth(weights: torch.Tensor, starts: torch.Tensor, ends: torch.Tensor) -> torch.Tensor:
    """``DepthRenderer`` (median) - nerfstudio/model_components/renderers.py:260-270.  Returns ``[N,1]``."""
    steps = (starts + ends) / 2
    cumulative = torch.cumsum(weights, dim=-1)
    split = torch.ones((*weights.shape[:-1], 1)) * 0.5
    idx = torch.searchsorted(cumulative, split, side="left")
    idx = torch.clamp(idx, 0, steps.shape[-1] - 1)
    return torch.gather(steps, dim=-1, index=idx)


def composite_rgb(rgb: torch.Tensor, weights: torch.Tensor, background=None, training: bool = False) -> torch.Tensor:
    """``RGBRenderer.forward`` with ``background_color="last_sample"`` (or a fixed RGB override)
    - renderers.py:69-140; nerfacto.py:77,220.  ``rgb[N,S,3]``, ``weights[N,S]`` -> ``[N,3]``.
    Eval mode adds ``nan_to_num`` on the samples and the final clamp (renderers.py:132-139); training has neither."""
    if not training:
        rgb = torch.nan_to_num(rgb)
    comp = torch.sum(weights[..., None] * rgb, dim=-2)
    acc = torch.sum(weights[..., None], dim=-2)
    bg = rgb[..., -1, :] if background is None else torch.as_tensor(background, dtype=torch.float32)
    comp = comp + bg * (1.0 - acc)
    return comp if training else torch.clamp(comp, 0.0, 1.0)


def get_feature_size(h: int, w: int, largesize: int = 64) -> Tuple[int, int]:
    """samnerf/sam_utils.py:7-14; the square case (undefined in the reference -> UnboundLocalError)
    is defined here as (largesize, largesize) - the documented deviation."""
    if h < w:
        return int(math.ceil((h / w) * largesize)), largesize
    if h > w:
        return largesize, int(math.ceil((w / h) * largesize))
    return largesize, largesize


# --------------------------------------------------------------------------------------------
# the model
# --------------------------------------------------------------------------------------------
class Oracle:
    """Holds parameters (reference ``state_dict`` names, flat tcnn layout) and evaluates the path."""

    def __init__(self, cfg, params: Dict[str, torch.Tensor]):
        self.cfg = cfg
        self.p = params
        g = cfg.proposal_grid
        self.prop_levels = T.grid_levels(g.n_levels, g.base_resolution, g.per_level_scale, g.log2_hashmap_size)
        flat = params["proposal_networks.0.mlp_base.params"]
        n_net = cfg.proposal_hidden * cfg.proposal_in + 16 * cfg.proposal_hidden
        self.prop_w = T.split_mlp_params(flat[:n_net], [cfg.proposal_in, cfg.proposal_hidden, 16])
        self.prop_table = flat[n_net:]

        g = cfg.field_grid
        self.field_levels = T.grid_levels(g.n_levels, g.base_resolution, g.per_level_scale, g.log2_hashmap_size)
        flat = params["field.mlp_base.params"]
        n_net = cfg.field_hidden * g.n_output_dims + 16 * cfg.field_hidden
        self.base_w = T.split_mlp_params(flat[:n_net], [g.n_output_dims, cfg.field_hidden, 16])
        self.field_table = flat[n_net:]
        self.head_w = T.split_mlp_params(
            params["field.mlp_head.params"], [cfg.head_in, cfg.head_hidden, cfg.head_hidden, 16]
        )
        if cfg.distill_sam:
            self.sam_levels = [
                T.grid_levels(g.n_levels, g.base_resolution, g.per_level_scale, g.log2_hashmap_size)
                for g in cfg.sam_grids
            ]
            self.sam_w = T.split_mlp_params(
                params["sam_field.sam_net.params"], [cfg.sam_in, cfg.sam_hidden, cfg.sam_out]
            )
            if cfg.use_clipseg_feature:
                self.clipseg_w = T.split_mlp_params(
                    params["sam_field.clipseg_net.params"], [cfg.sam_in, cfg.sam_hidden, cfg.clipseg_out]
                )

    # ---- fields ---------------------------------------------------------------------------
    def _normalized(self, positions: torch.Tensor, order) -> Tuple[torch.Tensor, torch.Tensor]:
        """density_fields.py:102-112 / nerfacto_field.py:244-253: contraction, (p+2)/4, (0,1) selector."""
        x = (contract(positions, order) + 2.0) / 4.0
        sel = ((x > 0.0) & (x < 1.0)).all(dim=-1)
        return x * sel[..., None], sel

    def proposal_density(self, positions: torch.Tensor) -> torch.Tensor:
        """``HashMLPDensityField.get_density`` - nerfstudio/fields/density_fields.py:102-125 (via
        ``Field.density_fn`` base_field.py:38-56).  ``positions[...,3]`` -> ``[...]``."""
        shp = positions.shape[:-1]
        x, sel = self._normalized(positions.reshape(-1, 3), float("inf"))
        enc = T.hash_grid_encode(x, self.prop_table, self.prop_levels, self.cfg.proposal_grid.n_features)
        pad = self.cfg.proposal_in - enc.shape[-1]
        if pad:
            enc = torch.cat([enc, torch.zeros(enc.shape[0], pad)], dim=-1)  # grid encodings pad with 0
        h = T.mlp_forward(enc, self.prop_w)[:, 0]
        return (trunc_exp(h) * sel).view(shp)

    def field_density(self, positions: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """``TCNNNerfactoField.get_density`` - nerfstudio/fields/nerfacto_field.py:242-266.
        Returns ``density[...]`` fp32 and ``geo[...,15]`` (fp16-rounded)."""
        shp = positions.shape[:-1]
        x, sel = self._normalized(positions.reshape(-1, 3), float("inf"))
        enc = T.hash_grid_encode(x, self.field_table, self.field_levels, self.cfg.field_grid.n_features)
        h = T.mlp_forward(enc, self.base_w)
        density = trunc_exp(h[:, 0]) * sel
        return density.view(shp), h[:, 1 : 1 + self.cfg.geo_feat_dim].reshape(*shp, -1)

    def field_rgb(self, directions: torch.Tensor, geo: torch.Tensor) -> torch.Tensor:
        """``TCNNNerfactoField.get_outputs`` with appearance embedding off - nerfacto_field.py:268-351
        (``get_normalized_directions`` :58-64).  ``directions[...,3]`` broadcast per sample, ``geo[...,15]``."""
        shp = geo.shape[:-1]
        d = T.sh4(((directions + 1.0) / 2.0).reshape(-1, 3))
        h = torch.cat([d, geo.reshape(-1, geo.shape[-1])], dim=-1)
        pad = self.cfg.head_in - h.shape[-1]
        if pad:
            h = torch.cat([h, torch.ones(h.shape[0], pad)], dim=-1)  # tcnn.Network pads identity inputs with 1
        rgb = T.mlp_forward(h, self.head_w, output_activation="Sigmoid")[:, :3]
        return rgb.reshape(*shp, 3)

    def sam_field(self, positions: torch.Tensor, which: Sequence[str] = ("sam",)) -> Dict[str, torch.Tensor]:
        """``SAMField.get_outputs`` - samnerf/sam_field.py:112-140 (L2 contraction :32, no selector)."""
        shp = positions.shape[:-1]
        x = (contract(positions.reshape(-1, 3), None) + 2.0) / 4.0
        out: Dict[str, torch.Tensor] = {}
        if "sam" in which:
            enc = torch.cat(
                [
                    T.hash_grid_encode(x, self.p[f"sam_field.clip_encs.{i}.params"], lv, g.n_features)
                    for i, (lv, g) in enumerate(zip(self.sam_levels, self.cfg.sam_grids))
                ],
                dim=-1,
            )
            out["hashgrid"] = enc.view(*shp, -1)
            out["sam"] = T.mlp_forward(enc, self.sam_w).view(*shp, -1)
        if "clipseg" in which and self.cfg.use_clipseg_feature:
            enc = torch.cat(
                [
                    T.hash_grid_encode(x, self.p[f"sam_field.clipseg_encs.{i}.params"], lv, g.n_features)
                    for i, (lv, g) in enumerate(zip(self.sam_levels, self.cfg.sam_grids))
                ],
                dim=-1,
            )
            out["clipseg"] = T.mlp_forward(enc, self.clipseg_w).view(*shp, -1)
        return out

    # ---- samplers -------------------------------------------------------------------------
    def initial_samples(self, nears: torch.Tensor, fars: torch.Tensor, t_rand: Optional[torch.Tensor] = None):
        """``SpacedSampler.generate_ray_samples`` with the piecewise spacing - ray_samplers.py:79-126,223-246.
        ``t_rand[N,1]``: the single-jitter random numbers of training mode (:104-112; ``single_jitter=True`` is the
        nerfacto default, nerfacto.py:113,211); ``None`` = eval (no jitter).  Returns spacing bins ``[1|N,S+1]``,
        euclidean bins ``[N,S+1]`` and ``(s_near, s_far)``."""
        s = self.cfg.num_proposal_samples
        bins = torch.linspace(0.0, 1.0, s + 1)[None, ...]
        if t_rand is not None:
            bin_centers = (bins[..., 1:] + bins[..., :-1]) / 2.0
            bin_upper = torch.cat([bin_centers, bins[..., -1:]], -1)
            bin_lower = torch.cat([bins[..., :1], bin_centers], -1)
            bins = bin_lower + (bin_upper - bin_lower) * t_rand
        s_near, s_far = spacing_fn(nears), spacing_fn(fars)
        eu = spacing_fn_inv(bins * s_far + (1 - bins) * s_near)
        return bins, eu, (s_near, s_far)

    def pdf_sample(self, weights: torch.Tensor, existing_bins: torch.Tensor, s_near, s_far,
                   rand: Optional[torch.Tensor] = None):
        """``PDFSampler.generate_ray_samples`` (``include_original=False``, padding 0.01) -
        ray_samplers.py:274-369.  ``weights[N,S]`` -> spacing bins and euclidean bins ``[N, S'+1]``.
        ``rand[N,1]``: the single-jitter random numbers of training mode (:314-322); ``None`` = eval."""
        num_samples = self.cfg.num_nerf_samples
        num_bins = num_samples + 1
        eps = 1e-5
        weights = weights + self.cfg.histogram_padding
        weights_sum = torch.sum(weights, dim=-1, keepdim=True)
        padding = torch.relu(eps - weights_sum)
        weights = weights + padding / weights.shape[-1]
        weights_sum = weights_sum + padding
        pdf = weights / weights_sum
        cdf = torch.min(torch.ones_like(pdf), torch.cumsum(pdf, dim=-1))
        cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], dim=-1)
        u = torch.linspace(0.0, 1.0 - (1.0 / num_bins), steps=num_bins)
        if rand is not None:
            u = u.expand(size=(*cdf.shape[:-1], num_bins)) + rand / num_bins
        else:
            u = u + 1.0 / (2 * num_bins)
        u = u.expand(size=(*cdf.shape[:-1], num_bins)).contiguous()
        existing_bins = existing_bins.expand(cdf.shape[0], -1)
        inds = torch.searchsorted(cdf, u, side="right")
        below = torch.clamp(inds - 1, 0, existing_bins.shape[-1] - 1)
        above = torch.clamp(inds, 0, existing_bins.shape[-1] - 1)
        cdf_g0 = torch.gather(cdf, -1, below)
        bins_g0 = torch.gather(existing_bins, -1, below)
        cdf_g1 = torch.gather(cdf, -1, above)
        bins_g1 = torch.gather(existing_bins, -1, above)
        t = torch.clip(torch.nan_to_num((u - cdf_g0) / (cdf_g1 - cdf_g0), 0), 0, 1)
        bins = bins_g0 + t * (bins_g1 - bins_g0)
        eu = spacing_fn_inv(bins * s_far + (1 - bins) * s_near)
        return bins, eu

    # ---- one chunk ------------------------------------------------------------------------
    def render_rays(
        self,
        origins: torch.Tensor,
        directions: torch.Tensor,
        nears: Optional[torch.Tensor] = None,
        fars: Optional[torch.Tensor] = None,
        get_feature: Sequence[str] = ("sam",),
        fast: bool = False,
        background=None,
        return_intermediates: bool = False,
        jitter: Optional[torch.Tensor] = None,
        anneal: float = 1.0,
    ) -> Dict[str, torch.Tensor]:
        """``SAMModel.forward`` -> ``get_outputs`` in eval mode - samnerf/sam_model.py:226-314
        (collider scene_colliders.py:183-188; sampler driver ray_samplers.py:558-599)."""
        cfg = self.cfg
        n = origins.shape[0]
        if nears is None:
            nears = torch.full((n, 1), float(cfg.near_plane_eval))
        if fars is None:
            fars = torch.full((n, 1), float(cfg.far_plane))
        o, d = origins[:, None, :], directions[:, None, :]

        # proposal level (ray_samplers.py:575-593; rays.py:48-57,226-270)
        # jitter[N,2]: training-mode single-jitter random numbers of the two sampling levels (None = eval)
        bins0, eu0, (s_near, s_far) = self.initial_samples(nears, fars, None if jitter is None else jitter[:, 0:1])
        starts0, ends0 = eu0[:, :-1], eu0[:, 1:]
        pos0 = o + d * ((starts0 + ends0) / 2)[..., None]
        dens0 = self.proposal_density(pos0)
        w0 = get_weights(ends0 - starts0, dens0)
        # anneal == 1.0 in eval (ray_samplers.py:545,583)
        # proposal-weight annealing (ray_samplers.py:583; 1.0 in eval and after the first 1000 training steps)
        w_pdf = w0 if anneal == 1.0 else torch.pow(w0, anneal)
        bins1, eu1 = self.pdf_sample(w_pdf, bins0, s_near, s_far, None if jitter is None else jitter[:, 1:2])
        starts, ends = eu1[:, :-1], eu1[:, 1:]
        pos = o + d * ((starts + ends) / 2)[..., None]

        # nerfacto field + renderers (sam_model.py:280-301)
        density, geo = self.field_density(pos)
        rgb_s = self.field_rgb(d.expand(-1, pos.shape[1], -1), geo)
        weights = get_weights(ends - starts, density)
        out: Dict[str, torch.Tensor] = {
            "rgb": composite_rgb(rgb_s, weights, background),
            "depth": median_depth(weights, starts, ends),
        }
        if not fast:
            out["accumulation"] = torch.sum(weights, dim=-1, keepdim=True)
            out["prop_depth_0"] = median_depth(w0, starts0, ends0)

        # feature branch (sam_model.py:243-277)
        if cfg.distill_sam and len(get_feature) > 0:
            k = cfg.num_sam_samples
            sam_w, best = torch.topk(weights, k, dim=-1, sorted=False)
            sam_w = sam_w**cfg.sharpening_temperature
            sam_w = sam_w / sam_w.sum(dim=-1, keepdim=True)
            sam_pos = torch.gather(pos, 1, best[..., None].expand(-1, -1, 3))
            fo = self.sam_field(sam_pos, which=get_feature)
            if "sam" in get_feature:
                feat = torch.sum(sam_w[..., None] * fo["sam"], dim=-2)  # MeanRenderer sam_model.py:126-137
                if cfg.patch_size > 1:
                    feat = self.patch_aggregate(feat)
                out["sam"] = feat
            if "clipseg" in get_feature and cfg.use_clipseg_feature:
                out["clipseg"] = torch.sum(sam_w[..., None] * fo["clipseg"], dim=-2)
            if return_intermediates:
                out["_sam_weights"], out["_best_ids"] = sam_w, best
        if return_intermediates:
            out.update(
                _w0=w0, _dens0=dens0, _bins1=bins1, _eu1=eu1, _density=density, _rgb_s=rgb_s, _weights=weights, _eu0=eu0
            )
        return out

    def patch_aggregate(self, feat: torch.Tensor) -> torch.Tensor:
        """Patch conv head - samnerf/sam_model.py:202-208,260-265: ``[P*p*p, C]`` patch-major rows ->
        Conv3x3(pad 1) -> ReLU -> Conv3x3(pad 1) -> mean over the p x p patch -> ``[P, C]``."""
        p_ = self.cfg.patch_size
        x = feat.reshape(-1, p_, p_, feat.shape[-1]).permute(0, 3, 1, 2)
        pad = (self.cfg.kernel_size - 1) // 2
        x = F.conv2d(x, self.p["conv_head.0.weight"], self.p["conv_head.0.bias"], padding=pad)
        x = F.relu(x)
        x = F.conv2d(x, self.p["conv_head.2.weight"], self.p["conv_head.2.bias"], padding=pad)
        return x.mean(dim=[2, 3])

    # ---- whole image ----------------------------------------------------------------------
    def render_image(
        self, origins: torch.Tensor, directions: torch.Tensor, fast: bool = False, chunk: Optional[int] = None
    ) -> Dict[str, torch.Tensor]:
        """``SAMModel.get_outputs_for_camera_ray_bundle`` up to the prompt handling -
        samnerf/sam_model.py:354-418 (``get_row_major_sliced_ray_bundle`` rays.py:213-224).
        ``origins/directions[H,W,3]``."""
        cfg = self.cfg
        chunk = chunk or cfg.eval_num_rays_per_chunk
        h, w = origins.shape[:2]
        fo, fd = origins.reshape(-1, 3), directions.reshape(-1, 3)
        lists: Dict[str, List[torch.Tensor]] = {}
        for i in range(0, h * w, chunk):  # LOOP A
            r = self.render_rays(fo[i : i + chunk], fd[i : i + chunk], get_feature=(), fast=fast)
            for k_, v in r.items():
                lists.setdefault(k_, []).append(v)
        out = {k_: torch.cat(v).view(h, w, -1) for k_, v in lists.items()}
        if cfg.distill_sam:
            fh, fw = get_feature_size(h, w)
            p_ = cfg.patch_size
            hi = torch.linspace(0, h - 1, fh * p_, dtype=torch.long)
            wi = torch.linspace(0, w - 1, fw * p_, dtype=torch.long)
            hind, wind = torch.meshgrid(hi, wi, indexing="ij")
            sel_o = origins[hind.flatten(), wind.flatten()].reshape(fh, p_, fw, p_, 3).transpose(1, 2).reshape(-1, 3)
            sel_d = directions[hind.flatten(), wind.flatten()].reshape(fh, p_, fw, p_, 3).transpose(1, 2).reshape(-1, 3)
            feats = []
            for i in range(0, sel_o.shape[0], chunk):  # LOOP B
                feats.append(self.render_rays(sel_o[i : i + chunk], sel_d[i : i + chunk], get_feature=("sam",))["sam"])
            out["sam"] = torch.cat(feats).view(fh, fw, -1)
            if cfg.use_clipseg_feature:  # LOOP C
                hi = torch.linspace(0, h - 1, 32, dtype=torch.long)
                wi = torch.linspace(0, w - 1, 32, dtype=torch.long)
                hind, wind = torch.meshgrid(hi, wi, indexing="ij")
                co = origins[hind.flatten(), wind.flatten()]
                cd = directions[hind.flatten(), wind.flatten()]
                feats = []
                for i in range(0, co.shape[0], chunk):
                    feats.append(self.render_rays(co[i : i + chunk], cd[i : i + chunk], get_feature=("clipseg",))["clipseg"])
                out["clipseg"] = torch.cat(feats).view(32, 32, -1)
        return out
