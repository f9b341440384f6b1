"""The reference's nerfstudio-facing surface for the hot path, backed by ``libsnrf``.

Same class names, argument meaning and return keys as the reference modules they replace, so that
``samnerf.train`` / the viewer / ``scripts/render.py`` call sites read the same:

    reference                                                   here
    nerfstudio/cameras/rays.py:31,97,166                        Frustums, RaySamples, RayBundle
    nerfstudio/model_components/scene_colliders.py:170-188      NearFarCollider
    nerfstudio/model_components/ray_samplers.py:509-599         ProposalNetworkSampler
    nerfstudio/fields/density_fields.py:39-125                  HashMLPDensityField
    nerfstudio/fields/nerfacto_field.py:67-351                  TCNNNerfactoField
    samnerf/sam_field.py:25-140                                 SAMField
    nerfstudio/model_components/renderers.py:58,197,226         RGBRenderer, AccumulationRenderer, DepthRenderer
    samnerf/sam_model.py:126,179                                MeanRenderer, SAMModel

Every method that computes something calls into the CUDA library (through ``renderer.Renderer``); nothing here
falls back to PyTorch math.  Inference (eval-mode) semantics everywhere; the first training slice (SURVEY.md section 8
f-1) is ``SAMModel.train()``: the ``sam_field`` parameter group is trained through ``snrf_feature_forward`` /
``snrf_feature_backward`` behind a ``torch.autograd.Function`` while the geometry (proposal + nerfacto fields) stays
frozen and deterministic (no jitter); their backward is the remaining part of that row.
"""
from __future__ import annotations

import contextlib
from dataclasses import dataclass, fields
from enum import Enum
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch

from .config import SAMNeRFConfig, get_feature_size
from .renderer import Camera, Renderer


class FieldHeadNames(Enum):
    """nerfstudio/field_components/field_heads.py (subset used on this path)."""

    RGB = "rgb"
    DENSITY = "density"


# ------------------------------------------------------------------------------------------------
# data carriers
# ------------------------------------------------------------------------------------------------
class _Carrier:
    """Minimal stand-in for ``TensorDataclass`` (nerfstudio/utils/tensor_dataclass.py): index / reshape /
    flatten applied to every tensor field over the leading (batch) dims."""

    def _map(self, fn):
        kw = {}
        for f in fields(self):
            v = getattr(self, f.name)
            if torch.is_tensor(v):
                kw[f.name] = fn(v)
            elif isinstance(v, _Carrier):
                kw[f.name] = v._map(fn)
            else:
                kw[f.name] = v
        return type(self)(**kw)

    @property
    def shape(self):
        return tuple(self._lead().shape[:-1])

    def __len__(self):
        t = self._lead()
        return t.numel() // t.shape[-1]

    def flatten(self):
        return self._map(lambda t: t.reshape(-1, t.shape[-1]))

    def reshape(self, shape):
        return self._map(lambda t: t.reshape(*shape, t.shape[-1]))

    def __getitem__(self, idx):
        if not isinstance(idx, tuple):
            idx = (idx,)
        return self._map(lambda t: t[idx + (slice(None),)] if Ellipsis not in idx else t[idx])

    def to(self, device):
        return self._map(lambda t: t.to(device))

    def _apply_fn_to_fields(self, fn, dataclass_fn=None):
        return self._map(fn)


@dataclass
class Frustums(_Carrier):
    origins: torch.Tensor
    directions: torch.Tensor
    starts: torch.Tensor
    ends: torch.Tensor
    pixel_area: Optional[torch.Tensor] = None

    def _lead(self):
        return self.starts

    def get_positions(self) -> torch.Tensor:
        """rays.py:48-57 - frustum centres (pure indexing arithmetic on the carrier; kernels recompute it)."""
        return self.origins + self.directions * (self.starts + self.ends) / 2


@dataclass
class RaySamples(_Carrier):
    frustums: Frustums
    camera_indices: Optional[torch.Tensor] = None
    deltas: Optional[torch.Tensor] = None
    spacing_starts: Optional[torch.Tensor] = None
    spacing_ends: Optional[torch.Tensor] = None
    spacing_to_euclidean_fn: Optional[Callable] = None
    renderer: Optional[Renderer] = None

    def _lead(self):
        return self.frustums.starts

    def get_weights(self, densities: torch.Tensor) -> torch.Tensor:
        """rays.py:141-163 -> ``snrf_ray_op`` mode 0."""
        n, s = densities.shape[0], densities.shape[1]
        if densities.requires_grad and torch.is_grad_enabled():
            return _GetWeightsFn.apply(densities.reshape(n, s), self.deltas.reshape(n, s), self.renderer)[..., None]
        w = self.renderer.ray_op(0, self.deltas.reshape(n, s), densities.reshape(n, s))
        return w[..., None]


@dataclass
class RayBundle(_Carrier):
    origins: torch.Tensor
    directions: torch.Tensor
    pixel_area: Optional[torch.Tensor] = None
    camera_indices: Optional[torch.Tensor] = None
    nears: Optional[torch.Tensor] = None
    fars: Optional[torch.Tensor] = None

    def _lead(self):
        return self.origins

    def get_row_major_sliced_ray_bundle(self, start_idx: int, end_idx: int) -> "RayBundle":
        """rays.py:213-224."""
        return self.flatten()[start_idx:end_idx]


class NearFarCollider:
    """scene_colliders.py:170-188 (eval mode: near plane 0).  Bundles that already carry nears/fars pass through
    (scene_colliders.py:40-44)."""

    def __init__(self, near_plane: float, far_plane: float, training: bool = False):
        self.near_plane, self.far_plane, self.training = near_plane, far_plane, training

    def __call__(self, ray_bundle: RayBundle) -> RayBundle:
        if ray_bundle.nears is not None and ray_bundle.fars is not None:
            return ray_bundle
        ones = torch.ones_like(ray_bundle.origins[..., 0:1])
        ray_bundle.nears = ones * (self.near_plane if self.training else 0)
        ray_bundle.fars = ones * self.far_plane
        return ray_bundle


# ------------------------------------------------------------------------------------------------
# fields
# ------------------------------------------------------------------------------------------------
class _TcnnParams(torch.nn.Module):
    """Parameter holder standing in for a tinycudann module (``tcnn.Encoding`` / ``tcnn.Network`` /
    ``tcnn.NetworkWithInputEncoding``): ONE flat fp32 ``params`` tensor - network weights first, then the grid - which
    is all such a module contributes to a ``state_dict`` (sam_field.py:51-109, nerfacto_field.py:144-175,228-240,
    density_fields.py:92-100).  The arithmetic lives in libsnrf, which keeps its own packed fp16 copy."""

    def __init__(self, n: int, device=None):
        super().__init__()
        self.params = torch.nn.Parameter(torch.zeros(int(n), device=device))


_AABB = ((-1.0, -1.0, -1.0), (1.0, 1.0, 1.0))  # scene box of the nerfstudio dataparsers (scene_box.py)


class _Field(torch.nn.Module):
    """Common part of the field shims: ``nn.Module``s whose parameters / buffers carry the reference's names, so that
    ``state_dict()`` / ``load_state_dict(strict=True)`` interoperate with the reference's checkpoints."""

    def __init__(self, renderer: Renderer):
        super().__init__()
        self.renderer = renderer

    def _grid_buffers(self, g) -> None:
        """The geometry the reference's fields register as buffers (nerfacto_field.py:121-125, density_fields.py:66-71)."""
        dev = self.renderer.device if self.renderer is not None else None
        self.register_buffer("aabb", torch.tensor(_AABB, device=dev))
        self.register_buffer("max_res", torch.tensor(g.max_resolution, device=dev))
        self.register_buffer("num_levels", torch.tensor(g.n_levels, device=dev))
        self.register_buffer("log2_hashmap_size", torch.tensor(g.log2_hashmap_size, device=dev))

    def _differentiable(self) -> bool:
        return self.training and torch.is_grad_enabled()

    def density_fn(self, positions: torch.Tensor) -> torch.Tensor:
        """base_field.py:38-56."""
        return self._density(positions)[0]


class HashMLPDensityField(_Field):
    """density_fields.py:39-125 (the proposal network)."""

    def __init__(self, renderer: Renderer, cfg: Optional[SAMNeRFConfig] = None):
        super().__init__(renderer)
        cfg = cfg or renderer.cfg
        self.mlp_base = _TcnnParams(cfg.proposal_mlp_params + cfg.proposal_grid.n_params, renderer.device)
        self._grid_buffers(cfg.proposal_grid)

    def _density(self, positions):
        if self._differentiable():
            return _ProposalDensityFn.apply(self.mlp_base.params, self.renderer, positions.detach()), None
        return self.renderer.query_density("proposal", positions)

    def get_density(self, ray_samples: RaySamples):
        return self._density(ray_samples.frustums.get_positions())[0], None

    def get_outputs(self, ray_samples, density_embedding=None) -> dict:
        return {}

    def forward(self, ray_samples, compute_normals: bool = False):
        density, _ = self.get_density(ray_samples)
        return {FieldHeadNames.DENSITY: density}


class TCNNNerfactoField(_Field):
    """nerfacto_field.py:67-351 with appearance embedding off (samconfigs.py:80,134)."""

    def __init__(self, renderer: Renderer, cfg: Optional[SAMNeRFConfig] = None):
        super().__init__(renderer)
        cfg = cfg or renderer.cfg
        dev = renderer.device
        self.mlp_base = _TcnnParams(cfg.field_mlp_params + cfg.field_grid.n_params, dev)
        self.mlp_head = _TcnnParams(cfg.head_mlp_params, dev)
        # parameter-free tcnn encodings: they still show up in the reference's state_dict, as empty tensors
        self.direction_encoding = _TcnnParams(0, dev)
        self.position_encoding = _TcnnParams(0, dev)
        self._grid_buffers(cfg.field_grid)

    def _density(self, positions):
        return self.renderer.query_density("field", positions)

    def get_density(self, ray_samples: RaySamples):
        return self._density(ray_samples.frustums.get_positions())

    def get_outputs(self, ray_samples: RaySamples, density_embedding=None):
        assert density_embedding is not None
        if ray_samples.camera_indices is None:
            raise AttributeError("Camera indices are not provided.")  # nerfacto_field.py:273-275
        rgb = self.renderer.query_rgb(ray_samples.frustums.directions, density_embedding)
        return {FieldHeadNames.RGB: rgb}

    def forward(self, ray_samples: RaySamples, compute_normals: bool = False):
        if self._differentiable():
            # training: density and colour in one differentiable call (tinycudann's autograd in the reference)
            if ray_samples.camera_indices is None:
                raise AttributeError("Camera indices are not provided.")  # nerfacto_field.py:273-275
            density, rgb = _NerfactoFieldFn.apply(self.mlp_base.params, self.mlp_head.params, self.renderer,
                                                  ray_samples.frustums.get_positions().detach(),
                                                  ray_samples.frustums.directions.detach())
            return {FieldHeadNames.RGB: rgb, FieldHeadNames.DENSITY: density}
        density, emb = self.get_density(ray_samples)
        out = self.get_outputs(ray_samples, density_embedding=emb)
        out[FieldHeadNames.DENSITY] = density
        return out


class SAMField(_Field):
    """sam_field.py:25-140.  Constructor arguments in the reference's order (sam_field.py:26-35); ``renderer`` is the one
    addition.  ``get_feautre`` keeps the reference's spelling of the keyword."""

    def __init__(self, grid_layers, grid_sizes, grid_resolutions, hidden_layers=2, spatial_distortion=None,
                 use_dino_features: bool = False, use_clipseg_features: bool = False, renderer: Optional[Renderer] = None):
        super().__init__(renderer)
        assert len(grid_layers) == len(grid_sizes) and len(grid_resolutions) == len(grid_layers)  # sam_field.py:37
        if renderer is None:
            raise RuntimeError("SAMField needs renderer=Renderer(...): the arithmetic lives in libsnrf (no PyTorch fallback)")
        if hidden_layers != 1 or use_dino_features:
            raise ValueError("libsnrf builds the shipped configuration: hidden_layers = 1, no DINO head (samconfigs.py:81-83)")
        from .config import GridConfig

        self.spatial_distortion = spatial_distortion  # L2 scene contraction inside the kernels (sam_field.py:32)
        self.use_dino_features, self.use_clipseg_features = use_dino_features, use_clipseg_features
        dev = renderer.device
        grids = [GridConfig(int(l), 8, int(t), int(r[0]), int(r[1])) for l, t, r in zip(grid_layers, grid_sizes, grid_resolutions)]
        width = sum(g.n_levels * g.n_features for g in grids)
        self.clip_encs = torch.nn.ModuleList([_TcnnParams(g.n_params, dev) for g in grids])
        self.sam_net = _TcnnParams(256 * width + 256 * 256, dev)
        if use_clipseg_features:
            self.clipseg_encs = torch.nn.ModuleList([_TcnnParams(g.n_params, dev) for g in grids])
            self.clipseg_net = _TcnnParams(256 * width + 192 * 256, dev)

    def get_outputs(self, ray_samples: RaySamples, get_feautre=("sam", "dino", "clipseg")):
        pos = ray_samples.frustums.get_positions().detach()
        out = {}
        if "sam" in get_feautre:
            out["hashgrid"], out["sam"] = self.renderer.query_features("sam", pos)
        if "clipseg" in get_feautre and self.renderer.cfg.use_clipseg_feature:
            _, out["clipseg"] = self.renderer.query_features("clipseg", pos)
        return out

    def forward(self, ray_samples: RaySamples, get_feautre=("sam", "dino", "clipseg")):
        return self.get_outputs(ray_samples, get_feautre=get_feautre)


# ------------------------------------------------------------------------------------------------
# sampler
# ------------------------------------------------------------------------------------------------
class ProposalNetworkSampler:
    """ray_samplers.py:509-599, eval mode, one proposal iteration (samconfigs.py:84,138)."""

    def __init__(self, renderer: Renderer, train_stratified: bool = True, update_sched: Callable = lambda step: 1):
        self.renderer = renderer
        self.training = False             # set by SAMModel.train()
        self.train_stratified = train_stratified  # ray_samplers.py:67
        self.update_sched = update_sched  # ray_samplers.py:528,533 (nerfacto.py:196-200 builds the warm-up schedule)
        self.last_jitter: Optional[torch.Tensor] = None
        self._anneal, self._steps_since_update, self._step = 1.0, 0, 0

    def set_anneal(self, anneal: float) -> None:
        """ray_samplers.py:546-548."""
        self._anneal = float(anneal)
        self.renderer.set_anneal(self._anneal)

    def step_cb(self, step) -> None:
        """ray_samplers.py:550-553."""
        self._step = step
        self._steps_since_update += 1

    def _samples(self, bundle: RayBundle, edges: torch.Tensor, spacing: Optional[torch.Tensor]) -> RaySamples:
        o, d = bundle.origins[:, None, :], bundle.directions[:, None, :]
        starts, ends = edges[:, :-1, None], edges[:, 1:, None]
        fr = Frustums(origins=o.expand(-1, starts.shape[1], -1), directions=d.expand(-1, starts.shape[1], -1),
                      starts=starts, ends=ends,
                      pixel_area=None if bundle.pixel_area is None else bundle.pixel_area[:, None, :].expand(-1, starts.shape[1], -1))
        cam = None if bundle.camera_indices is None else bundle.camera_indices[:, None, :].expand(-1, starts.shape[1], -1)
        return RaySamples(frustums=fr, camera_indices=cam, deltas=ends - starts,
                          spacing_starts=None if spacing is None else spacing[:, :-1, None],
                          spacing_ends=None if spacing is None else spacing[:, 1:, None], renderer=self.renderer)

    def generate_ray_samples(self, ray_bundle: RayBundle, density_fns=None) -> Tuple[RaySamples, List, List]:
        r = self.renderer
        # training: stratified sampling with one draw per ray and level (single_jitter, nerfacto.py:113,211) -
        # drawn here with torch.rand like the reference (ray_samplers.py:104-112,314-322), consumed by the kernel
        jitter = None
        if self.training and self.train_stratified:
            n_rays = ray_bundle.origins.reshape(-1, 3).shape[0]
            jitter = torch.rand((n_rays, 2), device=r.device)
        self.last_jitter = jitter
        w0, edges1, _ = r.sample(ray_bundle.origins, ray_bundle.directions, ray_bundle.nears, ray_bundle.fars, jitter=jitter)
        n = w0.shape[0]
        # level-0 samples for ray_samples_list: the piecewise bins are a closed form of (near, far)
        nears = ray_bundle.nears if ray_bundle.nears is not None else torch.zeros(n, 1, device=w0.device)
        fars = ray_bundle.fars if ray_bundle.fars is not None else torch.full((n, 1), r.cfg.far_plane, device=w0.device)
        bins = torch.linspace(0.0, 1.0, r.cfg.num_proposal_samples + 1, device=w0.device)[None]
        if jitter is not None:  # ray_samplers.py:108-111
            centers = (bins[..., 1:] + bins[..., :-1]) / 2.0
            upper, lower = torch.cat([centers, bins[..., -1:]], -1), torch.cat([bins[..., :1], centers], -1)
            bins = lower + (upper - lower) * jitter[:, 0:1].to(w0.device)
        sp = lambda x: torch.where(x < 1, x / 2, 1 - 1 / (2 * x))  # noqa: E731  ray_samplers.py:242
        sp_inv = lambda x: torch.where(x < 0.5, 2 * x, 1 / (2 - 2 * x))  # noqa: E731  ray_samplers.py:243
        s_near, s_far = sp(nears.to(w0.device)), sp(fars.to(w0.device))
        edges0 = sp_inv(bins * s_far + (1 - bins) * s_near)
        rs0 = self._samples(ray_bundle, edges0, bins.expand(n, -1) if bins.shape[0] == 1 else bins)
        # level-1 spacing bins (what interlevel_loss / distortion_loss read through ray_samples_to_sdist,
        # model_components/losses.py:100-143): the kernel reports euclidean edges; the spacing transform is invertible
        spacing1 = (sp(edges1) - s_near) / (s_far - s_near)
        rs1 = self._samples(ray_bundle, edges1, spacing1)
        to_euclidean = lambda x: sp_inv(x * s_far + (1 - x) * s_near)  # noqa: E731  ray_samplers.py:115
        rs0.spacing_to_euclidean_fn = rs1.spacing_to_euclidean_fn = to_euclidean
        if self.training and density_fns and torch.is_grad_enabled():
            # training (ray_samplers.py:572,586-596): on "updated" steps the proposal weights handed to the interlevel
            # loss carry the gradient of the proposal network; the sample positions are detached (ray_samplers.py:357)
            updated = self._steps_since_update > self.update_sched(self._step) or self._step < 10
            if updated:
                self._steps_since_update = 0
                dens0 = density_fns[0](rs0.frustums.get_positions())
                if dens0.requires_grad:
                    return rs1, [rs0.get_weights(dens0)], [rs0]
        return rs1, [w0[..., None]], [rs0]

    __call__ = generate_ray_samples
    forward = generate_ray_samples


# ------------------------------------------------------------------------------------------------
# renderers
# ------------------------------------------------------------------------------------------------
BACKGROUND_COLOR_OVERRIDE: Optional[torch.Tensor] = None


@contextlib.contextmanager
def background_color_override_context(mode):
    """renderers.py:46-55."""
    global BACKGROUND_COLOR_OVERRIDE
    old = BACKGROUND_COLOR_OVERRIDE
    try:
        BACKGROUND_COLOR_OVERRIDE = mode
        yield
    finally:
        BACKGROUND_COLOR_OVERRIDE = old


_COLORS = {"white": (1.0, 1.0, 1.0), "black": (0.0, 0.0, 0.0)}


def _resolve_background(background_color):
    bg = BACKGROUND_COLOR_OVERRIDE if BACKGROUND_COLOR_OVERRIDE is not None else background_color
    if isinstance(bg, str):
        if bg == "last_sample":
            return None
        if bg in _COLORS:
            return _COLORS[bg]
        raise NotImplementedError(f"background {bg!r}: random backgrounds are a training-only mode")
    return [float(v) for v in bg]


class RGBRenderer:
    """renderers.py:58-140 (eval: nan_to_num + clamp)."""

    def __init__(self, renderer: Renderer, background_color="last_sample"):
        self.renderer, self.background_color = renderer, background_color

    def forward(self, rgb, weights, ray_indices=None, num_rays=None):
        if ray_indices is not None:
            raise NotImplementedError("packed samples are not on this path (renderers.py:90-95)")
        n, s = weights.shape[0], weights.shape[1]
        bg = _resolve_background(self.background_color)
        if (rgb.requires_grad or weights.requires_grad) and torch.is_grad_enabled():
            # training: no nan_to_num / clamp (renderers.py:132-139); sum_i w_i <= 1 keeps the value in [0, 1] anyway
            return _RGBCompositeFn.apply(rgb.reshape(n, s, 3), weights.reshape(n, s), self.renderer, bg)
        return self.renderer.ray_op(3, rgb.reshape(n, s, 3), weights.reshape(n, s), background=bg)

    __call__ = forward


class AccumulationRenderer:
    """renderers.py:197-223."""

    def __init__(self, renderer: Renderer):
        self.renderer = renderer

    def forward(self, weights, ray_indices=None, num_rays=None):
        return self.renderer.ray_op(1, weights.reshape(weights.shape[0], weights.shape[1]))

    __call__ = forward


class DepthRenderer:
    """renderers.py:226-270, method "median"."""

    def __init__(self, renderer: Renderer, method: str = "median"):
        if method != "median":
            raise NotImplementedError("SAMModel uses the median depth (nerfacto.py:222)")
        self.renderer = renderer

    def forward(self, weights, ray_samples: RaySamples, ray_indices=None, num_rays=None):
        n, s = weights.shape[0], weights.shape[1]
        return self.renderer.ray_op(2, weights.reshape(n, s), ray_samples.frustums.starts.reshape(n, s),
                                    ray_samples.frustums.ends.reshape(n, s))

    __call__ = forward


class MeanRenderer:
    """sam_model.py:126-137."""

    def __init__(self, renderer: Renderer):
        self.renderer = renderer

    def forward(self, embeds, weights):
        n, s, c = embeds.shape
        return self.renderer.ray_op(4, embeds.to(torch.float32), weights.reshape(n, s), n_channels=c)

    __call__ = forward


# ------------------------------------------------------------------------------------------------
# training side: autograd Functions over libsnrf's forward / backward kernels.  The flat parameters
# are inputs only so that autograd routes their gradients; the values used are the ones uploaded to the library
# (``SAMModel._sync_params``).
# ------------------------------------------------------------------------------------------------
class _ProposalDensityFn(torch.autograd.Function):
    """``HashMLPDensityField.density_fn`` (density_fields.py:102-125)."""

    @staticmethod
    def forward(ctx, params, renderer, positions):
        density, _ = renderer.query_density("proposal", positions)
        ctx.renderer = renderer
        ctx.save_for_backward(positions)
        return density

    @staticmethod
    def backward(ctx, d_density):
        (positions,) = ctx.saved_tensors
        g = ctx.renderer.field_backward("proposal", positions, d_density=d_density.contiguous())
        return g["base"], None, None


class _NerfactoFieldFn(torch.autograd.Function):
    """``TCNNNerfactoField.forward`` (nerfacto_field.py:242-351): density and rgb of every sample."""

    @staticmethod
    def forward(ctx, base, head, renderer, positions, directions):
        density, geo = renderer.query_density("field", positions)
        rgb = renderer.query_rgb(directions, geo)
        ctx.renderer = renderer
        ctx.save_for_backward(positions, directions)
        return density, rgb

    @staticmethod
    def backward(ctx, d_density, d_rgb):
        positions, directions = ctx.saved_tensors
        g = ctx.renderer.field_backward("field", positions, directions,
                                        d_density=None if d_density is None else d_density.contiguous(),
                                        d_rgb=None if d_rgb is None else d_rgb.contiguous())
        return g["base"], g.get("head"), None, None, None


class _GetWeightsFn(torch.autograd.Function):
    """``RaySamples.get_weights`` (rays.py:141-163); deltas are detached (ray_samplers.py:357)."""

    @staticmethod
    def forward(ctx, densities, deltas, renderer):
        ctx.renderer = renderer
        ctx.save_for_backward(densities, deltas)
        return renderer.ray_op(0, deltas, densities)

    @staticmethod
    def backward(ctx, g_w):
        densities, deltas = ctx.saved_tensors
        return ctx.renderer.ray_op_backward(0, deltas, densities, g_w.contiguous()), None, None


class _RGBCompositeFn(torch.autograd.Function):
    """``RGBRenderer.combine_rgb`` (renderers.py:69-112)."""

    @staticmethod
    def forward(ctx, rgb, weights, renderer, background):
        ctx.renderer, ctx.background = renderer, background
        ctx.save_for_backward(rgb, weights)
        return renderer.ray_op(3, rgb, weights, background=background)

    @staticmethod
    def backward(ctx, g_out):
        rgb, weights = ctx.saved_tensors
        d_rgb, d_w = ctx.renderer.ray_op_backward(3, rgb, weights, g_out.contiguous(), background=ctx.background)
        return d_rgb, d_w, None, None


class _PatchAggregateFn(torch.autograd.Function):
    """``conv_head(feat).mean(dim=[2, 3])`` over 4 x 4 ray patches (sam_model.py:202-208,260-265)."""

    @staticmethod
    def forward(ctx, feat, w0, b0, w2, b2, renderer):
        ctx.renderer = renderer
        ctx.save_for_backward(feat)
        return renderer.patch_aggregate(feat)

    @staticmethod
    def backward(ctx, d_out):
        (feat,) = ctx.saved_tensors
        g, d_feat = ctx.renderer.patch_aggregate_backward(feat, d_out.contiguous(), want_d_feat=ctx.needs_input_grad[0])
        return (d_feat,) + tuple(g[n] for n in Renderer.CONV_PARAMS) + (None,)


class _FeatureBranchFn(torch.autograd.Function):
    """``MeanRenderer(SAMField.get_outputs(sam_samples))`` (sam_model.py:256-277, sam_field.py:112-140)."""

    @staticmethod
    def forward(ctx, net, grid0, grid1, renderer, which, origins, directions, sam_t, sam_w):
        out, enc = renderer.feature_forward(which, origins, directions, sam_t, sam_w, save_for_backward=True)
        ctx.renderer, ctx.which = renderer, which
        ctx.save_for_backward(origins, directions, sam_t, sam_w, enc)
        return out

    @staticmethod
    def backward(ctx, d_out):
        origins, directions, sam_t, sam_w, enc = ctx.saved_tensors
        want = [k for k, need in zip(("net", "grid0", "grid1"), ctx.needs_input_grad[:3]) if need]
        g = ctx.renderer.feature_backward(ctx.which, origins, directions, sam_t, sam_w, enc, d_out.contiguous(), want=want)
        return g.get("net"), g.get("grid0"), g.get("grid1"), None, None, None, None, None, None


# ------------------------------------------------------------------------------------------------
# the model
# ------------------------------------------------------------------------------------------------
class SAMModel(torch.nn.Module):
    """samnerf/sam_model.py:179-418, inference side (+ the training shim of SURVEY 8 f-1).

    An ``nn.Module`` whose parameter tree is the reference's (``SAMModel(NerfactoModel)``, sam_model.py:179-224;
    nerfacto.py:149-225): ``proposal_networks.0.mlp_base.params``, ``field.mlp_base.params``, ``field.mlp_head.params``,
    ``sam_field.{clip_encs.i,sam_net,clipseg_encs.i,clipseg_net}.params``, ``conv_head.{0,2}.{weight,bias}`` plus the
    geometry buffers - so ``state_dict()`` / ``load_state_dict(strict=True)`` exchange checkpoints with the reference
    (tests/golden/state_dict_layout.json holds the reference modules' own key list).  The tensors are fp32 masters on the
    device; libsnrf keeps packed fp16 copies that are re-uploaded whenever a master's version counter moves.

    ``get_outputs`` / ``forward`` run one fused ``snrf_render`` call per chunk; the component objects
    (``proposal_sampler``, ``field``, ``sam_field``, renderers) expose the same pieces individually.
    Prompt lifting / projection (sam_model.py:426-475) is done here as in the reference; the 2-D mask decoders that
    consume the prompts and the rendered feature map (sam_model.py:485-548) are out of scope.
    Constructed in eval mode (a bare ``nn.Module`` starts in training mode; the reference's pipelines call ``eval()`` /
    ``train()`` explicitly, and so may callers of this class)."""

    def __init__(self, config: SAMNeRFConfig, device: int = 0, engine: str = "tcgen05"):
        super().__init__()
        self.config = config
        self.renderer = Renderer(config, device=device, engine=engine)
        r = self.renderer
        self.collider = NearFarCollider(near_plane=0.05, far_plane=config.far_plane)
        self.proposal_networks = torch.nn.ModuleList([HashMLPDensityField(r, config)])
        self.density_fns = [n.density_fn for n in self.proposal_networks]
        from .training import proposal_update_schedule

        self.proposal_sampler = ProposalNetworkSampler(
            r, update_sched=proposal_update_schedule(config.proposal_warmup, config.proposal_update_every))
        self.field = TCNNNerfactoField(r, config)
        if config.distill_sam:
            gs = config.sam_grids
            self.sam_field = SAMField(
                tuple(g.n_levels for g in gs), tuple(g.log2_hashmap_size for g in gs),
                tuple((g.base_resolution, g.max_resolution) for g in gs), hidden_layers=1,
                use_dino_features=False, use_clipseg_features=config.use_clipseg_feature, renderer=r)
            k = config.kernel_size  # sam_model.py:202-208: parameter holders; the convolutions run in libsnrf
            self.conv_head = torch.nn.Sequential(
                torch.nn.Conv2d(256, 256, k, stride=1, padding=(k - 1) // 2), torch.nn.ReLU(inplace=True),
                torch.nn.Conv2d(256, 256, k, stride=1, padding=(k - 1) // 2)).to(r.device)
        else:
            self.sam_field = None
        self.renderer_rgb = RGBRenderer(r, background_color="last_sample")
        self.renderer_accumulation = AccumulationRenderer(r)
        self.renderer_depth = DepthRenderer(r)
        self.renderer_mean = MeanRenderer(r)
        self.prompts = None
        self._uploaded: Dict[str, int] = {name: p._version for name, p in self.params.items()}
        self.train(False)

    @property
    def params(self) -> Dict[str, torch.nn.Parameter]:
        """The hot-path parameters under the reference's names (the empty tcnn encodings are left out; the conv head
        only when the patch head is in use, as before)."""
        from .checkpoint import HOT_PATH_KEYS

        use_conv = self.config.distill_sam and self.config.patch_size > 1
        return {n: p for n, p in self.named_parameters()
                if n in HOT_PATH_KEYS and p.numel() > 0 and (use_conv or not n.startswith("conv_head."))}

    def load_state_dict(self, state_dict: Dict[str, torch.Tensor], strict: bool = False):
        """Accepts the reference's pipeline keys (with or without the ``module.`` / ``_model.`` prefixes,
        base_pipeline.py:109-115,366-375).  ``strict=True`` is ``nn.Module``'s own check against this model's key set -
        what the reference's trainer enforces; the default ignores tensors off the hot path (camera optimiser, ...)."""
        from .checkpoint import params_from_state_dict, strip_prefixes

        params_from_state_dict(state_dict)  # raises KeyError when a required hot-path tensor is missing
        own = super().state_dict()
        mapped = {}
        for k, v in state_dict.items():
            name = strip_prefixes(k)
            if name in own and torch.is_tensor(v):
                v = v.detach()
                mapped[name] = v.reshape(own[name].shape) if v.numel() == own[name].numel() else v
            elif strict:
                mapped[name] = v
        if not strict:  # tensors this configuration does not use (e.g. an unused conv head) keep their initial values
            mapped = {k: v for k, v in mapped.items() if k in own}
        res = super().load_state_dict(mapped, strict=strict)
        self.renderer.load_params({n: p.detach() for n, p in self.named_parameters() if p.numel() > 0 and n in mapped})
        self._uploaded = {name: p._version for name, p in self.params.items()}
        return res

    @classmethod
    def from_checkpoint(cls, path: str, device: int = 0, engine: str = "tcgen05", base: Optional[SAMNeRFConfig] = None):
        """Build the model from a reference training checkpoint (``step-*.ckpt`` file or its directory; layout
        nerfstudio/engine/trainer.py:389-400): the configuration is inferred from the tensors it carries."""
        from .checkpoint import load_checkpoint

        cfg, params, step = load_checkpoint(path, base)
        model = cls(cfg, device=device, engine=engine)
        model.load_state_dict(params)
        model.step = step
        return model

    # ---- training -------------------------------------------------------------------
    def train(self, mode: bool = True):
        """Training mode: the gradients of the flat fp32 parameters come from libsnrf's backward kernels
        (``snrf_field_backward``, ``snrf_feature_backward``, ``snrf_patch_aggregate_backward``, ``snrf_ray_op_backward``).
        The collider switches to its training near plane (scene_colliders.py:185) and the sampler to stratified
        single-jitter sampling (ray_samplers.py:104-112,314-322).  Leaving training mode pushes every parameter an
        optimiser has changed into the library's packed copies."""
        super().train(mode)
        if hasattr(self, "collider"):
            self.collider.training = bool(mode)
            self.proposal_sampler.training = bool(mode)
            if not mode and hasattr(self, "_uploaded"):
                self._sync_params()
        return self

    def get_param_groups(self) -> Dict[str, List[torch.nn.Parameter]]:
        """nerfacto.py:236-240 + sam_model.py:330-335."""
        p = self.params
        groups = {
            "proposal_networks": [v for k, v in p.items() if k.startswith("proposal_networks.")],
            "fields": [v for k, v in p.items() if k.startswith("field.")],
        }
        if self.config.distill_sam:
            groups["sam_field"] = [v for k, v in p.items() if k.startswith("sam_field.")]
            conv = [v for k, v in p.items() if k.startswith("conv_head.")]
            if conv:
                groups["conv"] = conv
        return groups

    # nerfacto.py:316-344 + sam_model.py:316-328
    def get_metrics_dict(self, outputs, batch) -> Dict[str, torch.Tensor]:
        from .training import metrics_dict

        return metrics_dict(outputs, batch, self.training)

    def get_loss_dict(self, outputs, batch, metrics_dict=None) -> Dict[str, torch.Tensor]:
        from .training import loss_dict

        if self.training:
            assert metrics_dict is not None and "distortion" in metrics_dict  # nerfacto.py:332
        return loss_dict(outputs, batch, metrics_dict or {}, self.config, self.training)

    # nerfacto.py:242-271: the two per-iteration callbacks of the trainer
    def before_train_iteration(self, step: int) -> None:
        from .training import proposal_anneal

        if self.config.use_proposal_weight_anneal:
            self.proposal_sampler.set_anneal(proposal_anneal(step, self.config.proposal_weights_anneal_max_num_iters,
                                                             self.config.proposal_weights_anneal_slope))

    def after_train_iteration(self, step: int) -> None:
        if self.config.use_proposal_weight_anneal:
            self.proposal_sampler.step_cb(step)

    def all_reduce_gradients(self, group=None) -> None:
        """Data-parallel training as the reference runs it (one process per GPU under torch DDP, samnerf/train.py:171-199;
        DDP averages every gradient across ranks after backward): call between ``backward()`` and ``optimizer.step()``.
        The parameters are a dozen large flat tensors, so one ``all_reduce`` per tensor is already bucketed."""
        import torch.distributed as dist

        if not (dist.is_available() and dist.is_initialized()):
            return
        world = dist.get_world_size(group)
        if world == 1:
            return
        for p in self.params.values():
            if p.grad is not None:
                dist.all_reduce(p.grad, op=dist.ReduceOp.SUM, group=group)
                p.grad.div_(world)

    def get_training_callbacks(self, training_callback_attributes=None) -> List:
        """nerfacto.py:242-271: the anneal setter before, the sampler's step counter after every iteration."""
        from .training import TrainingCallback, TrainingCallbackLocation

        if not self.config.use_proposal_weight_anneal:
            return []
        return [
            TrainingCallback([TrainingCallbackLocation.BEFORE_TRAIN_ITERATION], self.before_train_iteration, update_every_num_iters=1),
            TrainingCallback([TrainingCallbackLocation.AFTER_TRAIN_ITERATION], self.proposal_sampler.step_cb, update_every_num_iters=1),
        ]

    def _sync_params(self) -> None:
        """Push parameters an optimiser step has changed (tensor version counters) back into the library's packed
        fp16 copies - the re-upload SURVEY 8 b asks for."""
        r = self.renderer
        params = self.params
        for name in Renderer.DENSITY_PARAMS:
            p = params.get(name)
            if p is not None and p._version != self._uploaded[name]:
                r.upload_density_params(name, p)
                self._uploaded[name] = p._version
        for which in ("sam", "clipseg"):
            names = Renderer.FEATURE_PARAMS[which]
            changed = {}
            for slot, name in zip(("net", "grid0", "grid1"), names):
                p = params.get(name)
                if p is not None and p._version != self._uploaded[name]:
                    changed[slot] = p
                    self._uploaded[name] = p._version
            if changed:
                r.upload_feature_params(which, **changed)
        conv = [params.get(n) for n in Renderer.CONV_PARAMS]
        if conv[0] is not None and any(p._version != self._uploaded[n] for p, n in zip(conv, Renderer.CONV_PARAMS)):
            r.upload_conv_head(*conv)
            for p, n in zip(conv, Renderer.CONV_PARAMS):
                self._uploaded[n] = p._version

    def _get_outputs_training(self, ray_bundle: RayBundle, get_feature, fast: bool):
        """sam_model.py:226-301 in training mode, component by component like the reference: the samplers' positions
        are detached (ray_samplers.py:357), the two density fields / get_weights / the RGB composite are autograd
        Functions over libsnrf kernels, and the feature branch takes its top-k picks from those same weights."""
        cfg, r = self.config, self.renderer
        self._sync_params()
        if ray_bundle.camera_indices is None:  # the field insists on them (nerfacto_field.py:273-275)
            ray_bundle.camera_indices = torch.zeros_like(ray_bundle.origins[..., :1], dtype=torch.long)
        ray_samples, weights_list, ray_samples_list = self.proposal_sampler(ray_bundle, density_fns=self.density_fns)
        ray_samples_list.append(ray_samples)
        field_outputs = self.field(ray_samples)
        weights = ray_samples.get_weights(field_outputs[FieldHeadNames.DENSITY])
        weights_list.append(weights)
        out = {"rgb": self.renderer_rgb(rgb=field_outputs[FieldHeadNames.RGB], weights=weights)}
        with torch.no_grad():
            out["depth"] = self.renderer_depth(weights=weights, ray_samples=ray_samples)
            if not fast:
                out["accumulation"] = self.renderer_accumulation(weights=weights)
                out["prop_depth_0"] = self.renderer_depth(weights=weights_list[0], ray_samples=ray_samples_list[0])
        out["weights_list"], out["ray_samples_list"] = weights_list, ray_samples_list
        params = self.params
        feats = [f for f in get_feature if Renderer.FEATURE_PARAMS[f][0] in params]
        if feats:
            # top-k + sharpen on the (detached) weights of this very pass (sam_model.py:244-255)
            sam_t, sam_w = r.pick_samples(weights, ray_samples.frustums.starts, ray_samples.frustums.ends)
            o = r._prep(ray_bundle.origins, 3)
            d = r._prep(ray_bundle.directions, 3)
            for which in feats:
                names = Renderer.FEATURE_PARAMS[which]
                feat = _FeatureBranchFn.apply(*[params[n] for n in names], r, which, o, d, sam_t, sam_w)
                if which == "sam" and cfg.patch_size > 1:  # sam_model.py:260-265
                    feat = _PatchAggregateFn.apply(feat, *[params[n] for n in Renderer.CONV_PARAMS], r)
                out[which] = feat
        return out

    # sam_model.py:303-314
    def forward(self, ray_bundle: RayBundle, **kwargs):
        if self.collider is not None:
            ray_bundle = self.collider(ray_bundle)
        return self.get_outputs(ray_bundle, **kwargs)

    # sam_model.py:226-278
    def get_outputs(self, ray_bundle: RayBundle, get_rgbsigma=True, get_feature=("sam", "dino", "clipseg"), fast=False):
        cfg = self.config
        feats = [f for f in get_feature if f in ("sam", "clipseg")] if cfg.distill_sam else []
        if self.training:
            return self._get_outputs_training(ray_bundle, feats, fast)
        bg = _resolve_background(self.renderer_rgb.background_color)
        out = self.renderer.render(
            ray_bundle.origins, ray_bundle.directions, ray_bundle.nears, ray_bundle.fars, get_feature=feats,
            patch=cfg.patch_size > 1 and "sam" in feats, fast=fast, background=bg,
        )
        return out

    # sam_model.py:338-418 (up to the prompt handling)
    @torch.no_grad()
    def get_outputs_for_camera_ray_bundle(self, camera_ray_bundle: RayBundle, points=None, intrin=None, c2w=None,
                                          text_prompt=None, topk=5, thresh=0.5, fast=False):
        cfg = self.config
        chunk = cfg.eval_num_rays_per_chunk
        h, w = camera_ray_bundle.origins.shape[:2]
        dev = self.renderer.device
        bundle = camera_ray_bundle.to(dev)
        flat = bundle.flatten()
        lists: Dict[str, List[torch.Tensor]] = {}
        for i in range(0, h * w, chunk):  # LOOP A
            o = self.forward(flat[i:i + chunk], get_feature=[], fast=fast)
            for k, v in o.items():
                lists.setdefault(k, []).append(v)
        outputs = {k: torch.cat(v).view(h, w, -1) for k, v in lists.items()}
        if cfg.distill_sam:
            fh, fw = get_feature_size(h, w)
            p = cfg.patch_size
            hi = torch.linspace(0, h - 1, fh * p, dtype=torch.long, device=dev)
            wi = torch.linspace(0, w - 1, fw * p, dtype=torch.long, device=dev)
            hind, wind = torch.meshgrid(hi, wi, indexing="ij")
            fb = bundle[hind.flatten(), wind.flatten()].reshape((fh, p, fw, p))
            fb = fb._apply_fn_to_fields(lambda x: x.transpose(1, 2)).flatten()
            feats = []
            for i in range(0, len(fb), chunk):  # LOOP B
                feats.append(self.forward(fb[i:i + chunk], get_feature=["sam"])["sam"])
            outputs["sam"] = torch.cat(feats).view(fh, fw, -1)
            if cfg.use_clipseg_feature:  # LOOP C
                hi = torch.linspace(0, h - 1, 32, dtype=torch.long, device=dev)
                wi = torch.linspace(0, w - 1, 32, dtype=torch.long, device=dev)
                hind, wind = torch.meshgrid(hi, wi, indexing="ij")
                cb = bundle[hind.flatten(), wind.flatten()]
                feats = []
                for i in range(0, len(cb), chunk):
                    feats.append(self.forward(cb[i:i + chunk], get_feature=["clipseg"])["clipseg"])
                outputs["clipseg"] = torch.cat(feats).view(32, 32, -1)
        self._handle_prompts(outputs, points, intrin, c2w, text_prompt)
        return outputs

    def _handle_prompts(self, outputs: Dict[str, torch.Tensor], points, intrin, c2w, text_prompt=None) -> None:
        """The prompt bookkeeping of sam_model.py:426-475 up to the point where the 2-D decoders take over: new clicks are
        lifted to 3-D once (at the rendered depth minus ``TOR``) and remembered in ``self.prompts``; every frame the
        remembered prompts are projected into the current view and the ones inside the image become
        ``outputs["prompt_points"]`` - what the reference hands to ``SamPredictor.predict`` (with
        ``outputs["sam_embedding"]``, the zero-padded ``[1,256,S,S]`` feature map, sam_model.py:485 / predictor.py:100-127).
        ``masked_rgb`` stays the plain rgb, as in the reference before a mask exists (sam_model.py:426)."""
        from . import prompts as P

        outputs["masked_rgb"] = outputs["rgb"]
        if "sam" in outputs:
            outputs["sam_embedding"] = P.pad_feature_map(outputs["sam"])
        self._decode_clipseg(outputs, text_prompt)
        if points is None:
            self.prompts = None
            self._decode_mask(outputs)  # ClipSeg clicks alone can prompt the mask (sam_model.py:509-516)
            return
        assert intrin is not None and c2w is not None
        intrin, c2w = torch.as_tensor(intrin, dtype=torch.float32).cpu(), torch.as_tensor(c2w, dtype=torch.float32).cpu()
        if len(points) > 0:
            pts = torch.as_tensor(points).to(torch.long)
            known = 0 if getattr(self, "prompts", None) is None else self.prompts.shape[0]
            if len(pts) > known:  # only the clicks that arrived since the last frame are lifted (sam_model.py:437-444)
                new = P.lift_points(pts[known:], outputs["depth"].detach().cpu(), intrin, c2w)
                self.prompts = new if known == 0 else torch.cat([self.prompts, new], dim=0)
        else:
            self.prompts = None
        if getattr(self, "prompts", None) is not None:
            h, w = outputs["rgb"].shape[:2]
            outputs["prompt_points"] = P.prompts_in_image(self.prompts, intrin, c2w, w, h)
        self._decode_mask(outputs)

    def attach_mask_decoder(self, predictor) -> None:
        """Give the model SAM's prompt encoder + mask decoder (``mask_decoder.SamMaskPredictor``, or a SAM checkpoint path /
        state dict for ``SamMaskPredictor.from_sam_checkpoint``): the whole-image entry points then turn the remembered
        prompts into ``outputs["masked_rgb"]`` like the reference (sam_model.py:485-486,514-527).  Kept out of the module
        tree, like the reference's ``self.predictor``: the decoder is not part of the NeRF's ``state_dict()``."""
        from .mask_decoder import SamMaskPredictor

        if predictor is not None and not isinstance(predictor, SamMaskPredictor):
            predictor = SamMaskPredictor.from_sam_checkpoint(predictor, device=self.renderer.device)
        self.__dict__["predictor"] = predictor

    def attach_clipseg_decoder(self, decoder, text_encoder=None) -> None:
        """Give the model the ClipSeg decoder behind the rendered ClipSeg map (``mask_decoder.ClipSegDecoder``, or the path /
        state dict of ``rd64-uni.pth``, sam_model.py:216-222).  ``text_prompt`` of the whole-image entry points is then
        either the prompt's CLIP embedding ``[1,512]`` or - with ``text_encoder``, a callable ``str -> [1,512]`` such as
        ``lambda s: clip_model.encode_text(clip.tokenize([s]))`` - the string itself; the CLIP text tower is OpenAI's ``clip``
        package and not part of this repository.  Kept out of the module tree like ``self.clipseg`` weights are kept out of
        the reference's checkpoints."""
        from .mask_decoder import ClipSegDecoder

        if decoder is not None and not isinstance(decoder, ClipSegDecoder):
            decoder = ClipSegDecoder.from_checkpoint(decoder, device=self.renderer.device)
        self.__dict__["clipseg_decoder"] = decoder
        self.__dict__["text_encoder"] = text_encoder

    def _decode_clipseg(self, outputs: Dict[str, torch.Tensor], text_prompt) -> None:
        """sam_model.py:487-512: rendered ClipSeg map + text prompt -> ``outputs["clipseg_feature"]`` (the 512 x 512 heat map)
        and ``outputs["clipseg_points"]`` (click prompts in image pixels, fed to the mask decoder with the user's clicks)."""
        from .mask_decoder import clipseg_heat_and_clicks

        dec = self.__dict__.get("clipseg_decoder")
        if dec is None or "clipseg" not in outputs or text_prompt is None:
            return
        if isinstance(text_prompt, str):
            enc = self.__dict__.get("text_encoder")
            if enc is None:
                raise RuntimeError("a text prompt needs CLIP's text tower: pass the prompt's 512-d CLIP embedding as text_prompt, "
                                   "or attach_clipseg_decoder(decoder, text_encoder=...)")
            text_prompt = enc(text_prompt)
        cond = torch.as_tensor(text_prompt, dtype=torch.float32).reshape(1, -1)
        h, w = outputs["rgb"].shape[:2]
        heat, clicks = clipseg_heat_and_clicks(dec, outputs["clipseg"], cond, w, h)
        outputs["clipseg_feature"] = heat.to(outputs["rgb"].device)
        outputs["clipseg_points"] = torch.from_numpy(clicks)

    def _decode_mask(self, outputs: Dict[str, torch.Tensor]) -> None:
        """``predictor.set_feature(outputs["sam"])`` + ``generate_masked_img(predictor, prompts, 1s, rgb)`` for the prompts
        that fall inside this view (sam_model.py:485-486,514-527).  Without an attached decoder, without a rendered SAM
        map or without a prompt in view ``masked_rgb`` stays the plain rgb.  (The reference then draws the visible prompts
        as red key points with torchvision - ``show_prompts``, sam_model.py:39-92 - which is viewer decoration and not
        reproduced; ``prompts.visible`` gives the same selection.)"""
        from .mask_decoder import generate_masked_img

        pred = self.__dict__.get("predictor")
        if pred is None or "sam" not in outputs:
            return
        pts = [p.cpu().numpy().astype("float32") for p in (outputs.get("prompt_points"), outputs.get("clipseg_points"))
               if p is not None and len(p) > 0]  # the user's clicks first, then ClipSeg's (sam_model.py:513-514)
        if not pts:
            return
        import numpy as np

        pts = np.concatenate(pts, axis=0)
        h, w = outputs["rgb"].shape[:2]
        pred.set_feature(outputs["sam"], (h, w))
        rgb = outputs["rgb"].to(pred.device)
        outputs["masked_rgb"] = generate_masked_img(pred, pts, [1] * len(pts), rgb).to(outputs["rgb"].device)

    @torch.no_grad()
    def get_outputs_for_camera(self, camera: Camera, points=None, fast: bool = False, text_prompt=None) -> Dict[str, torch.Tensor]:
        """``get_outputs_for_camera_ray_bundle(cameras.generate_rays(i, keep_shape=True))`` without the ray bundle
        (SURVEY.md 8 f-2): the three loops of sam_model.py:354-406 each become one ``snrf_render_camera`` call that
        generates its rays on the device - LOOP A every pixel, LOOP B the ``fh*p x fw*p`` strided sub-grid in
        patch-major order through the conv head, LOOP C the 32 x 32 ClipSeg grid.  ``points`` are the viewer's clicks
        (pixel ``[x, y]`` rows, as for ``get_outputs_for_camera_ray_bundle``); the intrinsics and pose the prompt
        bookkeeping needs come from the camera itself.  ``points=None`` leaves remembered prompts alone (the viewer
        interleaves both calls), an empty list clears them as in the reference."""
        cfg, r = self.config, self.renderer
        h, w = camera.height, camera.width
        outputs = {k: v.view(h, w, -1) for k, v in r.render_camera(camera, get_feature=(), fast=fast).items()}
        if cfg.distill_sam:
            fh, fw = get_feature_size(h, w)
            p = cfg.patch_size
            hi = torch.linspace(0, h - 1, fh * p, dtype=torch.long)
            wi = torch.linspace(0, w - 1, fw * p, dtype=torch.long)
            outputs["sam"] = r.render_camera(camera, rows=hi, cols=wi, get_feature=("sam",), patch=p > 1)["sam"].view(fh, fw, -1)
            if cfg.use_clipseg_feature:
                hi = torch.linspace(0, h - 1, 32, dtype=torch.long)
                wi = torch.linspace(0, w - 1, 32, dtype=torch.long)
                outputs["clipseg"] = r.render_camera(camera, rows=hi, cols=wi, get_feature=("clipseg",))["clipseg"].view(32, 32, -1)
        if points is None:  # keep what get_outputs_for_camera_ray_bundle has remembered; re-project it into this view
            points = [[0, 0]] * int(self.prompts.shape[0]) if getattr(self, "prompts", None) is not None else None
        intrin = torch.tensor([[camera.fx, 0.0, camera.cx], [0.0, camera.fy, camera.cy], [0.0, 0.0, 1.0]])
        self._handle_prompts(outputs, points, intrin if points is not None else None,
                             camera.camera_to_world if points is not None else None, text_prompt)
        return outputs
