"""Generate ``tests/golden/*.npz`` by running the REFERENCE'S OWN Python for the hot path.

Runs only in the build container (needs /root/reference, which does not exist on the GPU box):

    python -m oracle.make_golden            # writes tests/golden/*.npz

What executes, unmodified, from /root/reference: ``ProposalNetworkSampler`` / ``PDFSampler`` /
``UniformLinDispPiecewiseSampler`` (model_components/ray_samplers.py), ``RayBundle`` / ``RaySamples`` /
``Frustums`` (cameras/rays.py), ``RGBRenderer`` / ``DepthRenderer`` / ``AccumulationRenderer``
(model_components/renderers.py), ``NearFarCollider`` (scene_colliders.py), ``SceneContraction``
(spatial_distortions.py), ``trunc_exp`` (activations.py), ``HashMLPDensityField`` (fields/density_fields.py),
``TCNNNerfactoField`` (fields/nerfacto_field.py), ``SAMField`` (samnerf/sam_field.py), and the *source text*
of ``MeanRenderer``, ``SAMModel.get_outputs``, ``_get_outputs_nerfacto``, ``forward`` and
``get_outputs_for_camera_ray_bundle`` (samnerf/sam_model.py:126-137,226-314,338-418) compiled straight out
of the file with ``ast`` (the module itself cannot be imported: it needs torchmetrics, the SAM checkpoint and
a GPU, sam_model.py:210-214).  Stubbed third-party modules: ``torchtyping`` (annotations only), ``nerfacc``
(packed-sample branches only, not on this path) and ``tinycudann`` -> ``oracle/fake_tinycudann.py``.

Parameters come from ``samnerf_b200.make_synthetic_params`` (seeded); each fixture stores a checksum of
them so drift of the generator is detected rather than silently compared.
"""
from __future__ import annotations

import ast
import os
import sys
import textwrap
import types
from types import SimpleNamespace

import numpy as np
import torch
from torch import nn

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


# ------------------------------------------------------------------------------------------------
# reference loader
# ------------------------------------------------------------------------------------------------
def _install_stubs():
    from oracle import fake_tinycudann

    class _TT:
        def __class_getitem__(cls, item):
            return cls

    tt = types.ModuleType("torchtyping")
    tt.TensorType = _TT
    sys.modules["torchtyping"] = tt
    na = types.ModuleType("nerfacc")
    na.OccupancyGrid = object
    na.ContractionType = object
    sys.modules["nerfacc"] = na
    sys.modules["tinycudann"] = fake_tinycudann
    if REF not in sys.path:
        sys.path.insert(0, REF)


def _extract(path: str, wanted):
    """Return {qualified name: source} for classes / methods named in ``wanted``."""
    src = open(path).read()
    tree = ast.parse(src)
    out = {}
    for node in tree.body:
        if isinstance(node, ast.ClassDef):
            if node.name in wanted:
                out[node.name] = ast.get_source_segment(src, node)
            for sub in node.body:
                if isinstance(sub, ast.FunctionDef) and f"{node.name}.{sub.name}" in wanted:
                    out[f"{node.name}.{sub.name}"] = ast.get_source_segment(src, sub)
    return out


def load_reference():
    """Import the reference modules and compile the SAMModel method bodies. Returns a namespace."""
    _install_stubs()
    from nerfstudio.cameras.rays import RayBundle, RaySamples  # noqa
    from nerfstudio.field_components.field_heads import FieldHeadNames
    from nerfstudio.field_components.spatial_distortions import SceneContraction
    from nerfstudio.fields.density_fields import HashMLPDensityField
    from nerfstudio.fields.nerfacto_field import TCNNNerfactoField
    from nerfstudio.model_components.ray_samplers import ProposalNetworkSampler
    from nerfstudio.model_components.renderers import AccumulationRenderer, DepthRenderer, RGBRenderer
    from nerfstudio.model_components.scene_colliders import NearFarCollider
    from samnerf.sam_field import SAMField
    from samnerf.sam_utils import get_feature_size

    def get_feature_size_sq(h, w, largesize=64):
        # the reference raises UnboundLocalError for h == w (sam_utils.py:7-14); documented deviation
        return (largesize, largesize) if h == w else get_feature_size(h, w, largesize)

    srcs = _extract(
        os.path.join(REF, "samnerf", "sam_model.py"),
        {
            "MeanRenderer",
            "SAMModel.get_outputs",
            "SAMModel._get_outputs_nerfacto",
            "SAMModel.forward",
            "SAMModel.get_outputs_for_camera_ray_bundle",
        },
    )
    # keep get_outputs_for_camera_ray_bundle up to the start of the prompt / mask post-processing
    body = srcs["SAMModel.get_outputs_for_camera_ray_bundle"]
    cut = body.index("# calculate SAM relevant")
    body = body[:cut].rstrip() + "\n        return outputs\n"
    srcs["SAMModel.get_outputs_for_camera_ray_bundle"] = body

    from collections import defaultdict
    from typing import Dict, List, Optional, Tuple  # noqa

    ns = dict(
        torch=torch, nn=nn, RayBundle=RayBundle, RaySamples=RaySamples, FieldHeadNames=FieldHeadNames,
        defaultdict=defaultdict, get_feature_size=get_feature_size_sq, TensorType=sys.modules["torchtyping"].TensorType,
        Dict=Dict, List=List, Optional=Optional, Tuple=Tuple,
    )
    exec(srcs["MeanRenderer"], ns)
    cls_src = "class RefSAMModel(nn.Module):\n" + "\n\n".join(
        textwrap.indent(textwrap.dedent(srcs[k]), "    ")
        for k in (
            "SAMModel.get_outputs",
            "SAMModel._get_outputs_nerfacto",
            "SAMModel.forward",
            "SAMModel.get_outputs_for_camera_ray_bundle",
        )
    )
    exec(cls_src, ns)
    return SimpleNamespace(
        RayBundle=RayBundle, SceneContraction=SceneContraction, HashMLPDensityField=HashMLPDensityField,
        TCNNNerfactoField=TCNNNerfactoField, ProposalNetworkSampler=ProposalNetworkSampler,
        AccumulationRenderer=AccumulationRenderer, DepthRenderer=DepthRenderer, RGBRenderer=RGBRenderer,
        NearFarCollider=NearFarCollider, SAMField=SAMField, MeanRenderer=ns["MeanRenderer"],
        RefSAMModel=ns["RefSAMModel"],
    )


def build_reference_model(ref, cfg, params):
    """Wire the reference modules exactly as ``NerfactoModel.populate_modules`` (models/nerfacto.py:149-225)
    and ``SAMModel.populate_modules`` (samnerf/sam_model.py:182-208) do, then load ``params``."""
    aabb = torch.tensor([[-1.0, -1.0, -1.0], [1.0, 1.0, 1.0]])
    contraction = ref.SceneContraction(order=float("inf"))
    m = ref.RefSAMModel()
    fg = cfg.field_grid
    m.field = ref.TCNNNerfactoField(
        aabb, hidden_dim=cfg.field_hidden, num_levels=fg.n_levels, max_res=fg.max_resolution,
        log2_hashmap_size=fg.log2_hashmap_size, hidden_dim_color=cfg.head_hidden, spatial_distortion=contraction,
        num_images=2, use_pred_normals=False, use_average_appearance_embedding=True, use_appearance_embedding=False,
    )
    pg = cfg.proposal_grid
    net = ref.HashMLPDensityField(
        aabb, spatial_distortion=contraction, hidden_dim=cfg.proposal_hidden, log2_hashmap_size=pg.log2_hashmap_size,
        num_levels=pg.n_levels, max_res=pg.max_resolution, use_linear=False,
    )
    m.proposal_networks = nn.ModuleList([net])
    m.density_fns = [net.density_fn]
    m.proposal_sampler = ref.ProposalNetworkSampler(
        num_nerf_samples_per_ray=cfg.num_nerf_samples, num_proposal_samples_per_ray=(cfg.num_proposal_samples,),
        num_proposal_network_iterations=1, single_jitter=True,
    )
    m.collider = ref.NearFarCollider(near_plane=0.05, far_plane=cfg.far_plane)
    m.renderer_rgb = ref.RGBRenderer(background_color="last_sample")
    m.renderer_accumulation = ref.AccumulationRenderer()
    m.renderer_depth = ref.DepthRenderer()
    m.renderer_mean = ref.MeanRenderer()
    m.config = SimpleNamespace(
        num_proposal_iterations=1, distill_sam=cfg.distill_sam, num_sam_samples=cfg.num_sam_samples,
        sharpening_temperature=cfg.sharpening_temperature, patch_size=cfg.patch_size, predict_normals=False,
        eval_num_rays_per_chunk=cfg.eval_num_rays_per_chunk, use_dino_feature=False,
        use_clipseg_feature=cfg.use_clipseg_feature,
    )
    if cfg.distill_sam:
        m.sam_field = ref.SAMField(
            tuple(g.n_levels for g in cfg.sam_grids), tuple(g.log2_hashmap_size for g in cfg.sam_grids),
            tuple((g.base_resolution, g.max_resolution) for g in cfg.sam_grids), hidden_layers=1,
            use_dino_features=False, use_clipseg_features=cfg.use_clipseg_feature,
        )
        k = cfg.kernel_size
        m.conv_head = nn.Sequential(
            nn.Conv2d(256, 256, k, stride=1, padding=(k - 1) // 2), nn.ReLU(inplace=True),
            nn.Conv2d(256, 256, k, stride=1, padding=(k - 1) // 2),
        )
    sd = m.state_dict()
    for name, val in params.items():
        assert name in sd, (name, [k for k in sd if "param" in k or "conv" in k])
        assert sd[name].shape == val.shape, (name, sd[name].shape, val.shape)
    missing = [k for k in sd if k not in params and (k.endswith("params") and sd[k].numel() > 0 or "conv_head" in k)]
    assert not missing, missing
    m.load_state_dict(params, strict=False)
    m.eval()
    return m


def make_bundle(ref, origins, directions):
    return ref.RayBundle(
        origins=origins, directions=directions, pixel_area=torch.ones_like(origins[..., :1]),
        camera_indices=torch.zeros_like(origins[..., :1]).long(),
    )


def params_checksum(params) -> float:
    return float(sum(float(v.double().abs().sum()) for v in params.values()))


# ------------------------------------------------------------------------------------------------
# fixtures
# ------------------------------------------------------------------------------------------------
def fixture_specs():
    """name -> (config factory args, regime, seed, ray source, kind). Shared with tests/test_oracle_golden.py."""
    return {
        "chunk_tiny_scene": dict(cfg="tiny", clipseg=True, patch=1, regime="scene", seed=3, rays="plumbing", n=384),
        "chunk_tiny_init": dict(cfg="tiny", clipseg=False, patch=1, regime="init", seed=4, rays="plumbing", n=256),
        "chunk_tiny_patch4": dict(cfg="tiny", clipseg=False, patch=4, regime="scene", seed=5, rays="plumbing", n=256),
        "chunk_full_scene": dict(cfg="full", clipseg=False, patch=1, regime="scene", seed=0, rays="orbit", n=192),
        "image_tiny": dict(cfg="tiny", clipseg=True, patch=4, regime="scene", seed=6, rays="image", n=24 * 32),
        # the shipped full-size configuration on 4 096 rays of the benchmark frame (SURVEY.md 8 d), without / with ClipSeg
        "chunk_full_4k": dict(cfg="full", clipseg=False, patch=1, regime="scene", seed=0, rays="orbit", n=4096),
        "chunk_full_clipseg_4k": dict(cfg="full", clipseg=True, patch=1, regime="scene", seed=2, rays="orbit", n=4096),
    }


def make_cfg(spec):
    from samnerf_b200 import SAMNeRFConfig

    if spec["cfg"] == "tiny":
        return SAMNeRFConfig.tiny(clipseg=spec["clipseg"], patch_size=spec["patch"])
    return SAMNeRFConfig.distill(clipseg=spec["clipseg"], patch_size=spec["patch"])


def make_rays(spec):
    from samnerf_b200.synthetic import look_at, orbit_rays, pinhole_rays, plumbing_rays

    if spec["rays"] == "plumbing":
        o, d = plumbing_rays()
        step = o.shape[0] // spec["n"]
        idx = torch.arange(spec["n"]) * step + (torch.arange(spec["n"]) % 7)
        if spec["patch"] > 1:  # keep whole patches of adjacent rays together
            idx = torch.arange(spec["n"]) + 2048
        return o[idx].contiguous(), d[idx].contiguous()
    if spec["rays"] == "orbit":
        o, d = orbit_rays()
        o, d = o.reshape(-1, 3), d.reshape(-1, 3)
        idx = (torch.arange(spec["n"]) * 3331 + 17) % o.shape[0]
        return o[idx].contiguous(), d[idx].contiguous()
    if spec["rays"] == "image":
        return pinhole_rays(24, 32, 32.0, 32.0, look_at((1.1, 0.6, 0.45)))
    raise ValueError(spec["rays"])


def main():
    from samnerf_b200 import make_synthetic_params

    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count() or 1)
    os.makedirs(GOLDEN, exist_ok=True)
    ref = load_reference()
    only = set(sys.argv[1:])  # python -m oracle.make_golden [fixture names]: regenerate just those
    for name, spec in fixture_specs().items():
        if only and name not in only:
            continue
        cfg = make_cfg(spec)
        params = make_synthetic_params(cfg, spec["regime"], spec["seed"])
        model = build_reference_model(ref, cfg, params)
        o, d = make_rays(spec)
        with torch.no_grad():
            if spec["rays"] == "image":
                out = model.get_outputs_for_camera_ray_bundle(make_bundle(ref, o, d))
            else:
                feats = ["sam", "clipseg"] if cfg.use_clipseg_feature else ["sam"]
                out = model(make_bundle(ref, o, d), get_feature=feats)
        arrays = {k: v.detach().float().numpy() for k, v in out.items() if torch.is_tensor(v)}
        if spec["rays"] == "image":  # keep the fixture small: every 3rd feature pixel
            arrays["sam"] = arrays["sam"][::3, ::3].copy()
            arrays["_sam_stride"] = np.array(3)
        arrays["_origins"] = o.numpy()
        arrays["_directions"] = d.numpy()
        arrays["_params_checksum"] = np.array(params_checksum(params))
        path = os.path.join(GOLDEN, name + ".npz")
        np.savez_compressed(path, **arrays)
        print(name, {k: v.shape for k, v in arrays.items()}, f"{os.path.getsize(path)/1024:.0f} KiB")


if __name__ == "__main__":
    main()
