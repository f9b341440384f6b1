"""On-disk formats either side of the hot path: the reference's training checkpoints and
its ground-truth feature files.  Host logic only - nothing here computes on the path.

Checkpoint layout restated from ``nerfstudio/engine/trainer.py:379-400`` (``torch.save`` of
``{"step", "pipeline", "optimizers", "scalers"}``; ``pipeline`` is ``Pipeline.state_dict()``), the prefixes from
``nerfstudio/pipelines/base_pipeline.py:109-115`` (``_model.``) and ``:373`` (``module.`` under DDP), the
"latest step" rule from ``trainer.py:357-365`` / ``nerfstudio/utils/eval_utils.py:45-59`` and the parameter names
from the modules themselves (``models/nerfacto.py:149-215``: ``field`` / ``proposal_networks``;
``samnerf/sam_model.py:193-208``: ``sam_field`` / ``conv_head``; tcnn modules expose one flat fp32 ``params``).

Feature files restated from ``samnerf/data/feature_loader.py:13-53`` (``sam_features/<image>.npy`` holding
``[256, h, w]``; ``clipseg_features/<image>.pt`` holding ``{"activations": [...]}``, reduced by
``samnerf/datamanager.py:87-95``) and ``samnerf/preprocessing/get_image_embeddings.py:23-35,60``.
"""
from __future__ import annotations

import os
import re
from typing import Callable, Dict, Mapping, Optional, Sequence, Tuple

import numpy as np
import torch

from .config import GridConfig, SAMNeRFConfig

#: tensors of a reference ``state_dict`` that lie on the hot path (everything else - camera optimiser, appearance
#: embeddings, LPIPS / SAM / ClipSeg 2-D networks, buffers - is ignored by the renderer)
HOT_PATH_KEYS = (
    "proposal_networks.0.mlp_base.params",
    "field.mlp_base.params",
    "field.mlp_head.params",
    "sam_field.clip_encs.0.params",
    "sam_field.clip_encs.1.params",
    "sam_field.sam_net.params",
    "sam_field.clipseg_encs.0.params",
    "sam_field.clipseg_encs.1.params",
    "sam_field.clipseg_net.params",
    "conv_head.0.weight",
    "conv_head.0.bias",
    "conv_head.2.weight",
    "conv_head.2.bias",
)
_REQUIRED = HOT_PATH_KEYS[:3]
_CKPT_RE = re.compile(r"^step-(\d+)\.ckpt$")


def strip_prefixes(key: str) -> str:
    """``module.`` (DDP, base_pipeline.py:373) and ``_model.`` (base_pipeline.py:110-112) in the order the reference
    strips them."""
    key = key.replace("module.", "")
    if key.startswith("_model."):
        key = key[len("_model."):]
    return key


def params_from_state_dict(state: Mapping[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Hot-path parameters out of a pipeline (or bare model) ``state_dict``; raises ``KeyError`` naming what is
    missing when the nerfacto fields are not all there."""
    out: Dict[str, torch.Tensor] = {}
    for key, val in state.items():
        k = strip_prefixes(key)
        if k in HOT_PATH_KEYS:
            out[k] = val.detach().to(torch.float32).contiguous()
    missing = [k for k in _REQUIRED if k not in out]
    if missing:
        raise KeyError(f"checkpoint lacks hot-path parameters {missing}")
    return out


def geometry_from_state_dict(state: Mapping[str, torch.Tensor], base: Optional[SAMNeRFConfig] = None) -> SAMNeRFConfig:
    """The nerfacto / proposal fields register their grid geometry as buffers (``num_levels``, ``max_res``,
    ``log2_hashmap_size``: nerfstudio/fields/nerfacto_field.py:122-124, density_fields.py:66-68), so a checkpoint
    carries it; fold whatever is present into ``base`` (``base_res`` is not stored: 16 in both fields)."""
    import dataclasses

    base = base or SAMNeRFConfig()
    st = {strip_prefixes(k): v for k, v in state.items()}

    def grid(prefix: str, g: GridConfig) -> GridConfig:
        def scalar(name, default):
            v = st.get(f"{prefix}.{name}")
            return int(v.item()) if v is not None and v.numel() == 1 else default

        return GridConfig(scalar("num_levels", g.n_levels), g.n_features, scalar("log2_hashmap_size", g.log2_hashmap_size),
                          g.base_resolution, scalar("max_res", g.max_resolution))

    return dataclasses.replace(base, field_grid=grid("field", base.field_grid),
                               proposal_grid=grid("proposal_networks.0", base.proposal_grid))


def _solve_log2_hashmap(n_grid_params: int, template: GridConfig) -> GridConfig:
    for log2_t in range(4, 31):
        g = GridConfig(template.n_levels, template.n_features, log2_t, template.base_resolution, template.max_resolution)
        if g.n_params == n_grid_params:
            return g
        if g.n_params > n_grid_params and all(not lv[4] for lv in g.levels()):
            break  # every level dense already: a larger table changes nothing
    raise ValueError(
        f"{n_grid_params} grid parameters do not fit a {template.n_levels}-level x {template.n_features}-feature "
        f"hash grid {template.base_resolution}..{template.max_resolution} for any log2_hashmap_size"
    )


def infer_config(params: Mapping[str, torch.Tensor], base: Optional[SAMNeRFConfig] = None) -> SAMNeRFConfig:
    """The renderer's configuration implied by the tensor sizes: hash-table sizes are solved for (the SAM grids'
    are not stored anywhere in a checkpoint; level counts and resolutions come from ``base`` - the shipped configs
    ``samconfigs.py`` / ``nerfacto.py:69-137`` or ``geometry_from_state_dict``), the feature heads present decide ``distill_sam`` / ``use_clipseg_feature`` and the conv head decides
    ``patch_size`` (``samconfigs.py:82,136``: 4 with the conv head, 1 without)."""
    import dataclasses

    base = base or SAMNeRFConfig()
    n = params["proposal_networks.0.mlp_base.params"].numel() - base.proposal_mlp_params
    prop = _solve_log2_hashmap(n, base.proposal_grid)
    n = params["field.mlp_base.params"].numel() - base.field_mlp_params
    fld = _solve_log2_hashmap(n, base.field_grid)
    if params["field.mlp_head.params"].numel() != base.head_mlp_params:
        raise ValueError(
            f"field.mlp_head.params has {params['field.mlp_head.params'].numel()} entries, expected "
            f"{base.head_mlp_params}: appearance embeddings are not supported (samconfigs.py:80,134 turn them off)"
        )
    distill = "sam_field.sam_net.params" in params
    clipseg = "sam_field.clipseg_net.params" in params
    sam_grids = base.sam_grids
    if distill:
        if params["sam_field.sam_net.params"].numel() != base.sam_mlp_params:
            raise ValueError("sam_field.sam_net.params: only hidden_layers = 1 (samconfigs.py:81,135) is supported")
        sam_grids = tuple(
            _solve_log2_hashmap(params[f"sam_field.clip_encs.{i}.params"].numel(), g)
            for i, g in enumerate(base.sam_grids)
        )
        if clipseg:
            for i, g in enumerate(sam_grids):
                if params[f"sam_field.clipseg_encs.{i}.params"].numel() != g.n_params:
                    raise ValueError(f"sam_field.clipseg_encs.{i}.params does not match the SAM grid geometry")
    has_conv = "conv_head.0.weight" in params
    if has_conv:
        k = int(params["conv_head.0.weight"].shape[-1])
        if tuple(params["conv_head.0.weight"].shape) != (256, 256, k, k) or k != base.kernel_size:
            raise ValueError(f"conv_head.0.weight {tuple(params['conv_head.0.weight'].shape)} unsupported")
    return dataclasses.replace(
        base, proposal_grid=prop, field_grid=fld, sam_grids=sam_grids, distill_sam=distill,
        use_clipseg_feature=clipseg and distill, patch_size=base.patch_size if has_conv else 1,
        num_sam_samples=base.num_sam_samples if distill else 3,
    )


def latest_checkpoint(load_dir: str) -> str:
    """``step-{n:09d}.ckpt`` with the largest ``n`` (trainer.py:362, eval_utils.py:56)."""
    steps = sorted(int(m.group(1)) for m in (_CKPT_RE.match(f) for f in os.listdir(load_dir)) if m)
    if not steps:
        raise FileNotFoundError(f"no step-*.ckpt under {load_dir}")
    return os.path.join(load_dir, f"step-{steps[-1]:09d}.ckpt")


def load_checkpoint(path: str, base: Optional[SAMNeRFConfig] = None) -> Tuple[SAMNeRFConfig, Dict[str, torch.Tensor], int]:
    """Read a reference checkpoint (file, or directory -> its latest step).  Returns ``(config, params, step)``;
    ``params`` goes straight into ``Renderer.load_params`` / ``SAMModel.load_state_dict``."""
    if os.path.isdir(path):
        path = latest_checkpoint(path)
    # weights_only: a checkpoint from an untrusted source must not unpickle arbitrary objects; the optimiser /
    # scaler dicts the trainer stores are plain tensors and numbers, which weights_only accepts
    loaded = torch.load(path, map_location="cpu", weights_only=True)
    if "pipeline" not in loaded:
        raise KeyError(f"{path}: not a trainer checkpoint (keys {sorted(loaded)[:5]})")
    params = params_from_state_dict(loaded["pipeline"])
    base = geometry_from_state_dict(loaded["pipeline"], base)
    return infer_config(params, base), params, int(loaded.get("step", 0))


def model_state_dict(cfg: SAMNeRFConfig, params: Mapping[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Every key the reference's ``SAMModel.state_dict()`` carries for this path (tests/golden/state_dict_layout.json,
    produced by the reference's own modules): the hot-path tensors of ``params`` plus the geometry buffers the fields
    register (``aabb``, ``max_res``, ``num_levels``, ``log2_hashmap_size``: nerfacto_field.py:121-125,
    density_fields.py:66-71) and the parameter-free tcnn encodings (``direction_encoding.params`` /
    ``position_encoding.params``, empty tensors), so that ``load_state_dict(strict=True)`` on the reference model finds
    all of them."""
    sd = {k: v.detach().cpu() for k, v in params.items()}
    aabb = torch.tensor([[-1.0, -1.0, -1.0], [1.0, 1.0, 1.0]])  # scene box of the nerfstudio dataparsers (scene_box.py)
    for prefix, g in (("field", cfg.field_grid), ("proposal_networks.0", cfg.proposal_grid)):
        sd.setdefault(f"{prefix}.aabb", aabb.clone())
        sd[f"{prefix}.max_res"] = torch.tensor(g.max_resolution)
        sd[f"{prefix}.num_levels"] = torch.tensor(g.n_levels)
        sd[f"{prefix}.log2_hashmap_size"] = torch.tensor(g.log2_hashmap_size)
    sd.setdefault("field.direction_encoding.params", torch.empty(0))
    sd.setdefault("field.position_encoding.params", torch.empty(0))
    return sd


def save_checkpoint(path: str, params: Mapping[str, torch.Tensor], step: int,
                    source: Optional[str] = None, cfg: Optional[SAMNeRFConfig] = None,
                    extra_pipeline_state: Optional[Mapping[str, torch.Tensor]] = None, ddp: bool = False) -> str:
    """Write ``params`` in the trainer's container layout (trainer.py:389-400).  ``path`` may be a directory
    (-> ``step-{step:09d}.ckpt`` inside it).  Two modes:

    * ``source`` = the reference checkpoint the run started from (file or directory): its WHOLE pipeline state dict,
      optimizer states and grad-scaler state are kept and only the hot-path tensors are replaced (under the prefix
      convention the source uses).  This is what ``Trainer._load_checkpoint`` needs to resume: it calls
      ``load_pipeline(..., strict=True)`` and ``grad_scaler.load_state_dict`` (trainer.py:372-380), both of which raise
      on missing entries.
    * no ``source``: an EVAL-ONLY export.  With ``cfg`` the model part is complete (``model_state_dict``: strict
      loading of the model succeeds), but tensors that live outside the model (camera optimizer, optimizer / scaler
      state) cannot be invented, so ``optimizers`` / ``scalers`` are empty and a training resume from it is refused by
      the reference.  ``eval_load_checkpoint`` (eval_utils.py:36-65, non-strict consumers such as the viewer and
      ``ns-render``) reads it as is."""
    if os.path.isdir(path) or path.endswith(os.sep):
        os.makedirs(path, exist_ok=True)
        path = os.path.join(path, f"step-{step:09d}.ckpt")
    if source is not None:
        if os.path.isdir(source):
            source = latest_checkpoint(source)
        loaded = torch.load(source, map_location="cpu", weights_only=True)
        if "pipeline" not in loaded:
            raise KeyError(f"{source}: not a trainer checkpoint (keys {sorted(loaded)[:5]})")
        pipe = dict(loaded["pipeline"])
        by_bare = {strip_prefixes(k): k for k in pipe}
        for k, v in params.items():
            if k not in by_bare:
                raise KeyError(f"{k} is not in the source checkpoint's pipeline state")
            old = pipe[by_bare[k]]
            if tuple(old.shape) != tuple(v.shape) and old.numel() != v.numel():
                raise ValueError(f"{k}: {tuple(v.shape)} does not fit the source checkpoint's {tuple(old.shape)}")
            pipe[by_bare[k]] = v.detach().cpu().reshape(old.shape).to(old.dtype)
        for k, v in (extra_pipeline_state or {}).items():
            pipe[k] = v
        out = dict(loaded)
        out.update(step=int(step), pipeline=pipe)
        torch.save(out, path)
        return path
    prefix = ("module." if ddp else "") + "_model."
    model = model_state_dict(cfg, params) if cfg is not None else {k: v.detach().cpu() for k, v in params.items()}
    pipe = {prefix + k: v for k, v in model.items()}
    for k, v in (extra_pipeline_state or {}).items():
        pipe[k] = v
    torch.save({"step": int(step), "pipeline": pipe, "optimizers": {}, "scalers": {}}, path)
    return path


# ---------------------------------------------------------------------------------------------------------
# ground-truth feature files
# ---------------------------------------------------------------------------------------------------------
def feature_filenames(image_filenames: Sequence[str], folder: str, ext: str) -> list:
    """``<scene>/<folder>/<image stem><ext>`` beside ``<scene>/images/<image>`` (datamanager.py:45-52,75-82)."""
    return [
        os.path.join(os.path.dirname(os.path.dirname(str(name))), folder, os.path.basename(str(name)).split(".")[0] + ext)
        for name in image_filenames
    ]


def clipseg_activations_to_map(saved: Mapping) -> torch.Tensor:
    """datamanager.py:92-94: concatenated ClipSeg activations minus the CLS token, as a ``[32, 32, C]`` map."""
    return torch.cat(list(saved["activations"]), dim=-1).squeeze()[1:, ...].reshape(512 // 16, 512 // 16, -1)


def crop_sam_embedding(feature: torch.Tensor, h: int, w: int) -> torch.Tensor:
    """get_image_embeddings.py:23-35: SAM pads the short side of its 64x64 embedding; keep the valid rows/columns."""
    size = feature.shape[-1]
    if h < w:
        return feature[..., : int(np.ceil((h / w) * size)), :]
    if h > w:
        return feature[..., :, : int(np.ceil((w / h) * size))]
    return feature


class FeatureDataloader:
    """``samnerf/data/feature_loader.py:13-53``: all views' feature maps in one ``[n, h, w, c]`` tensor, looked up
    at ``(image index, row, column)`` pixel coordinates with the floor of the scaled position."""

    def __init__(self, device, npy_paths: Sequence[str], image_shape: Tuple[int, int], patch_size: int = 1,
                 get_feature: Callable = lambda x: x):
        self.device, self.npy_path, self.image_shape, self.patch_size = device, list(npy_paths), tuple(image_shape), patch_size
        if self.npy_path[0].endswith(".npy"):
            maps = [np.transpose(np.load(p), (1, 2, 0)) for p in self.npy_path]  # c h w -> h w c
            self.features = torch.from_numpy(np.stack(maps, axis=0)).to(device)
        else:
            assert self.npy_path[0].endswith(".pt"), self.npy_path[0]
            self.features = torch.stack([get_feature(torch.load(p, weights_only=True)) for p in self.npy_path], dim=0).to(device)

    def __call__(self, img_points: torch.Tensor) -> torch.Tensor:
        scale = (self.features.shape[1] / self.image_shape[0], self.features.shape[2] / self.image_shape[1])
        x_ind, y_ind = (img_points[:, 1] * scale[0]).long(), (img_points[:, 2] * scale[1]).long()
        return self.features[img_points[:, 0].long(), x_ind, y_ind].to(self.device)
