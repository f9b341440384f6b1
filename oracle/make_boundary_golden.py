"""Generate ``tests/golden/chunk_tiny_boundary.npz``: the "inputs that look unused but are part of the contract"
(SURVEY.md 8 b) through the REFERENCE'S OWN code - rays that arrive with nears / fars already set (viewer crop box,
scene_colliders.py:40-44 bypasses the collider) rendered under ``background_color_override_context`` (renderers.py:46-55,
100-101; viewer and scripts/render.py use it) and in ``fast`` mode (sam_model.py:284-299).

TEST INFRASTRUCTURE ONLY.  Runs only in the build container (needs /root/reference):

    python -m oracle.make_boundary_golden
"""
from __future__ import annotations

import os

import numpy as np
import torch

from oracle.make_golden import GOLDEN, build_reference_model, load_reference, params_checksum


def main():
    from samnerf_b200 import SAMNeRFConfig, make_synthetic_params
    from samnerf_b200.synthetic import plumbing_rays

    ref = load_reference()
    from nerfstudio.model_components.renderers import background_color_override_context

    cfg = SAMNeRFConfig.tiny(clipseg=False, patch_size=1)
    params = make_synthetic_params(cfg, "scene", 9)
    model = build_reference_model(ref, cfg, params)
    o, d = plumbing_rays()
    idx = torch.arange(200) * 40 + 11
    o, d = o[idx].contiguous(), d[idx].contiguous()
    g = torch.Generator().manual_seed(1)
    nears = torch.rand(200, 1, generator=g) * 0.3 + 0.02
    fars = nears + torch.rand(200, 1, generator=g) * 4.0 + 0.5
    bg = torch.tensor([0.9, 0.2, 0.4])
    out = {"_origins": o.numpy(), "_directions": d.numpy(), "_nears": nears.numpy(), "_fars": fars.numpy(), "_bg": bg.numpy(),
           "_params_checksum": np.array(params_checksum(params))}
    for fast in (False, True):
        bundle = ref.RayBundle(origins=o, directions=d, pixel_area=torch.ones_like(o[..., :1]),
                               camera_indices=torch.zeros_like(o[..., :1]).long(), nears=nears.clone(), fars=fars.clone())
        with torch.no_grad(), background_color_override_context(bg):
            res = model(bundle, get_feature=["sam"], fast=fast)
        assert torch.equal(bundle.nears, nears)  # the collider left them alone
        for k, v in res.items():
            if torch.is_tensor(v):
                out[("fast." if fast else "") + k] = v.numpy()
    path = os.path.join(GOLDEN, "chunk_tiny_boundary.npz")
    np.savez_compressed(path, **out)
    print(sorted(out), os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
