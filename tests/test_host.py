"""CPU-side checks: the C-ABI library loads and exports what include/snrf.h declares, host-side geometry, configs."""
import ctypes as C
import os
import re

import pytest
import torch

from samnerf_b200 import GridConfig, SAMNeRFConfig, make_synthetic_params
from samnerf_b200 import _lib as L
from samnerf_b200.config import get_feature_size

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    L.build_library()
    return L.load()


def test_library_exports_every_declared_symbol(lib):
    header = open(os.path.join(ROOT, "include", "snrf.h")).read()
    declared = set(re.findall(r"\b(snrf_[a-z0-9_]+)\s*\(", header))
    assert declared == set(L.SYMBOLS), (declared ^ set(L.SYMBOLS))
    for name in declared:
        assert hasattr(lib, name), name


def test_no_context_without_a_gpu(lib):
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = C.c_void_p()
    assert lib.snrf_ctx_create(0, C.byref(h)) != 0 and not h.value
    from samnerf_b200.renderer import Renderer

    with pytest.raises(RuntimeError, match="no CPU path"):
        Renderer(SAMNeRFConfig.tiny())


@pytest.mark.parametrize(
    "g,entries",
    [(GridConfig(5, 2, 17, 16, 128), 383264), (GridConfig(16, 2, 19, 16, 2048), 6098120),
     (GridConfig(12, 8, 19, 16, 128), 2481152), (GridConfig(12, 8, 19, 128, 512), 6291456)],
)
def test_grid_geometry_matches_c_helper_and_survey(lib, g, entries):
    """Table sizes quoted in SURVEY.md 8 a-4/a-9/a-13, and host (Python) == C geometry."""
    assert g.n_entries == entries
    d = L.GridDesc()
    assert lib.snrf_grid_desc_init(C.byref(d), g.n_levels, g.n_features, g.log2_hashmap_size, g.base_resolution,
                                   g.per_level_scale) == 0
    for i, (scale, res, offset, size, hashed) in enumerate(g.levels()):
        assert (d.lv[i].res, d.lv[i].offset, d.lv[i].size, d.lv[i].hashed) == (res, offset, size, int(hashed))
        assert abs(d.lv[i].scale - scale) <= 2e-7 * max(1.0, abs(scale))


def test_param_counts_match_survey():
    cfg = SAMNeRFConfig.distill(clipseg=True)
    p = make_synthetic_params(cfg, "init", 0)
    assert p["proposal_networks.0.mlp_base.params"].numel() == 767040
    assert p["field.mlp_base.params"].numel() + p["field.mlp_head.params"].numel() == 12206480
    sam = sum(v.numel() for k, v in p.items() if "clip_encs" in k or "sam_net" in k)
    clip = sum(v.numel() for k, v in p.items() if "clipseg" in k)
    conv = sum(v.numel() for k, v in p.items() if "conv_head" in k)
    assert (sam, clip, conv) == (70295552, 70279168, 1180160)


def test_get_feature_size():
    assert get_feature_size(1060, 1600) == (43, 64)
    assert get_feature_size(840, 1297) == (42, 64)
    assert get_feature_size(1600, 1060) == (64, 43)
    assert get_feature_size(800, 800) == (64, 64)  # undefined (UnboundLocalError) in the reference


def test_scene_regime_is_scene_like():
    """The synthetic 'scene' parameters give trained-scene statistics (checked with the oracle on a few rays)."""
    from oracle.samnerf_oracle import Oracle
    from samnerf_b200.synthetic import orbit_rays

    cfg = SAMNeRFConfig.tiny(clipseg=False, patch_size=1)
    orc = Oracle(cfg, make_synthetic_params(cfg, "scene", 0))
    o, d = orbit_rays()
    o, d = o.reshape(-1, 3), d.reshape(-1, 3)
    idx = (torch.arange(512) * 1237 + 5) % o.shape[0]
    out = orc.render_rays(o[idx], d[idx], return_intermediates=True)
    depth = out["depth"][:, 0]
    assert float(torch.quantile(depth, 0.9) / torch.quantile(depth, 0.1)) > 4.0   # rays end at varied depths
    assert float(out["accumulation"].median()) > 0.99
    assert float((out["_sam_weights"] > 1e-3).float().sum(-1).mean()) < 4.0       # sharpened weights concentrate
    assert not torch.isnan(out["sam"]).any()
