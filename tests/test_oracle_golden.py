"""The CPU oracle (oracle/samnerf_oracle.py) against golden vectors produced by the reference's own
Python (oracle/make_golden.py, run in the build container where /root/reference exists)."""
import os

import numpy as np
import pytest
import torch

from oracle.make_golden import fixture_specs, make_cfg, params_checksum
from oracle.samnerf_oracle import Oracle
from samnerf_b200 import make_synthetic_params

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _load(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: z[k] for k in z.files}


def _oracle_for(name):
    spec = fixture_specs()[name]
    cfg = make_cfg(spec)
    params = make_synthetic_params(cfg, spec["regime"], spec["seed"])
    g = _load(name)
    assert abs(params_checksum(params) - float(g["_params_checksum"])) <= 1e-9 * float(g["_params_checksum"]), (
        "synthetic parameter generator drifted from the one that produced the fixture"
    )
    return Oracle(cfg, params), cfg, g


@pytest.mark.parametrize("name", ["chunk_tiny_scene", "chunk_tiny_init", "chunk_tiny_patch4", "chunk_full_scene", "chunk_full_4k",
                                  "chunk_full_clipseg_4k"])
def test_chunk_matches_reference(name):
    """Restated samplers / fields / renderers / top-k / MeanRenderer / conv head == the reference's code.
    Both sides are fp32 CPU torch with the same tcnn stand-in, so the tolerance is rounding-order only."""
    orc, cfg, g = _oracle_for(name)
    feats = ("sam", "clipseg") if cfg.use_clipseg_feature else ("sam",)
    out = orc.render_rays(torch.from_numpy(g["_origins"]), torch.from_numpy(g["_directions"]), get_feature=feats)
    for key in ("rgb", "accumulation", "depth", "prop_depth_0", "sam", "clipseg"):
        if key not in g:
            continue
        a, b = out[key].numpy(), g[key]
        assert a.shape == b.shape, (key, a.shape, b.shape)
        np.testing.assert_allclose(a, b, rtol=1e-5, atol=1e-6, err_msg=f"{name}:{key}")


def test_image_matches_reference():
    """Chunk loops A/B/C, the strided feature-ray grid and its patch-major order (sam_model.py:354-418)."""
    orc, cfg, g = _oracle_for("image_tiny")
    out = orc.render_image(torch.from_numpy(g["_origins"]), torch.from_numpy(g["_directions"]))
    st = int(g["_sam_stride"])
    for key in ("rgb", "accumulation", "depth", "prop_depth_0", "clipseg"):
        np.testing.assert_allclose(out[key].numpy(), g[key], rtol=1e-5, atol=1e-6, err_msg=key)
    np.testing.assert_allclose(out["sam"].numpy()[::st, ::st], g["sam"], rtol=1e-5, atol=2e-6)


def test_training_mode_sampler_matches_reference():
    """Stratified single-jitter sampling (ray_samplers.py:104-112,314-322): the reference's own sampler in training
    mode with its torch.rand draws recorded (oracle/make_jitter_golden.py) vs the oracle fed the same draws."""
    import numpy as np
    import torch

    from oracle.samnerf_oracle import Oracle
    from samnerf_b200 import SAMNeRFConfig, make_synthetic_params

    z = np.load(os.path.join(GOLDEN, "sampler_training.npz"))
    cfg = SAMNeRFConfig.tiny(clipseg=False, patch_size=1)
    orc = Oracle(cfg, make_synthetic_params(cfg, "scene", 8))
    o, d, jit = (torch.from_numpy(z[k]) for k in ("_origins", "_directions", "jitter"))
    res = orc.render_rays(o, d, get_feature=(), return_intermediates=True, jitter=jit)
    np.testing.assert_allclose(res["_eu0"].numpy(), z["edges0"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(res["_w0"].numpy(), z["w0"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(res["_bins1"].numpy(), z["spacing1"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(res["_eu1"].numpy(), z["edges1"], rtol=2e-5, atol=1e-6)
    # annealed proposal weights feed the PDF sampler, the reported weights stay raw (ray_samplers.py:583-593)
    ja = torch.from_numpy(z["jitter_anneal"])
    ra = orc.render_rays(o, d, get_feature=(), return_intermediates=True, jitter=ja, anneal=float(z["anneal"]))
    np.testing.assert_allclose(ra["_w0"].numpy(), z["w0_anneal"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(ra["_eu1"].numpy(), z["edges1_anneal"], rtol=2e-5, atol=1e-6)
    # and the jitter really moved the samples away from the eval-mode positions
    ev = orc.render_rays(o, d, get_feature=(), return_intermediates=True)
    assert float((ev["_eu1"] - res["_eu1"]).abs().max()) > 1e-3


def test_boundary_inputs_match_reference():
    """Rays that arrive with nears / fars (collider bypass, scene_colliders.py:40-44), a background colour override
    (renderers.py:46-55,100-101) and ``fast`` mode (sam_model.py:284-299), through the reference's own model
    (oracle/make_boundary_golden.py)."""
    from samnerf_b200 import SAMNeRFConfig

    z = np.load(os.path.join(GOLDEN, "chunk_tiny_boundary.npz"))
    cfg = SAMNeRFConfig.tiny(clipseg=False, patch_size=1)
    orc = Oracle(cfg, make_synthetic_params(cfg, "scene", 9))
    o, d, nr, fr = (torch.from_numpy(z[k]) for k in ("_origins", "_directions", "_nears", "_fars"))
    bg = tuple(float(v) for v in z["_bg"])
    for fast in (False, True):
        out = orc.render_rays(o, d, nr, fr, get_feature=("sam",), fast=fast, background=bg)
        pre = "fast." if fast else ""
        keys = [k[len(pre):] for k in z.files if not k.startswith("_") and (k.startswith("fast.") == fast)]
        assert set(keys) == set(out), (keys, sorted(out))
        for k in keys:
            np.testing.assert_allclose(out[k].numpy(), z[pre + k], rtol=1e-5, atol=1e-6, err_msg=pre + k)
    assert float(torch.from_numpy(z["depth"]).max()) <= float(fr.max())
