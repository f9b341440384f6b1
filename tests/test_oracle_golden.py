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


@pytest.mark.parametrize("name", ["chunk_tiny_scene", "chunk_tiny_init", "chunk_tiny_patch4", "chunk_full_scene"])
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
