"""The drop-in claim, exercised: the REFERENCE'S OWN ``SAMModel.get_outputs_for_camera_ray_bundle`` body
(samnerf/sam_model.py:336-418, compiled straight out of the reference file as oracle/make_golden.py does) runs unchanged
with this package's ``SAMModel`` as ``self`` and this package's ``RayBundle`` as its argument - chunk loops A / B / C, the
strided feature-ray grid, ``reshape`` / ``_apply_fn_to_fields`` / ``get_row_major_sliced_ray_bundle`` - and gives the
same image as the shim's own method and as the golden image the reference's modules produced.

Needs /root/reference (build container only; the GPU box skips it) and runs on the CPU through tests/fake_renderer.py."""
import os
import textwrap
import types
from collections import defaultdict

import numpy as np
import pytest
import torch

REF = "/root/reference"
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")

pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "samnerf", "sam_model.py")),
                                reason="the reference tree only exists in the build container")


def _reference_method():
    from oracle.make_golden import _extract

    src = _extract(os.path.join(REF, "samnerf", "sam_model.py"), {"SAMModel.get_outputs_for_camera_ray_bundle"})
    body = src["SAMModel.get_outputs_for_camera_ray_bundle"]
    body = body[: body.index("# calculate SAM relevant")].rstrip() + "\n        return outputs\n"
    from samnerf_b200.config import get_feature_size
    from samnerf_b200.nerfstudio_api import RayBundle

    ns = dict(torch=torch, defaultdict=defaultdict, RayBundle=RayBundle, get_feature_size=get_feature_size)
    exec(textwrap.dedent(body), ns)
    return ns["get_outputs_for_camera_ray_bundle"]


def test_reference_whole_image_body_runs_over_the_shim(monkeypatch):
    import samnerf_b200.nerfstudio_api as api
    from fake_renderer import FakeRenderer
    from oracle.make_golden import fixture_specs, make_cfg
    from samnerf_b200 import make_synthetic_params

    monkeypatch.setattr(api, "Renderer", FakeRenderer)
    spec = fixture_specs()["image_tiny"]
    cfg = make_cfg(spec)
    params = make_synthetic_params(cfg, spec["regime"], spec["seed"])
    z = np.load(os.path.join(GOLDEN, "image_tiny.npz"))
    m = api.SAMModel(cfg)
    assert isinstance(m, torch.nn.Module) and isinstance(m.sam_field, torch.nn.Module)
    m.load_state_dict({"_model." + k: v for k, v in params.items()})
    o, d = torch.from_numpy(z["_origins"]), torch.from_numpy(z["_directions"])
    bundle = api.RayBundle(origins=o, directions=d, pixel_area=torch.ones_like(o[..., :1]),
                           camera_indices=torch.zeros_like(o[..., :1]).long())
    ref_body = types.MethodType(_reference_method(), m)
    got = ref_body(bundle)
    own = m.get_outputs_for_camera_ray_bundle(bundle)
    assert got["rgb"].shape == (24, 32, 3) and got["sam"].shape == (48, 64, 256) and got["clipseg"].shape == (32, 32, 192)
    for k in ("rgb", "depth", "accumulation", "prop_depth_0", "sam", "clipseg"):
        assert torch.equal(torch.nan_to_num(got[k]), torch.nan_to_num(own[k])), k
    # and both are the image the reference's own modules rendered (fp32 CPU on both sides: rounding order only)
    st = int(z["_sam_stride"])
    np.testing.assert_allclose(got["rgb"].numpy(), z["rgb"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(got["sam"].numpy()[::st, ::st], z["sam"], rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(got["clipseg"].numpy(), z["clipseg"], rtol=1e-4, atol=2e-5)


def test_state_dict_is_the_reference_models_key_set(monkeypatch):
    """``nn.Module.state_dict()`` of the shim == the key list and shapes of the reference's own modules
    (tests/golden/state_dict_layout.json), and ``load_state_dict(strict=True)`` accepts exactly that."""
    import json

    import samnerf_b200.nerfstudio_api as api
    from fake_renderer import FakeRenderer
    from samnerf_b200 import SAMNeRFConfig, make_synthetic_params

    monkeypatch.setattr(api, "Renderer", FakeRenderer)
    cfg = SAMNeRFConfig.tiny(clipseg=True, patch_size=4)
    with open(os.path.join(GOLDEN, "state_dict_layout.json")) as f:
        layout = {k[len("_model."):]: v for k, v in json.load(f)["tiny_distill_clipseg_p4"].items() if k.startswith("_model.")}
    m = api.SAMModel(cfg)
    sd = m.state_dict()
    assert set(sd) == set(layout), (sorted(set(sd) - set(layout)), sorted(set(layout) - set(sd)))
    for k, shp in layout.items():
        assert list(sd[k].shape) == list(shp), (k, tuple(sd[k].shape), shp)
    params = make_synthetic_params(cfg, "init", 4)
    full = {k: (params[k].reshape(v.shape) if k in params else v) for k, v in sd.items()}
    res = m.load_state_dict(full, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    with pytest.raises(RuntimeError):  # nn.Module's strict check: an unknown key is refused, as in the reference's trainer
        m.load_state_dict(dict(full, **{"field.bogus.params": torch.zeros(3)}), strict=True)
    assert torch.equal(m.state_dict()["sam_field.sam_net.params"], params["sam_field.sam_net.params"])
