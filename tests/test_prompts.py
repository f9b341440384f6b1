"""Prompt lifting / projection and feature-map packing (host logic) against the reference's own
``project`` and ``SamPredictor.set_feature`` (tests/golden/prompts.npz, oracle/make_prompt_golden.py)."""
import os

import numpy as np
import torch

from samnerf_b200 import prompts as P

Z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "prompts.npz"))


def test_project_matches_reference():
    got = P.project(torch.from_numpy(Z["intrin"]), torch.from_numpy(Z["c2w"]), torch.from_numpy(Z["points"]))
    assert got.dtype == torch.int32 and np.array_equal(got.numpy(), Z["project"])


def test_feature_map_padding_matches_reference():
    for name in ("landscape", "square"):
        feat = torch.from_numpy(Z[f"{name}.feat"])
        assert np.array_equal(P.pad_feature_map(feat).numpy(), Z[f"{name}.padded"])
        assert P.predictor_input_size(tuple(Z[f"{name}.original"].tolist())) == tuple(Z[f"{name}.input_size"].tolist())
    # portrait: the reference raises (wrong concat axis, predictor.py:123-124); here the width is padded
    tall = P.pad_feature_map(torch.ones(64, 43, 4))
    assert tall.shape == (1, 4, 64, 64) and float(tall[..., 43:].abs().sum()) == 0 and float(tall[..., :43].min()) == 1
    assert P.predictor_input_size((1600, 1060)) == (1024, 679)


def test_lift_then_project_round_trip_and_visibility():
    """A click lifted to 3-D at the rendered depth projects back onto the same pixel in the same view, is visible
    there, and is occluded once the depth map says the surface is nearer."""
    intrin, c2w = torch.from_numpy(Z["intrin"]), torch.from_numpy(Z["c2w"])
    g = torch.Generator().manual_seed(1)
    px = torch.stack([torch.randint(2, 318, (40,), generator=g), torch.randint(2, 238, (40,), generator=g)], -1)
    depth = torch.rand(240, 320, 1, generator=g) * 3.0 + 0.5
    p3 = P.lift_points(px, depth, intrin, c2w)
    back = P.project(intrin, c2w, p3)
    assert int((back - px.to(torch.int32)).abs().max()) <= 1  # int32 truncation of x.999...
    inside = P.prompts_in_image(p3, intrin, c2w, 320, 240)
    assert inside.shape[0] == 40
    assert P.visible(px, p3, depth, intrin, c2w).all()
    assert not P.visible(px, p3, depth - 0.5, intrin, c2w).any()
    # prompts behind the image border are dropped
    far = p3 + torch.tensor([0.0, 0.0, 50.0])
    assert P.prompts_in_image(far, intrin, c2w, 320, 240).shape[0] < 40


def test_clipseg_activations_layout():
    m = torch.arange(32 * 32 * 192, dtype=torch.float32).reshape(32, 32, 192)
    acts = P.clipseg_activations(m)
    assert len(acts) == 3 and all(a.shape == (1025, 1, 64) for a in acts)
    assert torch.equal(acts[1][1:, 0], m[..., 64:128].reshape(-1, 64))
    assert torch.allclose(acts[2][0, 0], m[..., 128:].reshape(-1, 64).mean(0))


def test_model_prompt_bookkeeping(monkeypatch):
    """get_outputs_for_camera_ray_bundle(points, intrin, c2w): clicks are lifted once, remembered, re-projected into
    every later view, and dropped again when the viewer sends an empty list (sam_model.py:426-475)."""
    import samnerf_b200.nerfstudio_api as api
    from fake_renderer import FakeRenderer
    from helpers import model_pair
    from samnerf_b200.synthetic import look_at, pinhole_rays

    monkeypatch.setattr(api, "Renderer", FakeRenderer)
    cfg, params, _ = model_pair("tiny", "scene", 6, False, 1)  # patch 1: the 43 x 64 feature grid stays small for the CPU
    m = api.SAMModel(cfg)
    m.load_state_dict(params)
    H, W, f = 8, 12, 12.0
    c2w = look_at((1.1, 0.6, 0.45))[:3, :4]
    intrin = torch.tensor([[f, 0.0, W / 2.0], [0.0, f, H / 2.0], [0.0, 0.0, 1.0]])
    o, d = pinhole_rays(H, W, f, f, look_at((1.1, 0.6, 0.45)))
    bundle = api.RayBundle(origins=o, directions=d, pixel_area=torch.ones(H, W, 1), camera_indices=torch.zeros(H, W, 1, dtype=torch.long))
    out = m.get_outputs_for_camera_ray_bundle(bundle)
    assert "prompt_points" not in out and out["masked_rgb"] is out["rgb"] and out["sam_embedding"].shape == (1, 256, 64, 64)
    clicks = np.array([[3, 2], [9, 5]])
    out = m.get_outputs_for_camera_ray_bundle(bundle, points=clicks, intrin=intrin, c2w=c2w)
    assert m.prompts.shape == (2, 3)
    assert int((out["prompt_points"] - torch.from_numpy(clicks).to(torch.int32)).abs().max()) <= 1
    first = m.prompts.clone()
    out = m.get_outputs_for_camera_ray_bundle(bundle, points=np.concatenate([clicks, [[6, 6]]]), intrin=intrin, c2w=c2w)
    assert m.prompts.shape == (3, 3) and torch.equal(m.prompts[:2], first)   # old clicks are not lifted again
    m.get_outputs_for_camera_ray_bundle(bundle, points=np.zeros((0, 2)), intrin=intrin, c2w=c2w)
    assert m.prompts is None
