"""the SAM prompt encoder / mask decoder behind the rendered feature map (samnerf_b200/mask_decoder.py)
against the reference's own modules.

* ``tests/golden/sam_decoder.npz`` (oracle/make_decoder_golden.py): a reduced-width instance of the reference's
  ``Sam`` / ``SamPredictor`` - weights, feature maps, click prompts, outputs.  The weights load STRICTLY into this
  package's module tree, and ``set_feature`` + ``predict`` reproduce the reference's logits, IoU predictions and
  low-resolution masks to fp32 rounding (stated tolerance 1e-5 relative to the largest logit).
* ``tests/golden/sam_decoder_layout.json``: the key names and shapes of a real SAM checkpoint's ``prompt_encoder.*`` /
  ``mask_decoder.*`` entries - what ``SamMaskPredictor()`` at default widths has to expose.
* in the build container (needs /root/reference): the same comparison against the reference's modules at SAM's real widths.
"""
import json
import os

import numpy as np
import pytest
import torch

from samnerf_b200.mask_decoder import (MaskDecoder, PointPromptEncoder, SamMaskPredictor, TwoWayTransformer, clipseg_click_points,
                                       generate_masked_img, masked_image)

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
REF = "/root/reference"
TINY = dict(embed=32, grid=8, img=128, mask_in_chans=8, depth=2, mlp_dim=64, heads=4, iou_hidden=32)  # = make_decoder_golden.TINY


def _tiny_predictor():
    t = TINY
    return SamMaskPredictor(
        PointPromptEncoder(t["embed"], (t["grid"], t["grid"]), (t["img"], t["img"]), t["mask_in_chans"]),
        MaskDecoder(t["embed"], TwoWayTransformer(t["depth"], t["embed"], t["heads"], t["mlp_dim"]), 3, 3, t["iou_hidden"]),
        img_size=t["img"]).eval()


def _close(a, b, what):
    a, b = torch.as_tensor(a, dtype=torch.float32), torch.as_tensor(b, dtype=torch.float32)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    scale = float(b.abs().max()) + 1e-6
    err = float((a - b).abs().max())
    assert err <= 1e-5 * scale + 1e-6, f"{what}: max |err| {err:.3e} against scale {scale:.3e}"


def test_reference_weights_load_strictly_and_outputs_match():
    z = np.load(os.path.join(GOLDEN, "sam_decoder.npz"))
    p = _tiny_predictor()
    state = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("w.")}
    p.load_state_dict(state, strict=True)  # every reference key is consumed, none is missing
    for name in ("landscape", "square"):
        feat, size = torch.from_numpy(z[f"{name}.feat"]), tuple(int(v) for v in z[f"{name}.size"])
        pts, lab = z[f"{name}.points"], z[f"{name}.labels"]
        for layout in ("hwc", "chw"):  # this package's layout and the reference call site's (sam_model.py:486)
            p.set_feature(feat if layout == "hwc" else feat.permute(2, 0, 1), size)
            for multi in (False, True):
                tag = f"{name}.{'multi' if multi else 'single'}"
                logits, iou, low = p.predict(pts, lab, multimask_output=multi, return_logits=True)
                _close(low, z[tag + ".low"], tag + ".low")
                _close(iou, z[tag + ".iou"], tag + ".iou")
                _close(logits, z[tag + ".logits"], tag + ".logits")
                masks, _, _ = p.predict(pts, lab, multimask_output=multi)
                want = torch.from_numpy(z[tag + ".logits"]) > 0.0
                assert masks.dtype == torch.bool and float((masks != want).float().mean()) < 1e-4  # sign flips of ~0 logits only


def test_default_widths_expose_the_sam_checkpoint_layout():
    with open(os.path.join(GOLDEN, "sam_decoder_layout.json")) as f:
        layout = json.load(f)
    own = {k: list(v.shape) for k, v in SamMaskPredictor().state_dict().items()}
    assert own == layout
    # from_sam_checkpoint: a full SAM state dict (with image-encoder entries) loads, the encoder entries are ignored
    g = torch.Generator().manual_seed(0)
    state = {k: torch.randn(s, generator=g) for k, s in layout.items()}
    state["image_encoder.pos_embed"] = torch.zeros(1, 64, 64, 1280)
    p = SamMaskPredictor.from_sam_checkpoint(state)
    assert torch.equal(p.mask_decoder.iou_token.weight, state["mask_decoder.iou_token.weight"])
    with pytest.raises(RuntimeError):
        SamMaskPredictor.from_sam_checkpoint({k: v for k, v in state.items() if "iou_token" not in k})


def test_click_points_from_the_clipseg_heat_map():
    from oracle.make_decoder_golden import click_heat

    z = np.load(os.path.join(GOLDEN, "sam_decoder.npz"))
    got = clipseg_click_points(click_heat(), 1297, 840)
    assert got.dtype == np.float32 and np.array_equal(got, z["clicks.points"])
    assert clipseg_click_points(torch.zeros(512, 512), 1297, 840).shape == (0, 2)   # nothing above the threshold
    assert clipseg_click_points(torch.ones(64, 64), 100, 50).shape == (16, 2)       # fewer than k blocks: all of them


def test_masked_image_blend_and_errors():
    img = torch.rand(6, 9, 3, generator=torch.Generator().manual_seed(1))
    mask = torch.zeros(6, 9, dtype=torch.bool)
    mask[2:4, 3:7] = True
    out = masked_image(mask, img)
    col = torch.tensor([30 / 255, 144 / 255, 255 / 255])
    assert torch.equal(out[~mask], img[~mask])                                  # untouched outside the mask
    assert torch.allclose(out[mask], col * 0.6 + img[mask] * 0.4, atol=1e-6)   # sam_utils.py:37-42 with the fixed colour
    p = _tiny_predictor()
    with pytest.raises(RuntimeError):
        p.predict(np.zeros((1, 2)), np.ones(1))                                 # no feature map set
    with pytest.raises(NotImplementedError):
        p.prompt_encoder((torch.zeros(1, 1, 2), torch.ones(1, 1)), boxes=torch.zeros(1, 4))
    p.set_feature(torch.randn(5, 8, TINY["embed"]), (100, 160))
    blended = generate_masked_img(p, np.array([[10.0, 20.0]]), np.array([1]), torch.rand(100, 160, 3))
    assert blended.shape == (100, 160, 3) and torch.isfinite(blended).all()


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "samnerf", "segment_anything", "predictor.py")),
                    reason="the reference tree only exists in the build container")
def test_full_width_against_the_reference_modules():
    """SAM's real widths (256-d, 64 x 64 map, 1024 frame): the reference's randomly initialised modules and this package's,
    same state dict, same rendered-map stand-in for a 840 x 1297 image (get_feature_size -> 42 x 64), same clicks."""
    from oracle.make_decoder_golden import build_reference

    sam, ref = build_reference(embed=256, grid=64, img=1024, mask_in_chans=16, depth=2, mlp_dim=2048, heads=8, iou_hidden=256, seed=3)
    mine = SamMaskPredictor.from_sam_checkpoint({k: v for k, v in sam.state_dict().items()})
    g = torch.Generator().manual_seed(9)
    feat = torch.randn(42, 64, 256, generator=g)
    pts = np.array([[400.0, 300.5], [1200.0, 800.0], [3.0, 4.0]])
    lab = np.array([1, 1, 1])
    ref.set_feature(feat.permute(2, 0, 1), original_image_size=(840, 1297))
    mine.set_feature(feat, (840, 1297))
    for multi in (False, True):
        want = ref.predict(point_coords=pts, point_labels=lab, multimask_output=multi, return_logits=True, return_torch=True)
        got = mine.predict(pts, lab, multimask_output=multi, return_logits=True)
        for a, b, what in zip(got, want, ("logits", "iou", "low")):
            _close(a, b, f"full width, multi={multi}: {what}")


def test_model_turns_prompts_into_a_masked_image(monkeypatch):
    """``SAMModel.attach_mask_decoder`` + clicks: ``get_outputs_for_camera_ray_bundle`` hands the rendered SAM map and the
    prompts in view to the decoder and blends the mask over the rgb (sam_model.py:485-486,514-527); the decoder stays out
    of the model's own ``state_dict()``; without prompts (or without a decoder) ``masked_rgb`` is the plain rgb."""
    import samnerf_b200.nerfstudio_api as api
    from fake_renderer import FakeRenderer
    from helpers import model_pair
    from samnerf_b200.synthetic import look_at, pinhole_rays

    monkeypatch.setattr(api, "Renderer", FakeRenderer)
    cfg, params, _ = model_pair("tiny", "scene", 6, False, 1)
    m = api.SAMModel(cfg)
    m.load_state_dict(params)
    keys_before = set(m.state_dict())
    torch.manual_seed(0)
    m.attach_mask_decoder(SamMaskPredictor().eval())  # SAM's real widths, random weights: the plumbing is what is tested
    assert set(m.state_dict()) == keys_before
    H, W, f = 8, 12, 12.0
    c2w = look_at((1.1, 0.6, 0.45))[:3, :4]
    intrin = torch.tensor([[f, 0.0, W / 2.0], [0.0, f, H / 2.0], [0.0, 0.0, 1.0]])
    o, d = pinhole_rays(H, W, f, f, look_at((1.1, 0.6, 0.45)))
    bundle = api.RayBundle(origins=o, directions=d, pixel_area=torch.ones(H, W, 1), camera_indices=torch.zeros(H, W, 1, dtype=torch.long))
    plain = m.get_outputs_for_camera_ray_bundle(bundle)
    assert plain["masked_rgb"] is plain["rgb"]
    out = m.get_outputs_for_camera_ray_bundle(bundle, points=np.array([[3, 2], [9, 5]]), intrin=intrin, c2w=c2w)
    assert out["masked_rgb"].shape == out["rgb"].shape and torch.isfinite(out["masked_rgb"]).all()
    # the same mask straight from the predictor, blended by hand
    pred = m.__dict__["predictor"]
    pred.set_feature(out["sam"], (H, W))
    masks, _, _ = pred.predict(out["prompt_points"].numpy(), [1] * len(out["prompt_points"]))
    assert torch.allclose(out["masked_rgb"], masked_image(masks[0, 0], out["rgb"]))
    m.attach_mask_decoder(None)
    again = m.get_outputs_for_camera_ray_bundle(bundle, points=np.array([[3, 2], [9, 5]]), intrin=intrin, c2w=c2w)
    assert again["masked_rgb"] is again["rgb"]


@pytest.mark.gpu
def test_gpu_rendered_map_through_the_decoder():
    """The whole downstream chain on the device: camera -> rendered rgb / depth / SAM map (libsnrf) -> clicks lifted and
    re-projected -> prompt encoder + mask decoder on the same GPU -> ``masked_rgb``; the decoder gives the same logits on
    the GPU as on the CPU for the rendered map (library kernels, TF32 convolutions by torch's default: stated tolerance 5e-3 of
    the largest logit)."""
    from helpers import model_pair
    from samnerf_b200.nerfstudio_api import SAMModel
    from samnerf_b200.renderer import Camera
    from samnerf_b200.synthetic import look_at

    cfg, params, _ = model_pair("tiny", "scene", 6, False, 4)
    m = SAMModel(cfg)
    m.load_state_dict(params)
    torch.manual_seed(0)
    cpu_pred = SamMaskPredictor().eval()
    m.attach_mask_decoder({k: v.clone() for k, v in cpu_pred.state_dict().items()})  # the checkpoint route, onto the GPU
    assert m.__dict__["predictor"].device.type == "cuda"
    cam = Camera(32.0, 32.0, 16.0, 12.0, 32, 24, look_at((1.1, 0.6, 0.45))[:3, :4])
    out = m.get_outputs_for_camera(cam, points=[[10, 8], [20, 15]])
    torch.cuda.synchronize()
    assert out["masked_rgb"].shape == out["rgb"].shape and out["masked_rgb"].is_cuda and torch.isfinite(out["masked_rgb"]).all()
    pts = out["prompt_points"].cpu().numpy()
    gpu_pred = m.__dict__["predictor"]
    gpu_pred.set_feature(out["sam"], (24, 32))
    cpu_pred.set_feature(out["sam"].cpu(), (24, 32))
    got = gpu_pred.predict(pts, [1] * len(pts), return_logits=True)
    want = cpu_pred.predict(pts, [1] * len(pts), return_logits=True)
    for a, b, what in zip(got, want, ("logits", "iou", "low")):
        scale = float(b.abs().max()) + 1e-6
        assert float((a.cpu() - b).abs().max()) <= 5e-3 * scale + 1e-5, what  # measured on B200: 7e-4 (TF32 transposed convs)
    masks, _, _ = gpu_pred.predict(pts, [1] * len(pts))
    differs = (out["masked_rgb"] - masked_image(masks[0, 0], out["rgb"])).abs().amax(dim=-1) > 1e-6
    assert float(differs.float().mean()) < 1e-3  # the same mask up to sign flips of ~0 logits between two launches


def test_clipseg_decoder_matches_reference_golden():
    """``ClipSegDecoder`` on the fixture of oracle/make_decoder_golden.py: the reference's ``CLIPDensePredT`` (CLIP towers
    stubbed: the ``inp_feature`` branch never touches them) with seeded weights, a rendered-map stand-in and a text
    embedding, through the call of sam_model.py:487-499.  The weights are regenerated from the seed, loaded by NAME."""
    from oracle.make_decoder_golden import clipseg_inputs, seeded_state
    from samnerf_b200.mask_decoder import ClipSegDecoder, clipseg_heat_and_clicks

    z = np.load(os.path.join(GOLDEN, "sam_decoder.npz"))
    dec = ClipSegDecoder()
    state = seeded_state({k: tuple(v.shape) for k, v in dec.state_dict().items()}, 21)
    dec = ClipSegDecoder.from_checkpoint({**state, "clip_model.visual.proj": torch.zeros(768, 512)})  # tower keys are ignored
    cmap, cond = clipseg_inputs()
    heat, clicks = clipseg_heat_and_clicks(dec, cmap, cond, 1297, 840)
    assert heat.shape == (512, 512, 1)
    want = torch.from_numpy(z["clipseg.logits_every_3rd"]).sigmoid()
    _close(heat[::3, ::3, 0], want, "clipseg heat map")
    assert np.array_equal(clicks, clipseg_click_points(heat[..., 0], 1297, 840)) and clicks.shape[1] == 2
    with pytest.raises(RuntimeError):
        ClipSegDecoder.from_checkpoint({k: v for k, v in state.items() if not k.startswith("film_mul")})


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "samnerf", "clipseg", "models", "clipseg.py")),
                    reason="the reference tree only exists in the build container")
def test_clipseg_decoder_has_the_reference_modules_parameter_tree():
    """Key names and shapes of the reference's ``CLIPDensePredT(version="ViT-B/16", reduce_dim=64)`` minus its CLIP towers
    (what ``rd64-uni.pth`` holds) == ``ClipSegDecoder().state_dict()``, and the seeded fixture weights are the same tensors."""
    from oracle.make_decoder_golden import build_reference_clipseg
    from samnerf_b200.mask_decoder import ClipSegDecoder

    ref = {k: v for k, v in build_reference_clipseg(seed=21).state_dict().items() if not k.startswith(("clip_model.", "model."))}
    mine = ClipSegDecoder().state_dict()
    assert {k: tuple(v.shape) for k, v in ref.items()} == {k: tuple(v.shape) for k, v in mine.items()}


def test_model_text_prompt_through_clipseg_to_the_mask(monkeypatch):
    """sam_model.py:487-527 end to end on the shim: rendered ClipSeg map + text embedding -> ``clipseg_feature`` heat map ->
    ClipSeg click points -> (with the user's clicks) SAM mask decoder -> ``masked_rgb``.  A string prompt needs a text
    encoder; without the decoders nothing changes."""
    import samnerf_b200.nerfstudio_api as api
    from fake_renderer import FakeRenderer
    from helpers import model_pair
    from oracle.make_decoder_golden import seeded_state
    from samnerf_b200.mask_decoder import ClipSegDecoder
    from samnerf_b200.synthetic import look_at, pinhole_rays

    monkeypatch.setattr(api, "Renderer", FakeRenderer)
    cfg, params, _ = model_pair("tiny", "scene", 6, True, 1)
    m = api.SAMModel(cfg)
    m.load_state_dict(params)
    keys_before = set(m.state_dict())
    torch.manual_seed(0)
    m.attach_mask_decoder(SamMaskPredictor().eval())
    dec = ClipSegDecoder()
    state = seeded_state({k: tuple(v.shape) for k, v in dec.state_dict().items()}, 21)
    state["trans_conv.bias"] = state["trans_conv.bias"] + 8.0  # hot everywhere: every block passes the 0.7 threshold
    m.attach_clipseg_decoder(state)
    assert set(m.state_dict()) == keys_before
    H, W, f = 8, 12, 12.0
    o, d = pinhole_rays(H, W, f, f, look_at((1.1, 0.6, 0.45)))
    bundle = api.RayBundle(origins=o, directions=d, pixel_area=torch.ones(H, W, 1), camera_indices=torch.zeros(H, W, 1, dtype=torch.long))
    cond = torch.randn(1, 512, generator=torch.Generator().manual_seed(2))
    plain = m.get_outputs_for_camera_ray_bundle(bundle)
    assert "clipseg_feature" not in plain and plain["masked_rgb"] is plain["rgb"]
    out = m.get_outputs_for_camera_ray_bundle(bundle, text_prompt=cond)
    assert out["clipseg_feature"].shape == (512, 512, 1) and float(out["clipseg_feature"].min()) > 0.7
    assert out["clipseg_points"].shape == (1000, 2)
    assert float(out["clipseg_points"][:, 0].max()) < W and float(out["clipseg_points"][:, 1].max()) < H
    assert out["masked_rgb"] is not out["rgb"] and out["masked_rgb"].shape == out["rgb"].shape
    with pytest.raises(RuntimeError, match="text tower"):
        m.get_outputs_for_camera_ray_bundle(bundle, text_prompt="a chair")
    m.attach_clipseg_decoder(state, text_encoder=lambda s: cond)
    again = m.get_outputs_for_camera_ray_bundle(bundle, text_prompt="a chair")
    assert torch.equal(again["clipseg_feature"], out["clipseg_feature"]) and torch.equal(again["masked_rgb"], out["masked_rgb"])
