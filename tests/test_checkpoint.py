"""Checkpoint and feature-file formats against the key layout of the reference's own modules.

``tests/golden/state_dict_layout.json`` is written by ``oracle/make_ckpt_golden.py`` from the ``state_dict()`` of the
reference's ``TCNNNerfactoField`` / ``HashMLPDensityField`` / ``SAMField`` / conv head; the values are seeded here."""
import json
import os

import numpy as np
import pytest
import torch

from samnerf_b200 import SAMNeRFConfig, make_synthetic_params
from samnerf_b200 import checkpoint as ck

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _layout(name):
    with open(os.path.join(GOLDEN, "state_dict_layout.json")) as f:
        return json.load(f)[name]


def _reference_like_state(layout, params, geometry):
    """A pipeline state_dict with every key the reference writes: hot-path tensors from ``params``, the geometry
    buffers from ``geometry`` (cfg), everything else zeros of the recorded shape."""
    state = {}
    for key, shape in layout.items():
        k = ck.strip_prefixes(key)
        if k in params:
            assert list(params[k].shape) == shape, (k, params[k].shape, shape)
            state[key] = params[k].clone()
        elif k.endswith(("max_res", "num_levels", "log2_hashmap_size")):
            g = geometry.field_grid if k.startswith("field.") else geometry.proposal_grid
            state[key] = torch.tensor({"max_res": g.max_resolution, "num_levels": g.n_levels,
                                       "log2_hashmap_size": g.log2_hashmap_size}[k.rsplit(".", 1)[1]])
        else:
            state[key] = torch.zeros(shape)
    return state


@pytest.mark.parametrize("ddp", [False, True])
def test_reference_checkpoint_round_trip(tmp_path, ddp):
    cfg = SAMNeRFConfig.tiny(clipseg=True, patch_size=4)
    params = make_synthetic_params(cfg, "scene", 2)
    state = _reference_like_state(_layout("tiny_distill_clipseg_p4"), params, cfg)
    if ddp:
        state = {"module." + k: v for k, v in state.items()}
    d = tmp_path / "nerfstudio_models"
    d.mkdir()
    for step in (999, 29999, 2000):  # the loader must pick the numerically largest step, like trainer.py:362
        torch.save({"step": step, "pipeline": state if step == 29999 else {}, "optimizers": {}, "scalers": {}},
                   d / f"step-{step:09d}.ckpt")
    cfg2, params2, step = ck.load_checkpoint(str(d))
    assert step == 29999
    assert cfg2 == cfg, (cfg2, cfg)
    assert set(params2) == set(params)
    for k in params:
        assert torch.equal(params2[k], params[k]), k


def test_full_layout_sizes_match_the_survey():
    """The reference's own modules size their tensors as SURVEY 8 a-4/a-9/a-13 computed; infer_config recovers the
    shipped ``samnerf_distill`` geometry (hash sizes 2^17 / 2^19 / 2^19, 2^19) from sizes alone."""
    layout = _layout("full_distill_p4")
    cfg = SAMNeRFConfig.distill(clipseg=False, patch_size=4)
    sizes = {ck.strip_prefixes(k): int(np.prod(v)) for k, v in layout.items()}
    assert sizes["proposal_networks.0.mlp_base.params"] == cfg.proposal_mlp_params + 383264 * 2
    assert sizes["field.mlp_base.params"] == cfg.field_mlp_params + 6098120 * 2
    assert sizes["sam_field.clip_encs.0.params"] == 2481152 * 8 and sizes["sam_field.clip_encs.1.params"] == 6291456 * 8
    meta = {k: torch.empty(n, device="meta") for k, n in sizes.items() if k in ck.HOT_PATH_KEYS and "conv" not in k}
    meta["conv_head.0.weight"] = torch.empty(256, 256, 3, 3, device="meta")
    got = ck.infer_config(meta)
    assert got == cfg


def test_rgb_only_checkpoint_and_errors(tmp_path):
    cfg = SAMNeRFConfig.tiny(clipseg=False, patch_size=1)
    params = {k: v for k, v in make_synthetic_params(cfg, "init", 3).items() if not k.startswith(("sam_field", "conv"))}
    path = ck.save_checkpoint(str(tmp_path) + os.sep, params, step=7)
    assert os.path.basename(path) == "step-000000007.ckpt"
    cfg2, params2, step = ck.load_checkpoint(path, base=SAMNeRFConfig.tiny())
    assert step == 7 and not cfg2.distill_sam and cfg2.patch_size == 1 and cfg2.num_sam_samples == 3
    assert cfg2.proposal_grid == cfg.proposal_grid and cfg2.field_grid == cfg.field_grid
    with pytest.raises(KeyError, match="field.mlp_head"):
        ck.params_from_state_dict({k: v for k, v in params.items() if "mlp_head" not in k})
    bad = dict(params)
    bad["field.mlp_head.params"] = torch.zeros(64 * 64 + 64 * 64 + 16 * 64)  # 64-wide input = appearance embedding on
    with pytest.raises(ValueError, match="appearance"):
        ck.infer_config(bad, SAMNeRFConfig.tiny())
    bad = dict(params)
    bad["field.mlp_base.params"] = torch.zeros(cfg.field_mlp_params + 12345)
    with pytest.raises(ValueError, match="do not fit"):
        ck.infer_config(bad, SAMNeRFConfig.tiny())
    torch.save({"not": "a checkpoint"}, tmp_path / "x.ckpt")
    with pytest.raises(KeyError, match="not a trainer checkpoint"):
        ck.load_checkpoint(str(tmp_path / "x.ckpt"))
    os.makedirs(tmp_path / "empty")
    with pytest.raises(FileNotFoundError):
        ck.latest_checkpoint(str(tmp_path / "empty"))


def test_feature_files(tmp_path):
    """sam_features/*.npy are [256,h,w]; lookups floor the scaled pixel position (feature_loader.py:45-53)."""
    scene = tmp_path / "scene"
    (scene / "images").mkdir(parents=True)
    (scene / "sam_features").mkdir()
    (scene / "clipseg_features").mkdir()
    names = [str(scene / "images" / f"frame_{i:05d}.JPG") for i in range(2)]
    rng = np.random.default_rng(0)
    maps = [rng.standard_normal((256, 43, 64)).astype(np.float32) for _ in names]
    npy = ck.feature_filenames(names, "sam_features", ".npy")
    assert npy[1].endswith(os.path.join("scene", "sam_features", "frame_00001.npy"))
    for p, m in zip(npy, maps):
        np.save(p, m)
    loader = ck.FeatureDataloader("cpu", npy, image_shape=(1060, 1600), patch_size=4)
    assert tuple(loader.features.shape) == (2, 43, 64, 256)
    pts = torch.tensor([[0, 0, 0], [1, 1059, 1599], [1, 530, 800], [0, 24, 25]])
    got = loader(pts)
    for row, (i, y, x) in zip(got, pts.tolist()):
        fy, fx = int(y * (43 / 1060)), int(x * (64 / 1600))
        assert np.array_equal(row.numpy(), maps[i][:, fy, fx])
    # ClipSeg activations: 3 layers of [1 + 32*32 tokens, 1, 64] -> [32,32,192] (datamanager.py:92-94)
    acts = [torch.randn(1025, 1, 64, generator=torch.Generator().manual_seed(i)) for i in range(3)]
    pt = ck.feature_filenames(names[:1], "clipseg_features", ".pt")
    torch.save({"activations": acts}, pt[0])
    cl = ck.FeatureDataloader("cpu", pt, image_shape=(1060, 1600), get_feature=ck.clipseg_activations_to_map)
    assert tuple(cl.features.shape) == (1, 32, 32, 192)
    assert torch.equal(cl.features[0, 3, 5], torch.cat([a[1 + 3 * 32 + 5, 0] for a in acts]))
    # the short side of SAM's 64x64 embedding is cropped to the image's aspect (get_image_embeddings.py:23-35)
    emb = torch.zeros(1, 256, 64, 64)
    assert ck.crop_sam_embedding(emb, 1060, 1600).shape[-2:] == (43, 64)
    assert ck.crop_sam_embedding(emb, 1600, 1060).shape[-2:] == (64, 43)
    assert ck.crop_sam_embedding(emb, 800, 800).shape[-2:] == (64, 64)


def test_model_from_checkpoint(tmp_path, monkeypatch):
    """SAMModel.from_checkpoint: configuration inferred from the file, parameters handed to the renderer; after a
    training step, state_dict() -> save_checkpoint -> from_checkpoint carries the trained tensors."""
    import samnerf_b200.nerfstudio_api as api
    from fake_renderer import FakeRenderer

    monkeypatch.setattr(api, "Renderer", FakeRenderer)
    cfg = SAMNeRFConfig.tiny(clipseg=False, patch_size=1)
    params = {k: v for k, v in make_synthetic_params(cfg, "scene", 4).items() if not k.startswith("conv_head")}
    path = ck.save_checkpoint(str(tmp_path) + os.sep, params, step=1234, ddp=True)
    m = api.SAMModel.from_checkpoint(str(tmp_path), base=SAMNeRFConfig.tiny(clipseg=False, patch_size=1))
    assert m.step == 1234 and m.config == cfg and path.endswith("step-000001234.ckpt")
    for k, v in params.items():
        assert torch.equal(m.renderer.p[k], v), k
    m.train()
    with torch.no_grad():
        m.params["sam_field.sam_net.params"].add_(0.25)
    sd = m.state_dict()
    ck.save_checkpoint(str(tmp_path) + os.sep, sd, step=1300)
    m2 = api.SAMModel.from_checkpoint(str(tmp_path), base=SAMNeRFConfig.tiny(clipseg=False, patch_size=1))
    assert m2.step == 1300
    assert torch.allclose(m2.renderer.p["sam_field.sam_net.params"], params["sam_field.sam_net.params"] + 0.25)
    assert torch.equal(m2.renderer.p["field.mlp_base.params"], params["field.mlp_base.params"])


def _strict_load(target_keys_shapes, state):
    """What ``nn.Module.load_state_dict(strict=True)`` enforces (Trainer._load_checkpoint -> load_pipeline(strict=True),
    trainer.py:372-376, base_pipeline.py:366-375): no missing key, no unexpected key, equal shapes."""
    missing = sorted(set(target_keys_shapes) - set(state))
    unexpected = sorted(set(state) - set(target_keys_shapes))
    assert not missing and not unexpected, (missing, unexpected)
    for k, shp in target_keys_shapes.items():
        assert list(state[k].shape) == list(shp), (k, tuple(state[k].shape), shp)


@pytest.mark.parametrize("name,cfg", [("tiny_distill_clipseg_p4", SAMNeRFConfig.tiny(clipseg=True, patch_size=4)),
                                      ("full_distill_p4", None)])
def test_saved_checkpoint_passes_the_reference_trainers_strict_load(tmp_path, name, cfg):
    """Resume path: starting from a reference checkpoint (every key of the reference's own modules, the camera
    optimizer, Adam state, grad-scaler state), ``save_checkpoint(source=...)`` must hand back a file the reference's
    strict loader accepts, with the hot-path tensors replaced and everything else untouched."""
    layout = _layout(name)
    if cfg is None:  # full size: keep the test light - meta-sized fake tensors are enough for key / shape logic
        pytest.importorskip("torch")
        small = {k: v for k, v in layout.items() if int(np.prod(v)) < 1_000_000}
        layout = small
    gen = torch.Generator().manual_seed(5)
    pipe = {k: torch.randn(*shp, generator=gen) if shp else torch.tensor(7) for k, shp in layout.items()}
    opt = {"fields": {"state": {0: {"step": torch.tensor(3.0), "exp_avg": torch.randn(5, generator=gen)}}, "param_groups": [{"lr": 1e-2}]}}
    scaler = {"scale": 65536.0, "growth_factor": 2.0, "backoff_factor": 0.5, "growth_interval": 2000, "_growth_tracker": 0}
    src = tmp_path / "step-000000100.ckpt"
    torch.save({"step": 100, "pipeline": pipe, "optimizers": opt, "scalers": scaler}, src)

    hot = {ck.strip_prefixes(k): v for k, v in pipe.items() if ck.strip_prefixes(k) in ck.HOT_PATH_KEYS}
    assert hot, "layout has no hot-path tensors?"
    trained = {k: v + 1.0 for k, v in hot.items()}
    out = ck.save_checkpoint(str(tmp_path / "out") + os.sep, trained, step=200, source=str(src))
    got = torch.load(out, map_location="cpu", weights_only=True)
    assert got["step"] == 200
    _strict_load(layout, got["pipeline"])
    for k, v in pipe.items():
        bare = ck.strip_prefixes(k)
        want = trained[bare] if bare in trained else v
        assert torch.equal(got["pipeline"][k], want), k
    # optimizer and scaler state survive: GradScaler.load_state_dict({}) raises in the reference (mixed precision is on
    # in both shipped configs), and Adam moments must not be lost on resume
    assert got["scalers"] == scaler and torch.equal(got["optimizers"]["fields"]["state"][0]["exp_avg"], opt["fields"]["state"][0]["exp_avg"])
    with pytest.raises(KeyError, match="not in the source"):
        ck.save_checkpoint(str(tmp_path / "out") + os.sep, {"sam_field.bogus.params": torch.zeros(3)}, step=201, source=str(src))


def test_eval_export_has_every_model_key_of_the_reference(tmp_path):
    """Without a source checkpoint: the model part of the export is complete (strict load of the MODEL passes: the
    geometry buffers and the parameter-free encodings are synthesised from the config); what lives outside the model
    cannot be invented, and the file says so by carrying empty optimizer / scaler dicts (eval-only)."""
    cfg = SAMNeRFConfig.tiny(clipseg=True, patch_size=4)
    layout = {k: v for k, v in _layout("tiny_distill_clipseg_p4").items() if k.startswith("_model.")}
    params = make_synthetic_params(cfg, "init", 2)
    out = ck.save_checkpoint(str(tmp_path) + os.sep, params, step=9, cfg=cfg)
    got = torch.load(out, map_location="cpu", weights_only=True)
    _strict_load(layout, got["pipeline"])
    assert got["optimizers"] == {} and got["scalers"] == {}
    cfg2, params2, step = ck.load_checkpoint(out, base=SAMNeRFConfig.tiny())
    assert step == 9 and cfg2 == cfg and all(torch.equal(params2[k], params[k]) for k in params)
