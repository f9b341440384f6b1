"""Parity of the CUDA path (through the C ABI) against the CPU oracle and the committed golden fixtures.

Run on the B200 box with ``pytest -m gpu``.  Sizes are chosen so the oracle finishes in seconds; full-size
behaviour is covered through size-independent properties in test_gpu_properties.py.
"""
import os

import numpy as np
import pytest
import torch

from helpers import (FRAC_DISCRETE, FRAC_SMOOTH, TOL, assert_features_close, assert_mostly_close, make_renderer,
                     model_pair, test_rays)

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def tiny():
    cfg, params, orc = model_pair("tiny", "scene", 11, True, 1)
    return cfg, params, orc, make_renderer(cfg, params)


@pytest.fixture(scope="module")
def full():
    cfg, params, orc = model_pair("full", "scene", 0, False, 1)
    return cfg, params, orc, make_renderer(cfg, params)


@pytest.fixture(scope="module")
def full_clipseg():
    """The shipped full-size configuration with the ClipSeg grids and head (BASELINE.json configs[3])."""
    cfg, params, orc = model_pair("full", "scene", 2, True, 1)
    return cfg, params, orc, make_renderer(cfg, params)


def _positions(n, seed):
    g = torch.Generator().manual_seed(seed)
    inner = (torch.rand(n // 2, 3, generator=g) * 2 - 1) * 0.9
    outer = torch.randn(n - n // 2, 3, generator=g) * 4.0
    return torch.cat([inner, outer]).contiguous()


# ---- component level: the hash-grid gathers and MLP heads on given positions ------------------------------------
@pytest.mark.parametrize("which", ["tiny", "full"])
def test_density_fields(which, request):
    cfg, params, orc, r = request.getfixturevalue(which)
    x = _positions(4096, 1)
    d_ref = orc.proposal_density(x)
    d_gpu, _ = r.query_density("proposal", x)
    assert_mostly_close(d_gpu[..., 0], d_ref, TOL["density"], FRAC_SMOOTH, "proposal density")
    f_ref, geo_ref = orc.field_density(x)
    f_gpu, geo_gpu = r.query_density("field", x)
    assert_mostly_close(f_gpu[..., 0], f_ref, TOL["density"], FRAC_SMOOTH, "field density")
    assert_mostly_close(geo_gpu.float(), geo_ref, dict(rtol=1e-2, atol=2e-3), FRAC_SMOOTH, "geo features")
    dirs = torch.nn.functional.normalize(torch.randn(4096, 3, generator=torch.Generator().manual_seed(2)), dim=-1)
    rgb_ref = orc.field_rgb(dirs, geo_ref)
    rgb_gpu = r.query_rgb(dirs, geo_ref.to(torch.float16))
    assert_mostly_close(rgb_gpu, rgb_ref, dict(rtol=0.0, atol=2e-3), FRAC_SMOOTH, "per-sample rgb")


@pytest.mark.parametrize("which", ["tiny", "full", "full_clipseg"])
def test_feature_field(which, request):
    cfg, params, orc, r = request.getfixturevalue(which)
    x = _positions(2048, 3)
    ref = orc.sam_field(x, which=("sam", "clipseg") if cfg.use_clipseg_feature else ("sam",))
    hg, sam = r.query_features("sam", x)
    assert_mostly_close(hg.float(), ref["hashgrid"], TOL["encoding"], FRAC_SMOOTH, "hashgrid encoding")
    assert_mostly_close(sam, ref["sam"], TOL["features"], FRAC_SMOOTH, "per-sample sam")
    if cfg.use_clipseg_feature:
        _, cs = r.query_features("clipseg", x)
        assert_mostly_close(cs, ref["clipseg"], TOL["features"], FRAC_SMOOTH, "per-sample clipseg")


# ---- stage level: sampler, weights, compositing, top-k ------------------------------------------------------------
@pytest.mark.parametrize("which,regime", [("tiny", "scene"), ("full", "scene"), ("tiny", "init"), ("full_clipseg", "scene")])
def test_render_stages(which, regime):
    if which == "full_clipseg":
        cfg, params, orc = model_pair("full", regime, 2, True, 1)
    else:
        cfg, params, orc = model_pair(which, regime, 11 if which == "tiny" else 0, which == "tiny", 1)
    r = make_renderer(cfg, params)
    o, d = test_rays(1024, seed=5)
    feats = ("sam", "clipseg") if cfg.use_clipseg_feature else ("sam",)
    ref = orc.render_rays(o, d, get_feature=feats, return_intermediates=True)
    out = r.render(o, d, get_feature=feats, debug=True)
    torch.cuda.synchronize()
    assert_mostly_close(out["_prop_weights"], ref["_w0"], TOL["weights"], FRAC_SMOOTH, "proposal weights")
    assert_mostly_close(out["_edges"], ref["_eu1"], TOL["edges"], FRAC_SMOOTH, "nerf bin edges")
    assert_mostly_close(out["_density"], ref["_density"], dict(rtol=3e-2, atol=1e-3), 0.99, "nerf density")
    assert_mostly_close(out["_weights"], ref["_weights"], TOL["weights"], 0.99, "nerf weights")
    assert_mostly_close(out["_rgb_samples"], ref["_rgb_s"], dict(rtol=0.0, atol=4e-3), 0.99, "per-sample rgb")
    assert_mostly_close(out["rgb"], ref["rgb"], TOL["rgb"], 0.995, "rgb", per_row=True)
    assert_mostly_close(out["accumulation"], ref["accumulation"], TOL["accumulation"], 0.995, "accumulation")
    assert_mostly_close(out["depth"], ref["depth"], TOL["depth"], FRAC_DISCRETE, "median depth")
    assert_mostly_close(out["prop_depth_0"], ref["prop_depth_0"], TOL["depth"], FRAC_DISCRETE, "proposal depth")
    # sharpened weights: compare as sets (topk(sorted=False), sam_model.py:244) via the sorted values
    sw_ref = torch.sort(ref["_sam_weights"], dim=-1, descending=True).values
    sw_gpu = torch.sort(out["_sam_w"].cpu(), dim=-1, descending=True).values
    assert_mostly_close(sw_gpu, sw_ref, dict(rtol=5e-2, atol=2e-3), FRAC_DISCRETE, "sharpened top-k weights", per_row=True)
    assert_features_close(out["sam"], ref["sam"], "sam feature")
    if cfg.use_clipseg_feature:
        assert_features_close(out["clipseg"], ref["clipseg"], "clipseg feature")


@pytest.mark.parametrize("engine", ["tcgen05", "mma_sync"])
def test_engines_agree_with_oracle(engine):
    """Both tensor-core engines against the oracle, and hence against each other."""
    cfg, params, orc = model_pair("tiny", "scene", 11, True, 1)
    r = make_renderer(cfg, params, engine=engine)
    o, d = test_rays(515, seed=9)  # not a multiple of 8: exercises the tail tile of the feature kernel
    ref = orc.render_rays(o, d, get_feature=("sam", "clipseg"))
    out = r.render(o, d, get_feature=("sam", "clipseg"))
    assert_features_close(out["sam"], ref["sam"], f"sam[{engine}]")
    assert_features_close(out["clipseg"], ref["clipseg"], f"clipseg[{engine}]")


# ---- golden fixtures produced by the reference's own Python (oracle/make_golden.py) ----------------------------
@pytest.mark.parametrize("name", ["chunk_tiny_scene", "chunk_tiny_init", "chunk_tiny_patch4", "chunk_full_scene", "chunk_full_4k",
                                  "chunk_full_clipseg_4k"])
def test_golden_chunks(name):
    from oracle.make_golden import fixture_specs, make_cfg, params_checksum
    from samnerf_b200 import make_synthetic_params

    spec = fixture_specs()[name]
    cfg = make_cfg(spec)
    params = make_synthetic_params(cfg, spec["regime"], spec["seed"])
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    assert abs(params_checksum(params) - float(z["_params_checksum"])) <= 1e-9 * float(z["_params_checksum"])
    r = make_renderer(cfg, params)
    feats = ("sam", "clipseg") if cfg.use_clipseg_feature else ("sam",)
    out = r.render(torch.from_numpy(z["_origins"]), torch.from_numpy(z["_directions"]), get_feature=feats,
                   patch=cfg.patch_size > 1)
    assert_mostly_close(out["rgb"], z["rgb"], TOL["rgb"], 0.99, "rgb", per_row=True)
    assert_mostly_close(out["accumulation"], z["accumulation"], TOL["accumulation"], 0.99, "accumulation")
    # budgets for the discrete picks were 0.95 in round 1; the observed fractions on B200 are >= 0.992 everywhere
    # (profiles/r02_parity_observed.json), so they are held to FRAC_DISCRETE = 0.99 now
    assert_mostly_close(out["depth"], z["depth"], TOL["depth"], FRAC_DISCRETE, "depth")
    assert_mostly_close(out["prop_depth_0"], z["prop_depth_0"], TOL["depth"], FRAC_DISCRETE, "prop depth")
    if cfg.patch_size == 1:
        assert_features_close(out["sam"], z["sam"], "sam", row_frac=FRAC_DISCRETE)
    else:  # 16 patch-aggregated rows: one flipped ray moves a whole row, so allow 1 of 16 (observed: 0)
        assert_features_close(out["sam"], z["sam"], "sam", row_frac=0.93, elem_frac=0.93, rel_l2=3e-2,
                              tol=dict(rtol=3e-2, atol=3e-3))
    if "clipseg" in z.files:
        assert_features_close(out["clipseg"], z["clipseg"], "clipseg", row_frac=FRAC_DISCRETE)


def test_golden_image_through_model_shim():
    """SAMModel.get_outputs_for_camera_ray_bundle (chunk loops A/B/C + patch head) against the reference image."""
    from oracle.make_golden import fixture_specs, make_cfg
    from samnerf_b200 import make_synthetic_params
    from samnerf_b200.nerfstudio_api import RayBundle, SAMModel

    spec = fixture_specs()["image_tiny"]
    cfg = make_cfg(spec)
    params = make_synthetic_params(cfg, spec["regime"], spec["seed"])
    z = np.load(os.path.join(GOLDEN, "image_tiny.npz"))
    m = SAMModel(cfg)
    m.load_state_dict(params)
    o, d = torch.from_numpy(z["_origins"]), torch.from_numpy(z["_directions"])
    bundle = RayBundle(origins=o, directions=d, pixel_area=torch.ones_like(o[..., :1]),
                       camera_indices=torch.zeros_like(o[..., :1]).long())
    out = m.get_outputs_for_camera_ray_bundle(bundle)
    assert out["rgb"].shape == (24, 32, 3) and out["sam"].shape == (48, 64, 256) and out["clipseg"].shape == (32, 32, 192)
    assert_mostly_close(out["rgb"], z["rgb"], TOL["rgb"], 0.99, "rgb", per_row=False)
    assert_mostly_close(out["depth"], z["depth"], TOL["depth"], FRAC_DISCRETE, "depth")
    st = int(z["_sam_stride"])
    assert_features_close(out["sam"][::st, ::st], z["sam"], "patch-aggregated sam", row_frac=0.97, elem_frac=0.97,
                          rel_l2=3e-2, tol=dict(rtol=3e-2, atol=3e-3))
    assert_features_close(out["clipseg"], z["clipseg"], "clipseg", row_frac=FRAC_DISCRETE)


# ---- patch aggregation kernel on its own -----------------------------------------------------------------------
@pytest.mark.parametrize("engine", ["tcgen05", "mma_sync"])
def test_patch_aggregate(engine):
    cfg, params, orc = model_pair("tiny", "scene", 5, False, 4)
    r = make_renderer(cfg, params, engine=engine)
    g = torch.Generator().manual_seed(3)
    feat = torch.randn(37 * 16, 256, generator=g) * 0.3  # 37 patches: not a multiple of the 8-patch tile
    ref = orc.patch_aggregate(feat)
    out = r.patch_aggregate(feat)
    # fp16 operands (the precision class of the TF32 convs cuDNN runs for the reference), fp32 accumulate, K = 2304
    assert_mostly_close(out, ref, dict(rtol=2e-2, atol=3e-3), FRAC_SMOOTH, f"patch head[{engine}]")


# ---- renderer / sampler shims -------------------------------------------------------------------------------------
def test_component_shims(tiny):
    from samnerf_b200.nerfstudio_api import (AccumulationRenderer, DepthRenderer, MeanRenderer, ProposalNetworkSampler,
                                            RayBundle, RGBRenderer, NearFarCollider)
    from oracle import samnerf_oracle as O

    cfg, params, orc, r = tiny
    o, d = test_rays(300, seed=21)
    bundle = NearFarCollider(0.05, cfg.far_plane)(RayBundle(origins=o.cuda(), directions=d.cuda()))
    rs, w_list, rs_list = ProposalNetworkSampler(r)(bundle, density_fns=None)
    ref = orc.render_rays(o, d, get_feature=(), return_intermediates=True)
    assert_mostly_close(w_list[0][..., 0], ref["_w0"], TOL["weights"], FRAC_SMOOTH, "weights_list[0]")
    assert_mostly_close(rs.frustums.starts[..., 0], ref["_eu1"][:, :-1], TOL["edges"], FRAC_SMOOTH, "starts")
    assert_mostly_close(rs_list[0].frustums.ends[..., 0], ref["_eu0"][:, 1:], dict(rtol=1e-5, atol=1e-6), 1.0, "level-0 ends")
    # renderers on oracle-provided inputs: exact same inputs, so tight tolerances
    w = ref["_weights"]
    dens, deltas = ref["_density"], ref["_eu1"][:, 1:] - ref["_eu1"][:, :-1]
    rs.deltas = deltas[..., None].cuda()
    assert_mostly_close(rs.get_weights(dens[..., None].cuda())[..., 0], O.get_weights(deltas, dens),
                        dict(rtol=1e-4, atol=1e-6), 1.0, "get_weights")
    assert_mostly_close(AccumulationRenderer(r)(w[..., None].cuda()), w.sum(-1, keepdim=True),
                        dict(rtol=1e-5, atol=1e-6), 1.0, "accumulation")
    rs.frustums.starts, rs.frustums.ends = ref["_eu1"][:, :-1, None].cuda(), ref["_eu1"][:, 1:, None].cuda()
    assert_mostly_close(DepthRenderer(r)(w[..., None].cuda(), rs), O.median_depth(w, ref["_eu1"][:, :-1], ref["_eu1"][:, 1:]),
                        dict(rtol=1e-6, atol=1e-6), 0.995, "median depth")
    assert_mostly_close(RGBRenderer(r)(ref["_rgb_s"].cuda(), w[..., None].cuda()), O.composite_rgb(ref["_rgb_s"], w),
                        dict(rtol=1e-5, atol=1e-6), 1.0, "rgb composite")
    assert_mostly_close(RGBRenderer(r, background_color="white")(ref["_rgb_s"].cuda(), w[..., None].cuda()),
                        O.composite_rgb(ref["_rgb_s"], w, background=(1.0, 1.0, 1.0)), dict(rtol=1e-5, atol=1e-6), 1.0, "rgb/white")
    emb = torch.randn(300, 16, 24, generator=torch.Generator().manual_seed(1))
    sw = torch.rand(300, 16, generator=torch.Generator().manual_seed(2))
    assert_mostly_close(MeanRenderer(r)(emb.cuda(), sw[..., None].cuda()), (sw[..., None] * emb).sum(-2),
                        dict(rtol=1e-5, atol=1e-5), 1.0, "mean renderer")


def test_training_mode_sampler_matches_reference():
    """snrf_sample with snrf_set_jitter (march_kernel<.., JIT>) against the reference's own ProposalNetworkSampler
    run in training mode with recorded torch.rand draws (tests/golden/sampler_training.npz)."""
    import os

    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sampler_training.npz"))
    from samnerf_b200 import SAMNeRFConfig, make_synthetic_params

    cfg = SAMNeRFConfig.tiny(clipseg=False, patch_size=1)
    params8 = make_synthetic_params(cfg, "scene", 8)
    r = make_renderer(cfg, params8)
    o, d, jit = (torch.from_numpy(z[k]) for k in ("_origins", "_directions", "jitter"))
    w0, edges1, _ = r.sample(o, d, jitter=jit)
    torch.cuda.synchronize()
    assert_mostly_close(w0, z["w0"], TOL["weights"], FRAC_SMOOTH, "training-mode proposal weights")
    assert_mostly_close(edges1, z["edges1"], TOL["edges"], 0.99, "training-mode nerf bin edges")
    # one-shot: the next call is an eval-mode call again, and a mismatched ray count is refused
    w0_eval, edges_eval, _ = r.sample(o, d)
    ref_eval = r.sample(o, d)
    assert torch.equal(edges_eval, ref_eval[1]) and float((edges_eval - edges1).abs().max()) > 1e-3
    with pytest.raises(RuntimeError, match="snrf_set_jitter"):
        r.sample(o[:10], d[:10], jitter=jit)
    out = r.render(o, d, get_feature=(), debug=True, jitter=jit)
    assert_mostly_close(out["_edges"], z["edges1"], TOL["edges"], 0.99, "render() with jitter uses the same samples")
    # annealed proposal weights in front of the PDF sampler (set_anneal, ray_samplers.py:583); reported weights stay raw
    r.set_anneal(float(z["anneal"]))
    try:
        w0a, edges_a, _ = r.sample(o, d, jitter=torch.from_numpy(z["jitter_anneal"]))
        # the exponent applies in every mode (ray_samplers.py:583 has no training / eval branch): an eval render
        # during the first 1000 training steps resamples from the annealed weights too
        edges_ev_a = r.sample(o, d)[1]
    finally:
        r.set_anneal(1.0)
    assert_mostly_close(w0a, z["w0_anneal"], TOL["weights"], FRAC_SMOOTH, "raw proposal weights under annealing")
    assert_mostly_close(edges_a, z["edges1_anneal"], TOL["edges"], 0.99, "nerf bin edges under annealing")
    from oracle.samnerf_oracle import Oracle

    ref_ev_a = Oracle(cfg, params8).render_rays(o, d, get_feature=(), return_intermediates=True, anneal=float(z["anneal"]))
    assert float((edges_ev_a.cpu() - edges_eval.cpu()).abs().max()) > 1e-3
    assert_mostly_close(edges_ev_a, ref_ev_a["_eu1"], TOL["edges"], 0.99, "eval-mode nerf bin edges under annealing")


def test_boundary_inputs_against_reference_golden():
    """Per-ray nears / fars, background override and fast mode through snrf_render against the reference's own model
    (tests/golden/chunk_tiny_boundary.npz)."""
    from samnerf_b200 import SAMNeRFConfig, make_synthetic_params

    z = np.load(os.path.join(GOLDEN, "chunk_tiny_boundary.npz"))
    cfg = SAMNeRFConfig.tiny(clipseg=False, patch_size=1)
    r = make_renderer(cfg, make_synthetic_params(cfg, "scene", 9))
    o, d, nr, fr = (torch.from_numpy(z[k]) for k in ("_origins", "_directions", "_nears", "_fars"))
    bg = tuple(float(v) for v in z["_bg"])
    for fast in (False, True):
        out = r.render(o, d, nears=nr, fars=fr, get_feature=("sam",), fast=fast, background=bg)
        torch.cuda.synchronize()
        pre = "fast." if fast else ""
        assert_mostly_close(out["rgb"], z[pre + "rgb"], TOL["rgb"], 0.99, pre + "rgb", per_row=True)
        assert_mostly_close(out["depth"], z[pre + "depth"], TOL["depth"], FRAC_DISCRETE - 0.02, pre + "depth")
        assert_features_close(out["sam"], z[pre + "sam"], pre + "sam", row_frac=FRAC_DISCRETE - 0.02)
        if not fast:
            assert_mostly_close(out["accumulation"], z["accumulation"], TOL["accumulation"], 0.99, "accumulation")
            assert_mostly_close(out["prop_depth_0"], z["prop_depth_0"], TOL["depth"], FRAC_DISCRETE - 0.02, "prop_depth_0")
        else:
            assert "accumulation" not in out
