"""Size-independent properties of the CUDA path at BASELINE.json's full sizes, and the edge cases of the boundary."""
import pytest
import torch

from helpers import make_renderer, model_pair, test_rays

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def full():
    cfg, params, orc = model_pair("full", "scene", 0, False, 1)
    return cfg, make_renderer(cfg, params)


def test_full_frame_properties(full):
    """800x800 frame (640 000 rays, 20 chunks of 32 768): invariants the reference's math guarantees."""
    from samnerf_b200.synthetic import orbit_rays

    cfg, r = full
    o, d = orbit_rays()
    o, d = o.reshape(-1, 3).cuda(), d.reshape(-1, 3).cuda()
    chunk = cfg.eval_num_rays_per_chunk
    outs = [r.render(o[i:i + chunk], d[i:i + chunk], get_feature=("sam",), debug=(i == 0)) for i in range(0, o.shape[0], chunk)]
    torch.cuda.synchronize()
    rgb = torch.cat([x["rgb"] for x in outs])
    acc = torch.cat([x["accumulation"] for x in outs])
    depth = torch.cat([x["depth"] for x in outs])
    sam = torch.cat([x["sam"] for x in outs])
    assert rgb.shape == (640000, 3) and sam.shape == (640000, 256)
    assert torch.isfinite(rgb).all() and float(rgb.min()) >= 0.0 and float(rgb.max()) <= 1.0
    assert float(acc.min()) >= 0.0 and float(acc.max()) <= 1.0 + 1e-5        # sum of alpha*T never exceeds 1
    assert float(depth.min()) >= 0.0 and float(depth.max()) <= cfg.far_plane
    d0 = outs[0]
    e = d0["_edges"]
    assert bool((e[:, 1:] >= e[:, :-1]).all()), "PDF samples must be monotone along the ray"
    assert bool((d0["_weights"] >= 0).all()) and bool((d0["_prop_weights"] >= 0).all())
    sw = d0["_sam_w"]
    ok = torch.isfinite(sw).all(dim=-1)
    assert float(ok.float().mean()) > 0.99
    assert torch.allclose(sw[ok].sum(-1), torch.ones_like(sw[ok].sum(-1)), atol=1e-5), "top-k weights renormalise to 1"
    assert bool((sw[ok][:, :-1] >= sw[ok][:, 1:]).all()), "slot order is descending weight"
    assert float(torch.isfinite(sam).all(dim=-1).float().mean()) > 0.99


def test_chunking_and_determinism(full):
    """Rays are independent: any chunking, any position in the batch, any repeat gives the same bits."""
    cfg, r = full
    o, d = test_rays(5000, seed=3)
    o, d = o.cuda(), d.cuda()
    a = r.render(o, d, get_feature=("sam",))
    b = r.render(o, d, get_feature=("sam",))
    parts = [r.render(o[i:i + 777], d[i:i + 777], get_feature=("sam",)) for i in range(0, 5000, 777)]
    perm = torch.randperm(5000, generator=torch.Generator().manual_seed(1)).cuda()
    c = r.render(o[perm], d[perm], get_feature=("sam",))
    torch.cuda.synchronize()
    for k in ("rgb", "depth", "accumulation", "prop_depth_0", "sam"):
        assert torch.equal(a[k], b[k]), f"{k}: not deterministic"
        assert torch.equal(a[k], torch.cat([p[k] for p in parts])), f"{k}: depends on chunking"
        assert torch.equal(a[k][perm], c[k]), f"{k}: depends on position in the batch"


def test_bricks_are_a_pure_relayout(full):
    """Cell-major "brick" copies of the leading grid levels (snrf_set_brick_budget, csrc/bricks.cu) hold verbatim table
    entries: any budget - none, the dense levels only, the default prefix - renders the same bits, per-sample debug
    outputs included."""
    cfg, r = full
    o, d = test_rays(3000, seed=8)
    o, d = o.cuda(), d.cuda()
    outs, levels = {}, {}
    try:
        for gb in (0.0, 0.02, 1.5, 4.0):
            levels[gb] = r.set_brick_budget(gb)
            outs[gb] = r.render(o, d, get_feature=("sam",), debug=True)
            torch.cuda.synchronize()
    finally:
        r.set_brick_budget(4.0)
    assert levels[0.0] == (0, 0) and levels[4.0][0] == 5 and levels[4.0][1] >= 10, levels
    assert levels[0.0] < levels[0.02] < levels[1.5] <= levels[4.0], levels
    for gb in (0.02, 1.5, 4.0):
        for k in ("rgb", "depth", "accumulation", "prop_depth_0", "sam", "_prop_weights", "_edges", "_density", "_weights"):
            assert torch.equal(outs[0.0][k], outs[gb][k]), f"{k}: bricks ({gb} GB -> {levels[gb]} levels) change the result"


def test_fp16_rows_and_march_first_are_the_same_frame(full):
    """snrf_set_feature_dtype(fp16): the rows are the fp32 rows rounded once to fp16, bit for bit (TMA-fed output layer
    with the fp16 epilogue); snrf_set_march_first: one march launch over the tile + feature chunks == alternating chunks."""
    cfg, r = full
    o, d = test_rays(7001, seed=12)  # ragged: 7 chunks of 1024 less a few rays
    o, d = o.cuda(), d.cuda()
    base = r.render_frame(o, d, get_feature=("sam",), chunk=1024)
    try:
        r.set_march_first(True)
        mf = r.render_frame(o, d, get_feature=("sam",), chunk=1024)
        r.set_feature_dtype(torch.float16)
        h = r.render_frame(o, d, get_feature=("sam",), chunk=1024)
        hc = r.render(o[:515], d[:515], get_feature=("sam",))
        torch.cuda.synchronize()
    finally:
        r.set_march_first(False)
        r.set_feature_dtype(torch.float32)
    for k in ("rgb", "depth", "accumulation", "prop_depth_0", "sam"):
        assert torch.equal(torch.nan_to_num(base[k]), torch.nan_to_num(mf[k])), f"{k}: march-first changes the frame"
    assert h["sam"].dtype == torch.float16 and hc["sam"].dtype == torch.float16
    assert torch.equal(torch.nan_to_num(h["sam"]), torch.nan_to_num(base["sam"].half())), "fp16 rows != fp32 rows rounded once"
    assert torch.equal(torch.nan_to_num(hc["sam"]), torch.nan_to_num(base["sam"][:515].half()))
    assert torch.equal(h["rgb"], base["rgb"])


def test_edge_cases(full):
    cfg, r = full
    o, d = test_rays(64, seed=4)
    # empty batch
    out = r.render(o[:0], d[:0], get_feature=("sam",))
    assert out["rgb"].shape == (0, 3) and out["sam"].shape == (0, 256)
    # single ray / ragged sizes around the warp-per-ray CTA (8) and the 8-ray feature tile
    ref = r.render(o, d, get_feature=("sam",))
    for n in (1, 7, 9, 63):
        part = r.render(o[:n], d[:n], get_feature=("sam",))
        for k in ("rgb", "depth", "sam"):
            assert torch.equal(part[k], ref[k][:n]), (k, n)
    # caller-supplied nears / fars (viewer crop box path, scene_colliders.py:40-44) equal to the defaults
    nf = r.render(o, d, nears=torch.zeros(64, 1), fars=torch.full((64, 1), cfg.far_plane), get_feature=("sam",))
    assert torch.equal(nf["rgb"], ref["rgb"]) and torch.equal(nf["sam"], ref["sam"])
    # a tighter far plane changes the samples
    near2 = r.render(o, d, nears=torch.full((64, 1), 0.05), fars=torch.full((64, 1), 6.0), get_feature=())
    assert float(near2["depth"].max()) <= 6.0
    # fixed background only matters where accumulation < 1
    white = r.render(o, d, get_feature=(), background=(1.0, 1.0, 1.0))
    black = r.render(o, d, get_feature=(), background=(0.0, 0.0, 0.0))
    diff = (white["rgb"] - black["rgb"]).sum(-1, keepdim=True) / 3.0
    assert torch.allclose(diff, (1.0 - ref["accumulation"]).clamp(0, 1), atol=2e-3)
    # fast mode drops accumulation / prop_depth_0 (sam_model.py:284-299) and leaves the rest untouched
    fast = r.render(o, d, get_feature=(), fast=True)
    assert "accumulation" not in fast and "prop_depth_0" not in fast and torch.equal(fast["rgb"], ref["rgb"])


def test_errors_are_loud(full):
    cfg, r = full
    o, d = test_rays(16, seed=5)
    with pytest.raises(RuntimeError, match="clipseg"):
        # this renderer was built without ClipSeg parameters
        import ctypes as C
        from samnerf_b200 import _lib as L
        o_, d_ = o.cuda(), d.cuda()
        rgb, dep = torch.empty(16, 3, device="cuda"), torch.empty(16, 1, device="cuda")
        cs = torch.empty(16, 192, device="cuda")
        opts = r._opts()
        rc = r.lib.snrf_render(r.h, o_.data_ptr(), d_.data_ptr(), None, None, 16, L.WANT_CLIPSEG, C.byref(opts),
                               rgb.data_ptr(), dep.data_ptr(), None, None, None, cs.data_ptr(), None, r.stream)
        r._check(rc)
    with pytest.raises(RuntimeError, match="divisible by 16"):
        r.render(o[:10], d[:10], get_feature=("sam",), patch=True)


def test_frame_api_equals_chunk_loop(full):
    """snrf_render_frame (sequential and 3-stream pipelined) == the per-chunk loop, bit for bit."""
    cfg, r = full
    o, d = test_rays(9000, seed=8)
    o, d = o.cuda(), d.cuda()
    ref = [r.render(o[i:i + 2048], d[i:i + 2048], get_feature=("sam",)) for i in range(0, 9000, 2048)]
    for mode in (0, 2):
        r.set_pipeline(mode)
        out = r.render_frame(o, d, get_feature=("sam",), chunk=2048)
        torch.cuda.synchronize()
        for k in ("rgb", "depth", "accumulation", "prop_depth_0", "sam"):
            assert torch.equal(out[k], torch.cat([x[k] for x in ref])), (k, mode)
    r.set_pipeline(1)


def test_replication_descriptor_is_validated(full):
    import ctypes as C

    cfg, r = full
    t = torch.empty(16, 256, device="cuda")
    assert r.lib.snrf_set_replication(r.h, 7, t.data_ptr(), t.numel() * 4, None, None, 0) != 0       # bad slot
    assert r.lib.snrf_set_replication(r.h, 0, t.data_ptr(), t.numel() * 4, None, None, 9) != 0       # too many peers
    assert r.lib.snrf_set_replication(r.h, 0, None, 0, None, None, 0) == 0                           # clear is fine
    # replicating into a second buffer on the same device through the peer-pointer path mirrors the rows exactly
    o, d = test_rays(256, seed=2)
    frame, mirror = torch.zeros(256, 256, device="cuda"), torch.zeros(256, 256, device="cuda")
    r.set_replication("sam", frame, peer_ptrs=[mirror.data_ptr()])
    out = r.render(o, d, get_feature=("sam",), out={"sam": frame})
    torch.cuda.synchronize()
    r.set_replication("sam", None)
    assert torch.equal(frame, mirror) and torch.isfinite(frame).any()


def test_early_termination_stays_within_its_bounds(full):
    """Opt-in early termination (snrf_set_early_termination): rgb / accumulation within eps, depth and features
    unchanged up to the eps-weighted tail; switching it off restores the exact bits."""
    cfg, r = full
    o, d = test_rays(6000, seed=12)
    o, d = o.cuda(), d.cuda()
    exact = r.render(o, d, get_feature=("sam",))
    eps = 1e-4
    r.set_early_termination(eps)
    try:
        et = r.render(o, d, get_feature=("sam",))
        torch.cuda.synchronize()
    finally:
        r.set_early_termination(0.0)
    assert float((et["rgb"] - exact["rgb"]).abs().max()) <= 2 * eps
    assert float((et["accumulation"] - exact["accumulation"]).abs().max()) <= 2 * eps
    assert float((et["depth"] == exact["depth"]).float().mean()) > 0.999
    ok = torch.isfinite(exact["sam"]).all(-1) & torch.isfinite(et["sam"]).all(-1)
    rel = torch.linalg.norm(et["sam"][ok] - exact["sam"][ok], dim=-1) / torch.linalg.norm(exact["sam"][ok], dim=-1).clamp_min(1e-6)
    assert float((rel < 1e-3).float().mean()) > 0.999
    again = r.render(o, d, get_feature=("sam",))
    assert torch.equal(again["rgb"], exact["rgb"]) and torch.equal(again["sam"], exact["sam"])
    with pytest.raises(RuntimeError, match="threshold"):
        r.set_early_termination(0.7)


@pytest.mark.parametrize("cutoff,rel", [(0.0, 2e-4), (2.0 ** -24, 2e-4)])
def test_bucketed_feature_kernel_matches_kernel_b(full, cutoff, rel):
    """snrf_set_feature_cutoff: rays bucketed by their significant-slot count give the same feature rows as the
    kernel that evaluates all 16 slots - up to fp32 summation order (cut-off 0) / one fp32 ulp of the sum (2^-24)."""
    cfg, r = full
    o, d = test_rays(9000, seed=14)
    o, d = o.cuda(), d.cuda()
    try:
        r.set_feature_cutoff(-1.0)  # every slot of every ray (csrc/sam.cu)
        exact = r.render(o, d, get_feature=("sam",))
        r.set_feature_cutoff(cutoff)
        got = r.render(o, d, get_feature=("sam",))
        part = r.render(o[:1003], d[:1003], get_feature=("sam",))  # ragged: partial tiles in every bucket
        torch.cuda.synchronize()
        r.set_feature_cutoff(-1.0)
        again = r.render(o, d, get_feature=("sam",))
    finally:
        r.set_feature_cutoff(2.0 ** -24)  # the library default
    for k in ("rgb", "depth", "accumulation"):
        assert torch.equal(got[k], exact[k]), k
    a, b = got["sam"], exact["sam"]
    assert torch.equal(torch.isnan(a), torch.isnan(b))
    ok = torch.isfinite(b).all(-1)
    err = (a[ok] - b[ok]).abs().max(-1).values / b[ok].abs().max(-1).values.clamp_min(1e-6)
    # the per-ray hidden sum is stored in fp16: a different fp32 summation order flips an entry by one fp16 ulp (2^-11 of
    # that entry) now and then - a few % of the rays carry such a flip, worth ~1e-5..1e-4 of the ray's largest channel
    assert float(err.max()) <= 2e-3, float(err.max())
    assert float((err <= rel).float().mean()) > 0.99, float((err <= rel).float().mean())
    assert float((err == 0).float().mean()) > 0.5  # most rays agree bit for bit
    assert torch.equal(part["sam"], got["sam"][:1003])
    assert torch.equal(again["sam"], exact["sam"])
