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


def test_bucket_prepass_body():
    """Pre-pass of the bucketed feature kernel (csrc/sam_bucket.cu): every ray lands in exactly one bucket, the
    bucket covers all of its significant slots, NaN rays keep all 16."""
    import ctypes as C

    import numpy as np

    from emu.build_emu import load

    n = 5000
    g = torch.Generator().manual_seed(0)
    w = torch.rand(n, 16, generator=g) ** 12  # a few large, many tiny
    w, _ = (w / w.sum(-1, keepdim=True)).sort(dim=-1, descending=True)
    w[torch.rand(n, 16, generator=g) < 0.2] = 0.0
    w, _ = w.sort(dim=-1, descending=True)
    w[7] = float("nan")
    arr = np.ascontiguousarray(w.numpy(), np.float32)
    for eps in (0.0, 2.0 ** -24, 1e-4):
        counts = np.zeros(5, np.int32)
        lists = np.full((5, n), -1, np.int32)
        load().emu_bucket_assign(arr.ctypes.data_as(C.c_void_p), C.c_float(eps), C.c_longlong(n),
                                 counts.ctypes.data_as(C.c_void_p), lists.ctypes.data_as(C.c_void_p))
        assert counts.sum() == n
        seen = np.concatenate([lists[b, :counts[b]] for b in range(5)])
        assert np.array_equal(np.sort(seen), np.arange(n))
        sig = (~(w < eps)) & (w != 0)
        k = torch.where(sig.any(-1), 16 - sig.flip(-1).float().argmax(-1), torch.zeros(n, dtype=torch.long)).numpy()
        for b, slots in enumerate((1, 2, 4, 8, 16)):
            rays = lists[b, :counts[b]]
            assert (k[rays] <= slots).all() and (b == 0 or (k[rays] > slots // 2).all())
            # what is dropped is below the cut-off
            if slots < 16 and len(rays):
                dropped = torch.nan_to_num(w[torch.from_numpy(rays).long(), slots:]).sum(-1)
                assert float(dropped.max()) <= 16 * eps
        assert 7 in lists[4, :counts[4]]


def test_bucket_kernel_index_arithmetic_model():
    """A Python model of the warp-level index arithmetic of sam_bucket_kernel (it cannot be emulated like the
    per-thread kernels): global tile -> (bucket, local tile), tile row -> (list entry, slot) in the gather, TMEM row
    -> lane in the epilogue, the log2(SLOTS)-step halving reduction with its column base, and the store - every
    (ray, column) of hbar must be written exactly once with the sum over the bucket's slots of w * h.  The shuffles
    are modelled as lane permutations of numpy arrays."""
    import numpy as np

    rng = np.random.default_rng(0)
    n_rays = 700
    h = rng.standard_normal((n_rays, 16, 64)).astype(np.float32)   # stands for relu(W1 x) of (ray, slot); 64 of the 256 columns
    w = rng.random((n_rays, 16)).astype(np.float32)
    # bucket lists as the pre-pass leaves them: [5][n_rays] with counts (ragged tiles in every bucket)
    perm = rng.permutation(n_rays)
    counts = [128 + 5, 64 * 2 + 3, 32 + 7, 16 * 3 + 1, 8 + 2]
    lists = np.full((5, n_rays), -1, np.int64)
    o = 0
    for b, c in enumerate(counts):
        lists[b, :c] = perm[o:o + c]
        o += c
    hbar = np.full((n_rays, 64), np.nan, np.float32)
    writes = np.zeros((n_rays, 64), np.int32)
    t1 = (counts[0] + 127) >> 7
    t2 = t1 + ((counts[1] + 63) >> 6)
    t3 = t2 + ((counts[2] + 31) >> 5)
    t4 = t3 + ((counts[3] + 15) >> 4)
    t5 = t4 + ((counts[4] + 7) >> 3)

    def bucket_of(t):
        for b, (hi, first) in enumerate(((t1, 0), (t2, t1), (t3, t2), (t4, t3))):
            if t < hi:
                return b, first, counts[b]
        return 4, t4, counts[4]

    for t in range(t5):
        b, first, count = bucket_of(t)
        tile = t - first
        acc = np.zeros((128, 64), np.float32)
        s_sw = np.zeros(128, np.float32)
        for warp in range(8):  # e = 0 warps write the weights; both encodings write disjoint A columns
            for lane in range(32):
                row = (warp & 7) * 16 + (lane >> 1)
                r_loc, slot = row >> b, row & ((1 << b) - 1)
                li = tile * (128 >> b) + r_loc
                valid = li < count
                ray = lists[b, li if valid else count - 1]
                assert ray >= 0
                s_sw[row] = w[ray, slot] if valid else 0.0
                acc[row] = h[ray, slot]
        log = b
        for warp in range(4):  # the 64 modelled columns = column group cq = 0; quarter = warp & 3
            quarter = warp & 3
            for chunk in range(2):
                col0 = chunk * 32
                v = np.zeros((32, 32), np.float32)  # [lane][i]
                for lane in range(32):
                    erow = quarter * 32 + lane
                    v[lane] = acc[erow, col0:col0 + 32] * s_sw[erow]
                base = np.zeros(32, np.int32)
                for step in range(log):
                    half = 16 >> step
                    nv = v.copy()
                    for lane in range(32):
                        up = (lane >> step) & 1
                        peer = lane ^ (1 << step)
                        peer_up = (peer >> step) & 1
                        for i in range(half):
                            keep = v[lane, i + half] if up else v[lane, i]
                            recv = v[peer, i] if peer_up else v[peer, i + half]  # what the peer sends
                            nv[lane, i] = keep + recv
                        base[lane] += half if up else 0
                    v = nv
                keep_n = 32 >> log
                for lane in range(32):
                    erow = quarter * 32 + lane
                    li = tile * (128 >> log) + (erow >> log)
                    if li < count:
                        ray = lists[b, li]
                        c = col0 + base[lane]
                        hbar[ray, c:c + keep_n] = v[lane, :keep_n]
                        writes[ray, c:c + keep_n] += 1
    used = perm[:sum(counts)]
    assert (writes[used] == 1).all() and writes.sum() == sum(counts) * 64
    o = 0
    for b, c in enumerate(counts):
        rays = perm[o:o + c]
        o += c
        want = (w[rays, :1 << b, None] * h[rays, :1 << b]).sum(1)
        np.testing.assert_allclose(hbar[rays], want, rtol=1e-5, atol=1e-5)
