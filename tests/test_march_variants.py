"""The two alternative structures of the march kernel (csrc/march.cu: field MLPs as tcgen05 tiles, SNRF_MARCH_TC=1; sampling
half and field half as two launches, SNRF_MARCH_SPLIT=1) against the default fused mma.sync kernel on the same rays.

Both are opt-in (they measured slower, profiles/r02_kernel_experiments.txt) but stay selectable, so they stay tested:
the split kernel runs the default kernel's arithmetic in two launches and must reproduce it bit for bit; the tcgen05
variant accumulates the same fp16 products in a different order and is held to the tolerances of the oracle comparison."""
import pytest
import torch

from helpers import FRAC_DISCRETE, FRAC_SMOOTH, TOL, assert_features_close, assert_mostly_close, make_renderer, model_pair, test_rays


def _render(monkeypatch, env, n=1536):
    cfg, params, _ = model_pair("full", "scene", 7, False, 1)  # shipped grid geometry: the variants exist for it only
    o, d = test_rays(n, seed=5)
    monkeypatch.delenv("SNRF_MARCH_TC", raising=False)
    monkeypatch.delenv("SNRF_MARCH_SPLIT", raising=False)
    want = make_renderer(cfg, params).render(o, d, get_feature=("sam",))
    monkeypatch.setenv(env, "1")  # read by snrf_ctx_create
    got = make_renderer(cfg, params).render(o, d, get_feature=("sam",))
    torch.cuda.synchronize()
    return got, want


@pytest.mark.gpu
def test_split_march_is_bit_identical(monkeypatch):
    got, want = _render(monkeypatch, "SNRF_MARCH_SPLIT")
    for k in ("rgb", "depth", "accumulation", "sam"):
        assert torch.equal(got[k], want[k]), k


@pytest.mark.gpu
def test_tcgen05_march_matches_mma_sync(monkeypatch):
    got, want = _render(monkeypatch, "SNRF_MARCH_TC")
    assert_mostly_close(got["rgb"], want["rgb"], TOL["rgb"], FRAC_SMOOTH, "rgb", per_row=True)
    assert_mostly_close(got["accumulation"], want["accumulation"], TOL["accumulation"], FRAC_SMOOTH, "accumulation")
    assert_mostly_close(got["depth"], want["depth"], TOL["depth"], FRAC_DISCRETE, "depth")
    assert_features_close(got["sam"], want["sam"], "sam", row_frac=FRAC_DISCRETE)
