"""Loss terms of the training loop (samnerf_b200/training.py) against the reference's own
``interlevel_loss`` / ``distortion_loss`` (tests/golden/losses.npz, oracle/make_loss_golden.py) and the loss / metric
dictionaries and per-iteration callbacks of the model shim (CPU, oracle-backed fake renderer)."""
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from samnerf_b200 import training as T

Z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "losses.npz"))


def _lists(w0=None):
    e0, e1 = torch.from_numpy(Z["e0"]), torch.from_numpy(Z["e1"])
    w0 = torch.from_numpy(Z["w0"]) if w0 is None else w0
    rs = [SimpleNamespace(spacing_starts=e[:, :-1, None], spacing_ends=e[:, 1:, None]) for e in (e0, e1)]
    return [w0[..., None], torch.from_numpy(Z["w1"])[..., None]], rs


def test_interlevel_and_distortion_match_reference():
    ws, rs = _lists()
    assert abs(float(T.interlevel_loss(ws, rs)) - float(Z["interlevel"])) <= 1e-7 + 1e-5 * float(Z["interlevel"])
    assert abs(float(T.distortion_loss(ws, rs)) - float(Z["distortion"])) <= 1e-5 * float(Z["distortion"])
    w0 = torch.from_numpy(Z["w0"]).clone().requires_grad_(True)
    ws, rs = _lists(w0)
    T.interlevel_loss(ws, rs).backward()
    np.testing.assert_allclose(w0.grad.numpy(), Z["interlevel_grad_w0"], rtol=1e-4, atol=1e-9)


def test_schedules():
    assert T.proposal_anneal(0) == 0.0 and T.proposal_anneal(1000) == 1.0 and T.proposal_anneal(5000) == 1.0
    assert abs(T.proposal_anneal(100) - (10 * 0.1) / (9 * 0.1 + 1)) < 1e-12
    sched = T.proposal_update_schedule()
    assert sched(0) == 1.0 and sched(5000) == 5.0 and sched(2500) == 2.5 and sched(10**6) == 5.0


def test_model_loss_dict_and_callbacks(monkeypatch):
    import samnerf_b200.nerfstudio_api as api
    from fake_renderer import FakeRenderer
    from helpers import model_pair, test_rays

    monkeypatch.setattr(api, "Renderer", FakeRenderer)
    cfg, params, _ = model_pair("tiny", "scene", 25, False, 1)
    m = api.SAMModel(cfg)
    m.load_state_dict(params)
    m.train()
    m.before_train_iteration(100)
    assert abs(m.renderer.anneal - T.proposal_anneal(100)) < 1e-7 and m.proposal_sampler._anneal == m.renderer.anneal
    o, d = test_rays(32, seed=2)
    g = torch.Generator().manual_seed(0)
    batch = {"image": torch.rand(32, 3, generator=g), "sam": torch.randn(32, 256, generator=g) * 0.1}
    torch.manual_seed(5)
    out = m(api.RayBundle(origins=o, directions=d), get_feature=["sam"])
    metrics = m.get_metrics_dict(out, batch)
    losses = m.get_loss_dict(out, batch, metrics)
    assert set(losses) == {"rgb_loss", "interlevel_loss", "distortion_loss", "sam_loss"} and set(metrics) == {"psnr", "distortion"}
    assert abs(float(losses["distortion_loss"].detach()) - cfg.distortion_loss_mult * float(metrics["distortion"].detach())) < 1e-9
    total = sum(losses.values())
    total.backward()
    groups = m.get_param_groups()
    for name in ("proposal_networks", "fields", "sam_field"):
        assert all(p.grad is not None and torch.isfinite(p.grad).all() and float(p.grad.abs().max()) > 0 for p in groups[name]), name
    m.after_train_iteration(100)
    assert m.proposal_sampler._step == 100 and m.proposal_sampler._steps_since_update == 1
    # the same two hooks as trainer callbacks (nerfacto.py:242-271, engine/callbacks.py:49-110)
    cbs = m.get_training_callbacks()
    assert [c.where_to_run[0] for c in cbs] == [T.TrainingCallbackLocation.BEFORE_TRAIN_ITERATION, T.TrainingCallbackLocation.AFTER_TRAIN_ITERATION]
    for c in cbs:
        c.run_callback_at_location(step=500, location=T.TrainingCallbackLocation.BEFORE_TRAIN_ITERATION)
    assert abs(m.renderer.anneal - T.proposal_anneal(500)) < 1e-7 and m.proposal_sampler._step == 100
    cbs[1].run_callback_at_location(step=101, location=T.TrainingCallbackLocation.AFTER_TRAIN_ITERATION)
    assert m.proposal_sampler._step == 101 and m.proposal_sampler._steps_since_update == 2
    m.proposal_sampler._steps_since_update = 1
    # step 100 >= 10 and one step since the last update <= schedule(100) = 1: this pass keeps the proposal net frozen
    for p in groups["proposal_networks"]:
        p.grad = None
    out = m(api.RayBundle(origins=o, directions=d), get_feature=[])
    assert not out["weights_list"][0].requires_grad and out["weights_list"][1].requires_grad
    m.eval()
    ev = m(api.RayBundle(origins=o, directions=d), get_feature=[])
    assert set(m.get_loss_dict(ev, batch)) == {"rgb_loss"} and set(m.get_metrics_dict(ev, batch)) == {"psnr"}
