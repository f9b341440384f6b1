"""Data-parallel training step over 2 ranks (gloo, CPU): every rank back-propagates its own rays through the shim
(oracle forward, emulated kernel bodies backward), `all_reduce_gradients` averages the flat gradients like torch DDP in
the reference (samnerf/train.py:171-199), and the result equals the single-process gradient of the joint batch."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _step(rays_slice, seed_jitter=7):
    import samnerf_b200.nerfstudio_api as api
    from fake_renderer import FakeRenderer
    from helpers import model_pair, test_rays

    api.Renderer = FakeRenderer
    cfg, params, _ = model_pair("tiny", "scene", 25, False, 1)
    m = api.SAMModel(cfg)
    m.load_state_dict(params)
    m.train()
    m.proposal_sampler.train_stratified = False  # same samples in the joint and the split runs
    o, d = test_rays(32, seed=2)
    g = torch.Generator().manual_seed(0)
    image = torch.rand(32, 3, generator=g)
    o, d, image = o[rays_slice], d[rays_slice], image[rays_slice]
    out = m(api.RayBundle(origins=o, directions=d), get_feature=[])
    torch.nn.functional.mse_loss(out["rgb"], image).backward()
    return m


def _worker(rank, world, port, ret):
    import sys

    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        m = _step(slice(rank * 16, (rank + 1) * 16))
        m.all_reduce_gradients()
        ret[rank] = {k: v.grad.clone() for k, v in m.params.items() if v.grad is not None}
    finally:
        dist.destroy_process_group()


def test_two_rank_gradients_equal_the_joint_batch():
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, 2, 29731, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    joint = _step(slice(0, 32))
    for name, p in joint.params.items():
        if p.grad is None:
            continue
        g0, g1 = ret[0][name], ret[1][name]
        assert torch.equal(g0, g1), name                       # every rank holds the same averaged gradient
        scale = float(p.grad.abs().max())
        assert scale > 0 and float((g0 - p.grad).abs().max()) <= 1e-5 * scale, name  # mean over 32 = mean of the two means over 16
