"""Multi-rank tile sharding and exchange on CPU (gloo, world_size 2 and 3): the host-side logic of bench.py --gpus N."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from samnerf_b200.tiles import all_gather_tiles, ray_block, row_block


def test_row_blocks_cover_the_frame_and_respect_patches():
    for h, world, align in [(800, 1, 1), (800, 2, 1), (800, 8, 4), (1060, 8, 4), (43 * 4, 3, 4), (7, 8, 1)]:
        blocks = [row_block(r, world, h, align) for r in range(world)]
        assert blocks[0][0] == 0 and blocks[-1][1] == h
        for (a0, a1), (b0, b1) in zip(blocks, blocks[1:]):
            assert a1 == b0 and a0 <= a1
        for a0, a1 in blocks[:-1]:
            assert a0 % align == 0 and a1 % align == 0
    # the bench's case: 800 rows over 2/4/8 ranks are equal blocks (SURVEY 8 d config 5)
    for world in (2, 4, 8):
        assert {b - a for a, b in (row_block(r, world, 800) for r in range(world))} == {800 // world}
    assert ray_block(1, 2, 800, 800) == (320000, 640000)


def _worker(rank, world, port, h, w, align, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = {"rgb": torch.full((h * w, 3), -1.0), "sam": torch.full((h * w, 8), -1.0)}
        lo, hi = ray_block(rank, world, h, w, align)
        idx = torch.arange(lo, hi, dtype=torch.float32)
        full["rgb"][lo:hi] = idx[:, None] * torch.tensor([1.0, 2.0, 3.0])   # "render" this rank's tile
        full["sam"][lo:hi] = idx[:, None] + torch.arange(8.0)
        all_gather_tiles(full, h, w, align)
        all_idx = torch.arange(h * w, dtype=torch.float32)
        ok = torch.equal(full["rgb"], all_idx[:, None] * torch.tensor([1.0, 2.0, 3.0])) and torch.equal(
            full["sam"], all_idx[:, None] + torch.arange(8.0))
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,h,w,align", [(2, 16, 5, 1), (2, 12, 7, 4), (3, 16, 4, 4)])
def test_tiles_exchange_over_gloo(world, h, w, align):
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = 29600 + world * 10 + h
    procs = [ctx.Process(target=_worker, args=(r, world, port, h, w, align, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert all(ret.get(r) for r in range(world)), dict(ret)
