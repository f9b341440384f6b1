"""Screen-tile sharding of a frame across ranks (SURVEY.md section 8 e / config 5).

Rays are independent, so a frame is cut into contiguous row blocks, one per rank; the only cross-ray operation on the
path is the p x p patch conv, so blocks are aligned to ``align_rows`` (the patch size when the patch head is on).
After rendering, the blocks are exchanged with ONE all-gather per output tensor.  The same code runs over NCCL
(GPU ranks) and gloo (CPU tests).  The reference has no multi-GPU render (sam_model.py:358-364 loops on one device).
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.distributed as dist


def row_block(rank: int, world: int, height: int, align_rows: int = 1) -> Tuple[int, int]:
    """Rows ``[r0, r1)`` of rank ``rank``: ``height`` rows in ``world`` contiguous blocks, each a multiple of
    ``align_rows`` except possibly the last; earlier ranks take the extra units."""
    units = (height + align_rows - 1) // align_rows
    base, extra = divmod(units, world)
    u0 = rank * base + min(rank, extra)
    u1 = u0 + base + (1 if rank < extra else 0)
    return min(u0 * align_rows, height), min(u1 * align_rows, height)


def ray_block(rank: int, world: int, height: int, width: int, align_rows: int = 1) -> Tuple[int, int]:
    """Row-major ray range ``[lo, hi)`` of this rank's tile."""
    r0, r1 = row_block(rank, world, height, align_rows)
    return r0 * width, r1 * width


def all_gather_tiles(full: Dict[str, torch.Tensor], height: int, width: int, align_rows: int = 1,
                     group: Optional[dist.ProcessGroup] = None) -> None:
    """In-place exchange: every rank has written rows of its own block into ``full[name]`` (``[H*W, C]``); afterwards
    every rank holds the whole frame.  Equal blocks use one ``all_gather_into_tensor`` per tensor (in place, no
    staging); ragged blocks fall back to ``all_gather`` on per-rank views."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if world == 1:
        return
    blocks = [ray_block(r, world, height, width, align_rows) for r in range(world)]
    equal = len({hi - lo for lo, hi in blocks}) == 1
    lo, hi = blocks[rank]
    for t in full.values():
        if equal:
            dist.all_gather_into_tensor(t, t[lo:hi], group=group)
        else:
            _ragged_gather(t, blocks, rank, group)


def _ragged_gather(t: torch.Tensor, blocks, rank: int, group) -> None:
    """Ragged blocks: broadcast each rank's block from its owner (world is small: <= 8)."""
    for r, (a, b) in enumerate(blocks):
        if b > a:
            dist.broadcast(t[a:b], src=dist.get_global_rank(group, r) if group is not None else r, group=group)
