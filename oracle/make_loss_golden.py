"""Generate ``tests/golden/losses.npz`` with the REFERENCE'S OWN ``interlevel_loss`` / ``distortion_loss``
(nerfstudio/model_components/losses.py:33-143) on seeded sample lists.

TEST INFRASTRUCTURE ONLY.  Runs only in the build container (needs /root/reference):

    python -m oracle.make_loss_golden
"""
from __future__ import annotations

import os
from types import SimpleNamespace

import numpy as np
import torch

from oracle.make_golden import GOLDEN, _install_stubs


def sample_lists(seed: int = 0, n: int = 64):
    g = torch.Generator().manual_seed(seed)

    def level(s):
        edges = torch.sort(torch.rand(n, s + 1, generator=g), dim=-1).values
        edges[:, 0], edges[:, -1] = 0.0, 1.0
        w = torch.rand(n, s, generator=g) ** 3
        w = w / w.sum(-1, keepdim=True) * torch.rand(n, 1, generator=g)
        return edges, w

    return level(64), level(32)


def main():
    _install_stubs()
    from nerfstudio.model_components.losses import distortion_loss, interlevel_loss

    (e0, w0), (e1, w1) = sample_lists()
    rs = [SimpleNamespace(spacing_starts=e[:, :-1, None], spacing_ends=e[:, 1:, None]) for e in (e0, e1)]
    ws = [w0[..., None], w1[..., None]]
    out = {"e0": e0.numpy(), "w0": w0.numpy(), "e1": e1.numpy(), "w1": w1.numpy(),
           "interlevel": np.array(float(interlevel_loss(ws, rs))), "distortion": np.array(float(distortion_loss(ws, rs)))}
    w0g = w0.clone().requires_grad_(True)
    interlevel_loss([w0g[..., None], ws[1]], rs).backward()
    out["interlevel_grad_w0"] = w0g.grad.numpy()
    path = os.path.join(GOLDEN, "losses.npz")
    np.savez_compressed(path, **out)
    print({k: v.shape for k, v in out.items()}, float(out["interlevel"]), float(out["distortion"]), os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
