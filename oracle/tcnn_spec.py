"""CPU restatement of the tiny-cuda-nn modules the SAM-NeRF hot path calls.

TEST INFRASTRUCTURE ONLY - this file is the checker, never the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs may import it.

PARITY UNPINNED AT THIS BOUNDARY: ``tinycudann`` (NVlabs/tiny-cuda-nn, PyTorch bindings) is an
un-vendored, un-pinned, CUDA-only third-party dependency of the reference (install hint only:
``nerfstudio/utils/printing.py:24-33``; call sites ``samnerf/sam_field.py:51,63,84,99``,
``nerfstudio/fields/nerfacto_field.py:144,152,157,228``, ``nerfstudio/fields/density_fields.py:92,99``).
It is absent from /root/reference and from this image and the reference ships no tests or golden
vectors for it, so what follows restates its *published* algorithm:

* ``HashGrid``      - ``grid.h``: ``scale = exp2f(l*log2f(pls))*base - 1``; ``res = ceil(scale)+1``;
                      ``pos = x*scale + 0.5``; dense index ``x + y*res + z*res^2`` while the stride fits the
                      level, else the "coherent prime" hash ``x*1 ^ y*2654435761 ^ z*805459861`` (uint32);
                      ``% level_size``; level size ``min(round_up(res^3, 8), 2^log2T)``; trilinear blend.
                      In-tree cross-check of the indexing: ``nerfstudio/field_components/cuda/csrc/
                      temporal_gridencoder.cu:46-59`` (same primes / xor) and ``:62-88`` (dense-else-hash, ``%``).
* ``FullyFusedMLP`` / ``CutlassMLP`` - fp16 row-major ``[out, in]`` weights, no bias, ReLU hidden layers,
                      widths padded to 16; padded input columns are 0 after a grid encoding and 1 after
                      the identity encoding of a plain ``tcnn.Network``.
* ``SphericalHarmonics(4)`` - input ``[0,1] -> [-1,1]``, 16 real SH values with tcnn's sign pattern.

Precision model (stated, because the real library cannot be run here): parameters are rounded to fp16
(tcnn keeps fp16 copies of its fp32 ``params``); every module *output* is rounded to fp16 (tcnn outputs
``__half``); hidden activations are rounded to fp16 between layers (tcnn stores them as ``__half``);
all sums are accumulated in fp32 (tcnn's grid kernel and FullyFusedMLP accumulate in fp16 - the fp32
accumulation here is the *more* exact reading, and is what the CUDA path does too).
"""
from __future__ import annotations

import math
from typing import List, Sequence, Tuple

import numpy as np
import torch

PRIME_Y = 2654435761
PRIME_Z = 805459861
MASK32 = 0xFFFFFFFF


class _RoundF16(torch.autograd.Function):
    """fp16 rounding with a straight-through fp32 gradient (a plain ``.half().float()`` round trip would also round
    the *gradient* to fp16 on the way back, which is neither tcnn's loss-scaled fp16 backward nor useful as a
    reference; the stated gradient contract is "fp32 gradient of the rounded forward pass")."""

    @staticmethod
    def forward(ctx, x):
        return x.to(torch.float16).to(torch.float32)

    @staticmethod
    def backward(ctx, g):
        return g


def f16(x: torch.Tensor) -> torch.Tensor:
    """Round to fp16 and come back to fp32 (models a ``__half`` store)."""
    if x.requires_grad:
        return _RoundF16.apply(x)
    return x.to(torch.float16).to(torch.float32)


def grid_levels(n_levels: int, base_resolution: int, per_level_scale: float, log2_hashmap_size: int):
    """Per level ``(scale, res, offset, size, hashed)``; float32 arithmetic like ``grid.h``."""
    log2_pls = np.float32(np.log2(np.float32(per_level_scale)))
    out = []
    offset = 0
    for l in range(n_levels):
        scale = np.float32(np.exp2(np.float32(l) * log2_pls)) * np.float32(base_resolution) - np.float32(1.0)
        res = int(math.ceil(float(scale))) + 1
        dense = res**3
        size = min((dense + 7) // 8 * 8, 1 << log2_hashmap_size)
        out.append((float(scale), res, offset, size, dense > size))
        offset += size
    return out


def hash_grid_encode(
    x: torch.Tensor, table: torch.Tensor, levels: Sequence[Tuple[float, int, int, int, bool]], n_features: int
) -> torch.Tensor:
    """``x[N,3]`` fp32 in [0,1] -> ``[N, L*F]`` fp32 holding fp16-rounded values, level-major.

    ``table``: flat fp32 grid parameters (``n_entries * F``); rounded to fp16 here.
    """
    assert x.dtype == torch.float32 and x.shape[-1] == 3
    n = x.shape[0]
    tab = f16(table).view(-1, n_features)
    outs = []
    for scale, res, offset, size, hashed in levels:
        pos = x * np.float32(scale) + np.float32(0.5)
        g = torch.floor(pos)
        fr = pos - g
        g = g.to(torch.int64)
        acc = torch.zeros(n, n_features, dtype=torch.float32)
        for corner in range(8):
            w = torch.ones(n, dtype=torch.float32)
            c = []
            for d in range(3):
                if corner & (1 << d):
                    w = w * fr[:, d]
                    c.append(g[:, d] + 1)
                else:
                    w = w * (1.0 - fr[:, d])
                    c.append(g[:, d])
            if hashed:
                idx = (c[0] & MASK32) ^ ((c[1] * PRIME_Y) & MASK32) ^ ((c[2] * PRIME_Z) & MASK32)
            else:
                idx = (c[0] + c[1] * res + c[2] * res * res) & MASK32
            idx = idx % size + offset
            acc = acc + w[:, None] * tab[idx]
        outs.append(f16(acc))
    return torch.cat(outs, dim=-1)


def mlp_forward(
    x: torch.Tensor, weights: Sequence[torch.Tensor], output_activation: str = "None"
) -> torch.Tensor:
    """fp16-in / fp32-accumulate / fp16-out MLP.  ``weights[i]``: fp32 ``[out_i, in_i]`` (rounded to fp16 here).
    ``x``: fp32 holding fp16-representable values, width == ``in_0``.  ReLU between layers."""
    h = x
    for i, w in enumerate(weights):
        h = h @ f16(w).T
        if i + 1 < len(weights):
            h = torch.relu(h)
        h = f16(h)
    if output_activation == "Sigmoid":
        h = f16(torch.sigmoid(h))
    elif output_activation != "None":
        raise ValueError(output_activation)
    return h


def split_mlp_params(flat: torch.Tensor, dims: Sequence[int]) -> List[torch.Tensor]:
    """Carve ``[out,in]`` row-major matrices for layer widths ``dims = [in, h1, ..., out]`` out of a flat tensor."""
    mats, o = [], 0
    for i in range(len(dims) - 1):
        n = dims[i + 1] * dims[i]
        mats.append(flat[o : o + n].view(dims[i + 1], dims[i]))
        o += n
    assert o == flat.numel(), (o, flat.numel())
    return mats


def sh4(dirs01: torch.Tensor) -> torch.Tensor:
    """tcnn ``SphericalHarmonics`` degree 4 on inputs in [0,1]; returns fp16-rounded fp32 ``[N,16]``."""
    d = dirs01 * 2.0 - 1.0
    x, y, z = d[:, 0], d[:, 1], d[:, 2]
    xy, xz, yz = x * y, x * z, y * z
    x2, y2, z2 = x * x, y * y, z * z
    o = [
        torch.full_like(x, 0.28209479177387814),
        -0.48860251190291987 * y,
        0.48860251190291987 * z,
        -0.48860251190291987 * x,
        1.0925484305920792 * xy,
        -1.0925484305920792 * yz,
        0.94617469575755997 * z2 - 0.31539156525251999,
        -1.0925484305920792 * xz,
        0.54627421529603959 * x2 - 0.54627421529603959 * y2,
        0.59004358992664352 * y * (-3.0 * x2 + y2),
        2.8906114426405538 * xy * z,
        0.45704579946446572 * y * (1.0 - 5.0 * z2),
        0.3731763325901154 * z * (5.0 * z2 - 3.0),
        0.45704579946446572 * x * (1.0 - 5.0 * z2),
        1.4453057213202769 * z * (x2 - y2),
        0.59004358992664352 * x * (-x2 + 3.0 * y2),
    ]
    return f16(torch.stack(o, dim=-1))
