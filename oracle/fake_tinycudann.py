"""A CPU stand-in for the ``tinycudann`` PyTorch bindings, built on ``oracle/tcnn_spec.py``.

TEST INFRASTRUCTURE ONLY.  ``oracle/make_golden.py`` injects this module as ``sys.modules["tinycudann"]``
so that the reference's own field classes (``HashMLPDensityField``, ``TCNNNerfactoField``, ``SAMField``)
run unmodified on the CPU in the build container.  Same public surface as the real bindings for the
calls the reference makes: ``Encoding(n_input_dims, encoding_config)``, ``Network(n_input_dims,
n_output_dims, network_config)``, ``NetworkWithInputEncoding(n_input_dims, n_output_dims, encoding_config,
network_config)``; each an ``nn.Module`` with one flat fp32 ``params`` Parameter (network weights first,
then the grid), ``n_input_dims`` / ``n_output_dims`` attributes and a forward returning fp16.
PARITY UNPINNED against the real library (see ``tcnn_spec.py``).
"""
from __future__ import annotations

import math

import torch
from torch import nn

from . import tcnn_spec as T


def _pad16(n: int) -> int:
    return (n + 15) // 16 * 16


class Encoding(nn.Module):
    def __init__(self, n_input_dims, encoding_config, dtype=None):
        super().__init__()
        self.n_input_dims = n_input_dims
        self.cfg = dict(encoding_config)
        otype = self.cfg["otype"]
        if otype == "HashGrid":
            self.levels = T.grid_levels(
                self.cfg["n_levels"],
                self.cfg["base_resolution"],
                float(self.cfg["per_level_scale"]),
                self.cfg["log2_hashmap_size"],
            )
            self.n_features = self.cfg["n_features_per_level"]
            self.n_output_dims = self.cfg["n_levels"] * self.n_features
            n = (self.levels[-1][2] + self.levels[-1][3]) * self.n_features
            self.params = nn.Parameter((torch.rand(n) * 2 - 1) * 1e-4)
        elif otype == "SphericalHarmonics":
            assert self.cfg["degree"] == 4
            self.n_output_dims = 16
            self.params = nn.Parameter(torch.zeros(0))
        elif otype == "Frequency":
            self.n_output_dims = n_input_dims * 2 * self.cfg["n_frequencies"]
            self.params = nn.Parameter(torch.zeros(0))
        else:
            raise NotImplementedError(otype)

    def forward(self, x):
        otype = self.cfg["otype"]
        x = x.to(torch.float32)
        if otype == "HashGrid":
            return T.hash_grid_encode(x, self.params.detach(), self.levels, self.n_features).to(torch.float16)
        if otype == "SphericalHarmonics":
            return T.sh4(x).to(torch.float16)
        raise NotImplementedError(otype)  # Frequency only feeds the pred-normals MLP (off on this path)


class Network(nn.Module):
    """``tcnn.Network`` == NetworkWithInputEncoding with the identity encoding: padded inputs are 1."""

    pad_value = 1.0

    def __init__(self, n_input_dims, n_output_dims, network_config):
        super().__init__()
        self.n_input_dims = n_input_dims
        self.n_output_dims = n_output_dims
        c = dict(network_config)
        assert c["activation"] == "ReLU"
        self.out_act = c["output_activation"]
        width, hidden = c["n_neurons"], c["n_hidden_layers"]
        self.dims = [_pad16(n_input_dims)] + [width] * hidden + [_pad16(n_output_dims)]
        n = sum(self.dims[i] * self.dims[i + 1] for i in range(len(self.dims) - 1))
        ws = []
        for i in range(len(self.dims) - 1):
            a = math.sqrt(6.0 / (self.dims[i] + self.dims[i + 1]))
            ws.append((torch.rand(self.dims[i + 1] * self.dims[i]) * 2 - 1) * a)
        self.params = nn.Parameter(torch.cat(ws))
        assert self.params.numel() == n

    def _mlp(self, x):
        pad = self.dims[0] - x.shape[-1]
        if pad:
            x = torch.cat([x, torch.full((x.shape[0], pad), self.pad_value)], dim=-1)
        ws = T.split_mlp_params(self.params.detach(), self.dims)
        return T.mlp_forward(x, ws, self.out_act)[:, : self.n_output_dims]

    def forward(self, x):
        return self._mlp(T.f16(x.to(torch.float32))).to(torch.float16)


class NetworkWithInputEncoding(nn.Module):
    """Grid encoding + MLP sharing one flat ``params`` (network first, then encoding); grid pads with 0."""

    def __init__(self, n_input_dims, n_output_dims, encoding_config, network_config):
        super().__init__()
        self.n_input_dims = n_input_dims
        self.n_output_dims = n_output_dims
        enc = Encoding(n_input_dims, encoding_config)
        net = Network(enc.n_output_dims, n_output_dims, network_config)
        net.pad_value = 0.0
        self._enc = [enc]  # hidden from nn.Module registration: params live in the fused tensor
        self._net = [net]
        self.n_net = net.params.numel()
        self.params = nn.Parameter(torch.cat([net.params.detach(), enc.params.detach()]))

    def forward(self, x):
        enc, net = self._enc[0], self._net[0]
        flat = self.params.detach()
        feats = T.hash_grid_encode(x.to(torch.float32), flat[self.n_net :], enc.levels, enc.n_features)
        pad = net.dims[0] - feats.shape[-1]
        if pad:
            feats = torch.cat([feats, torch.zeros(feats.shape[0], pad)], dim=-1)
        ws = T.split_mlp_params(flat[: self.n_net], net.dims)
        return T.mlp_forward(feats, ws, net.out_act)[:, : self.n_output_dims].to(torch.float16)
