"""TEST INFRASTRUCTURE ONLY: a CPU stand-in for ``samnerf_b200.renderer.Renderer`` so that the *Python* logic of the
training shim (``SAMModel.train`` / ``_sync_params`` / ``_FeatureBranchFn`` / ``state_dict``) can be exercised in a
container without a GPU.  Forward values come from the oracle, gradients from the host emulation of the backward
kernel bodies (tests/emu).  Never imported by the package."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from samnerf_b200.renderer import Renderer


class FakeRenderer:
    FEATURE_PARAMS = Renderer.FEATURE_PARAMS

    def __init__(self, cfg, device: int = 0, engine: str = "tcgen05"):
        self.cfg, self.device, self.engine = cfg, torch.device("cpu"), engine
        self.uploads = []

    def load_params(self, params):
        from oracle.samnerf_oracle import Oracle

        self.p = {k: v.detach().clone() for k, v in params.items()}
        self.orc = Oracle(self.cfg, self.p)

    def _prep(self, t, cols):
        return None if t is None else t.to(torch.float32).reshape(-1, cols).contiguous()

    def upload_feature_params(self, which, net=None, grid0=None, grid1=None):
        from oracle.samnerf_oracle import Oracle

        names = self.FEATURE_PARAMS[which]
        for name, t in zip(names, (net, grid0, grid1)):
            if t is not None:
                self.p[name] = t.detach().clone()
                self.uploads.append(name)
        self.orc = Oracle(self.cfg, self.p)

    def upload_conv_head(self, w0, b0, w2, b2):
        from oracle.samnerf_oracle import Oracle

        for k, t in zip(("conv_head.0.weight", "conv_head.0.bias", "conv_head.2.weight", "conv_head.2.bias"), (w0, b0, w2, b2)):
            self.p[k] = t.detach().clone()
        self.uploads.append("conv_head")
        self.orc = Oracle(self.cfg, self.p)

    def render(self, origins, directions, nears=None, fars=None, get_feature=(), patch=False, fast=False, background=None,
               debug=False, out=None, picks=False):
        with torch.no_grad():
            res = self.orc.render_rays(self._prep(origins, 3), self._prep(directions, 3), self._prep(nears, 1),
                                       self._prep(fars, 1), get_feature=tuple(get_feature) or ("sam",), fast=fast,
                                       background=background, return_intermediates=True)
        eu = res["_eu1"]
        tm2 = eu[:, :-1] + eu[:, 1:]
        keep = {k: v for k, v in res.items() if not k.startswith("_") and (k in get_feature or k in ("rgb", "depth", "accumulation", "prop_depth_0"))}
        if picks:
            keep["_sam_t"] = torch.gather(tm2, 1, res["_best_ids"])
            keep["_sam_w"] = res["_sam_weights"]
        return keep

    def _enc(self, which, pos):
        from oracle import tcnn_spec as T
        from oracle.samnerf_oracle import contract

        names = self.FEATURE_PARAMS[which]
        pts = (contract(pos.reshape(-1, 3), None) + 2.0) / 4.0
        return torch.cat([T.hash_grid_encode(pts, self.p[names[1 + i]], self.orc.sam_levels[i], 8) for i in range(2)], -1)

    def feature_forward(self, which, origins, directions, sam_t, sam_w, save_for_backward=True):
        with torch.no_grad():
            pos = origins[:, None, :] + directions[:, None, :] * sam_t[..., None] / 2.0
            f = self.orc.sam_field(pos, which=(which,))
            out = (sam_w[..., None] * f[which]).sum(dim=-2)
            enc = self._enc(which, pos).view(origins.shape[0], 16, 192).to(torch.float16) if save_for_backward else None
        return out, enc

    def feature_backward(self, which, origins, directions, sam_t, sam_w, enc, d_out, grads=None, want=("net", "grid0", "grid1")):
        from emu.build_emu import load

        cfg = self.cfg
        names = self.FEATURE_PARAMS[which]
        n, n_out = origins.shape[0], d_out.shape[-1]
        net = self.p[names[0]]
        w1, w2 = net[: 256 * 192].view(256, 192), net[256 * 192:].view(n_out, 256)
        lv = np.zeros((2, 12, 5), np.float64)
        for e, g in enumerate(cfg.sam_grids):
            for l, (scale, res, offset, size, hashed) in enumerate(g.levels()):
                lv[e, l] = (scale, res, size, offset, float(hashed))
        g_w1, g_w2 = np.zeros((256, 192), np.float32), np.zeros((n_out, 256), np.float32)
        g_t = [np.zeros(cfg.sam_grids[i].n_params, np.float32) for i in range(2)]
        arr = lambda t: np.ascontiguousarray(t.detach().numpy(), np.float32)
        bits = lambda t: np.ascontiguousarray(t.detach().to(torch.float16).numpy().view(np.uint16))
        ptr = lambda a: a.ctypes.data_as(C.c_void_p)
        ins = [arr(origins), arr(directions), arr(sam_t), arr(sam_w), arr(d_out), bits(enc), bits(w1), bits(w2), lv]
        load().emu_feature_backward(ptr(ins[0]), ptr(ins[1]), ptr(ins[2]), ptr(ins[3]), C.c_longlong(n), ptr(ins[4]), n_out,
                                    ptr(ins[5]), ptr(ins[6]), ptr(ins[7]), ptr(ins[8]), ptr(g_w1), ptr(g_w2), ptr(g_t[0]),
                                    ptr(g_t[1]), None)
        full = {"net": torch.from_numpy(np.concatenate([g_w1.ravel(), g_w2.ravel()])), "grid0": torch.from_numpy(g_t[0]),
                "grid1": torch.from_numpy(g_t[1])}
        grads = {} if grads is None else grads
        res = {}
        for k in want:
            grads[k] = grads[k] + full[k] if k in grads else full[k]
            res[k] = grads[k]
        return res
