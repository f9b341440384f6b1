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
    DENSITY_PARAMS = Renderer.DENSITY_PARAMS
    CONV_PARAMS = Renderer.CONV_PARAMS

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

    def set_anneal(self, anneal):
        self.anneal = float(anneal)

    def upload_density_params(self, name, t):
        from oracle.samnerf_oracle import Oracle

        self.p[name] = t.detach().clone().view(-1)
        self.uploads.append(name)
        self.orc = Oracle(self.cfg, self.p)

    # ---- component queries / ray ops: oracle forward, emulated kernel bodies backward ----------------
    def query_density(self, which, positions):
        with torch.no_grad():
            if which == "proposal":
                return self.orc.proposal_density(positions)[..., None], None
            dens, geo = self.orc.field_density(positions)
            return dens[..., None], geo.to(torch.float16)

    def query_rgb(self, directions, geo):
        with torch.no_grad():
            return self.orc.field_rgb(directions.expand(*geo.shape[:-1], 3), geo.float())

    def sample(self, origins, directions, nears=None, fars=None, jitter=None):
        with torch.no_grad():
            res = self.orc.render_rays(self._prep(origins, 3), self._prep(directions, 3), self._prep(nears, 1),
                                       self._prep(fars, 1), get_feature=(), return_intermediates=True, jitter=jitter)
        return res["_w0"], res["_eu1"], res["prop_depth_0"]

    def ray_op(self, mode, a, b=None, c=None, n_channels=0, background=None):
        from oracle.samnerf_oracle import composite_rgb, get_weights, median_depth

        with torch.no_grad():
            if mode == 0:
                return get_weights(a, b)
            if mode == 1:
                return a.sum(-1, keepdim=True)
            if mode == 2:
                return median_depth(a, b, c)
            if mode == 3:
                return composite_rgb(a, b, background)
            return (a * b[..., None]).sum(-2)

    def pick_samples(self, weights, starts, ends):
        from emu.build_emu import load

        n, s = weights.shape[0], weights.shape[1]
        k = self.cfg.num_sam_samples
        arr = lambda t: np.ascontiguousarray(t.detach().reshape(n, s).numpy(), np.float32)
        ptr = lambda x: x.ctypes.data_as(C.c_void_p)
        keep = [arr(weights), arr(starts), arr(ends)]
        sam_t, sam_w = np.zeros((n, k), np.float32), np.zeros((n, k), np.float32)
        load().emu_pick_samples(ptr(keep[0]), ptr(keep[1]), ptr(keep[2]), C.c_longlong(n), s, k,
                                C.c_float(self.cfg.sharpening_temperature), ptr(sam_t), ptr(sam_w))
        return torch.from_numpy(sam_t), torch.from_numpy(sam_w)

    def ray_op_backward(self, mode, a, b, g, background=None):
        from emu.build_emu import load

        arr = lambda t: np.ascontiguousarray(t.detach().numpy(), np.float32)
        ptr = lambda x: None if x is None else x.ctypes.data_as(C.c_void_p)
        n, s = b.shape[0], b.shape[1]
        keep = [arr(a), arr(b), arr(g)]
        if mode == 0:
            out = np.zeros((n, s), np.float32)
            load().emu_weights_backward(ptr(keep[0]), ptr(keep[1]), ptr(keep[2]), ptr(out), C.c_longlong(n), s)
            return torch.from_numpy(out)
        bg = None if background is None else np.asarray([float(v) for v in background], np.float32)
        d_rgb, d_w = np.zeros((n, s, 3), np.float32), np.zeros((n, s), np.float32)
        load().emu_rgb_backward(ptr(keep[0]), ptr(keep[1]), ptr(keep[2]), int(bg is not None), ptr(bg), ptr(d_rgb), ptr(d_w),
                                C.c_longlong(n), s)
        return torch.from_numpy(d_rgb), torch.from_numpy(d_w)

    def field_backward(self, which, positions, directions=None, d_density=None, d_rgb=None, grads=None):
        from emu.build_emu import load
        from oracle import tcnn_spec as T

        cfg, orc = self.cfg, self.orc
        w = 0 if which == "proposal" else 1
        pos = positions.reshape(-1, 3).float()
        n = pos.shape[0]
        dirs = None if directions is None else directions.expand(*positions.shape[:-1], 3).reshape(-1, 3).float()
        with torch.no_grad():  # the activations launch_field_backward recomputes on the device
            x_pts, sel = orc._normalized(pos, float("inf"))
            table, levels, ws = (orc.field_table, orc.field_levels, orc.base_w) if w else (orc.prop_table, orc.prop_levels, orc.prop_w)
            x = T.hash_grid_encode(x_pts, table, levels, 2)
            xin = x if w else torch.cat([x, torch.zeros(n, 6)], -1)
            h1 = T.f16(torch.relu(xin @ T.f16(ws[0]).T))
            o = T.f16(h1 @ T.f16(ws[1]).T)
            acts = dict(x=x, h1=h1, o=o)
            if w and d_rgb is not None:
                hx = torch.cat([T.sh4((dirs + 1.0) / 2.0), o[:, 1:16], torch.ones(n, 1)], -1)
                g1 = T.f16(torch.relu(hx @ T.f16(orc.head_w[0]).T))
                g2 = T.f16(torch.relu(g1 @ T.f16(orc.head_w[1]).T))
                acts.update(hx=hx, g1=g1, g2=g2, pre3=T.f16(g2 @ T.f16(orc.head_w[2]).T))
        grid = cfg.field_grid if w else cfg.proposal_grid
        lv = np.zeros((grid.n_levels, 5), np.float64)
        for l, (scale, res, offset, size, hashed) in enumerate(grid.levels()):
            lv[l] = (scale, res, size, offset, float(hashed))
        n_base = (cfg.field_mlp_params + grid.n_params) if w else (cfg.proposal_mlp_params + grid.n_params)
        g_base, g_head = np.zeros(n_base, np.float32), np.zeros(cfg.head_mlp_params, np.float32)
        arr = lambda t: None if t is None else np.ascontiguousarray(t.detach().reshape(n, -1).numpy(), np.float32)
        bits = lambda t: None if t is None else np.ascontiguousarray(t.detach().to(torch.float16).numpy().view(np.uint16))
        ptr = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
        keep = [arr(pos), arr(d_density), arr(d_rgb), lv, bits(ws[0]), bits(ws[1])] + [bits(m) for m in orc.head_w]
        keep += [bits(acts.get(k)) for k in ("x", "h1", "o", "hx", "g1", "g2", "pre3")] + [arr(sel.float())]
        load().emu_field_backward(w, ptr(keep[0]), C.c_longlong(n), ptr(keep[1]), ptr(keep[2]), ptr(keep[3]), grid.n_levels,
                                  *[ptr(k) for k in keep[4:17]], ptr(g_base), ptr(g_head))
        grads = {} if grads is None else grads
        full = {"base": torch.from_numpy(g_base)}
        if d_rgb is not None:
            full["head"] = torch.from_numpy(g_head)
        for k, v in full.items():
            grads[k] = grads[k] + v if k in grads else v
        return grads

    def patch_aggregate(self, feat):
        with torch.no_grad():
            return self.orc.patch_aggregate(feat)

    def patch_aggregate_backward(self, feat, d_out, grads=None, want_d_feat=True):
        from emu.build_emu import load

        n_patches = feat.shape[0] // 16
        arr = lambda t: np.ascontiguousarray(t.detach().numpy(), np.float32)
        bits = lambda t: np.ascontiguousarray(t.detach().reshape(256, -1).to(torch.float16).numpy().view(np.uint16))
        ptr = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
        w0, b0, w2, b2 = [self.p[n] for n in self.CONV_PARAMS]
        g = [np.zeros(t.numel(), np.float32) for t in (w0, b0, w2, b2)]
        d_feat = np.zeros((feat.shape[0], 256), np.float32) if want_d_feat else None
        keep = [arr(feat), arr(d_out), bits(w0), arr(b0), bits(w2), arr(b2)]
        load().emu_conv_backward(ptr(keep[0]), C.c_longlong(n_patches), ptr(keep[1]), ptr(keep[2]), ptr(keep[3]), ptr(keep[4]),
                                 ptr(keep[5]), ptr(g[0]), ptr(g[1]), ptr(g[2]), ptr(g[3]), ptr(d_feat), None)
        out = {n: torch.from_numpy(x).view(self.p[n].shape) for n, x in zip(self.CONV_PARAMS, g)}
        grads = {} if grads is None else grads
        for n, v in out.items():
            grads[n] = grads[n] + v if n in grads else v
        return grads, (None if d_feat is None else torch.from_numpy(d_feat))

    def upload_conv_head(self, w0, b0, w2, b2):
        from oracle.samnerf_oracle import Oracle

        for k, t in zip(("conv_head.0.weight", "conv_head.0.bias", "conv_head.2.weight", "conv_head.2.bias"), (w0, b0, w2, b2)):
            self.p[k] = t.detach().clone()
        self.uploads.append("conv_head")
        self.orc = Oracle(self.cfg, self.p)

    def render(self, origins, directions, nears=None, fars=None, get_feature=(), patch=False, fast=False, background=None,
               debug=False, out=None, picks=False, jitter=None):
        with torch.no_grad():
            res = self.orc.render_rays(self._prep(origins, 3), self._prep(directions, 3), self._prep(nears, 1),
                                       self._prep(fars, 1), get_feature=tuple(get_feature) or ("sam",), fast=fast,
                                       background=background, return_intermediates=True, jitter=jitter)
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
                                    ptr(g_t[1]), None, C.c_float(-1.0), None)
        full = {"net": torch.from_numpy(np.concatenate([g_w1.ravel(), g_w2.ravel()])), "grid0": torch.from_numpy(g_t[0]),
                "grid1": torch.from_numpy(g_t[1])}
        grads = {} if grads is None else grads
        res = {}
        for k in want:
            grads[k] = grads[k] + full[k] if k in grads else full[k]
            res[k] = grads[k]
        return res
