"""Host-side driver of ``libsnrf``: owns one library context, uploads parameters, launches renders.

PyTorch is used for device memory and streams only; every computation happens inside the CUDA library.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Sequence

import torch

from . import _lib as L
from .config import GridConfig, SAMNeRFConfig


def grid_desc(g: GridConfig) -> L.GridDesc:
    """Level table computed on the host exactly as the oracle computes it (``GridConfig.levels``)."""
    d = L.GridDesc()
    d.n_levels, d.n_features = g.n_levels, g.n_features
    for i, (scale, res, offset, size, hashed) in enumerate(g.levels()):
        d.lv[i].scale, d.lv[i].res, d.lv[i].size, d.lv[i].offset, d.lv[i].hashed = scale, res, size, offset, int(hashed)
    return d


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


class Camera:
    """One camera of the reference's ``Cameras`` dataclass (nerfstudio/cameras/cameras.py:61-140): intrinsics,
    ``camera_to_worlds`` 3x4, ``CameraType`` value (1 perspective, 2 fisheye, 3 equirectangular) and the optional
    six distortion parameters ``[k1 k2 k3 k4 p1 p2]``."""

    def __init__(self, fx: float, fy: float, cx: float, cy: float, width: int, height: int, camera_to_world,
                 camera_type: int = 1, distortion_params=None, aabb=None):
        self.fx, self.fy, self.cx, self.cy = float(fx), float(fy), float(cx), float(cy)
        self.width, self.height, self.camera_type = int(width), int(height), int(camera_type)
        self.camera_to_world = torch.as_tensor(camera_to_world, dtype=torch.float32)[:3, :4].contiguous().cpu()
        self.distortion_params = None if distortion_params is None else [float(v) for v in distortion_params]
        # viewer crop box ``[x_min y_min z_min x_max y_max z_max]`` (generate_rays(aabb_box=...), cameras.py:463-482)
        self.aabb = None if aabb is None else [float(v) for v in torch.as_tensor(aabb).flatten().tolist()]

    def as_struct(self) -> L.Camera:
        c = L.Camera()
        c.fx, c.fy, c.cx, c.cy, c.width, c.height = self.fx, self.fy, self.cx, self.cy, self.width, self.height
        c.camera_type = self.camera_type
        c.has_distortion = int(self.distortion_params is not None)
        for i, v in enumerate(self.distortion_params or [0.0] * 6):
            c.distortion[i] = v
        for i, v in enumerate(self.camera_to_world.flatten().tolist()):
            c.c2w[i] = v
        c.has_aabb = int(self.aabb is not None)
        for i, v in enumerate(self.aabb or [0.0] * 6):
            c.aabb[i] = v
        return c


def _index_list(idx, n_default: int):
    """(ctypes int32 array or None, length) for an optional list of pixel rows / columns."""
    if idx is None:
        return None, n_default
    vals = [int(v) for v in torch.as_tensor(idx).flatten().tolist()]
    return (C.c_int32 * max(len(vals), 1))(*vals), len(vals)


class Renderer:
    """One ``snrf_ctx`` on one CUDA device."""

    def __init__(self, cfg: SAMNeRFConfig, device: int = 0, engine: str = "tcgen05"):
        if not torch.cuda.is_available():
            raise RuntimeError("samnerf_b200 needs a CUDA device (sm_100a); there is no CPU path")
        if cfg.num_proposal_samples != 64 or cfg.num_nerf_samples != 32:
            raise ValueError("kernels are built for 64 proposal / 32 nerf samples per ray (samconfigs.py:84-87)")
        self.cfg = cfg
        self.device = torch.device("cuda", device)
        self.lib = L.load()
        h = C.c_void_p()
        rc = self.lib.snrf_ctx_create(device, C.byref(h))
        if rc != 0:
            raise RuntimeError(f"snrf_ctx_create({device}) failed with {rc}")
        self.h = h
        self.set_engine(engine)
        # share the bits of the eval-mode PDF sample positions with torch (ray_samplers.py:325-327)
        nb = cfg.num_nerf_samples + 1
        base = torch.linspace(0.0, 1.0 - (1.0 / nb), steps=nb)
        u = torch.cat([base + 1.0 / (2 * nb), base]).contiguous()  # eval positions, then the training-mode base (:314-322)
        self._check(self.lib.snrf_set_pdf_u(self.h, u.data_ptr(), 2 * nb))
        self.have_sam = self.have_clipseg = self.have_conv = False
        self.feature_dtype = torch.float32
        # opt-in knobs from the environment, so that a whole test / bench run can be put through them unchanged
        # (e.g. SNRF_FEATURE_CUTOFF=5.96e-8 pytest -m gpu runs every parity test through the bucketed feature kernel)
        import os

        if os.environ.get("SNRF_FEATURE_CUTOFF"):
            self.set_feature_cutoff(float(os.environ["SNRF_FEATURE_CUTOFF"]))
        if os.environ.get("SNRF_EARLY_TERMINATION"):
            self.set_early_termination(float(os.environ["SNRF_EARLY_TERMINATION"]))

    # ------------------------------------------------------------------------------------------
    def _check(self, rc: int) -> None:
        if rc != 0:
            msg = self.lib.snrf_last_error(self.h)
            raise RuntimeError(f"libsnrf error {rc}: {msg.decode() if msg else ''}")

    def close(self) -> None:
        if getattr(self, "h", None) is not None and self.h:
            self.lib.snrf_ctx_destroy(self.h)
            self.h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    @property
    def stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    @property
    def launch_count(self) -> int:
        return int(self.lib.snrf_launch_count(self.h))

    REPLICATED = {"sam": 0, "rgb": 1, "depth": 2, "accumulation": 3, "prop_depth_0": 4}

    def set_replication(self, name: str, frame: Optional[torch.Tensor], peer_ptrs: Sequence[int] = (),
                        multicast_ptr: int = 0) -> None:
        """Fused tile all-gather: rows of output ``name`` rendered into (slices of) ``frame`` are also stored at the
        same offset of the other ranks' frame buffers - through ``multicast_ptr`` (NVSwitch multicast alias) or
        ``peer_ptrs`` (peer-mapped aliases on the other ranks).  ``frame=None`` switches it off."""
        which = self.REPLICATED[name]
        if frame is None:
            self._check(self.lib.snrf_set_replication(self.h, which, None, 0, None, None, 0))
            return
        arr = (C.c_void_p * max(len(peer_ptrs), 1))(*[C.c_void_p(p) for p in peer_ptrs])
        self._check(self.lib.snrf_set_replication(
            self.h, which, frame.data_ptr(), frame.numel() * frame.element_size(),
            C.c_void_p(multicast_ptr) if multicast_ptr else None, arr, len(peer_ptrs)))

    def set_replication_mode(self, mode: str) -> None:
        """``stores``: the kernels' own multimem.st / peer stores (default); ``dma``: copy engines move each chunk;
        ``push``: a small kernel on a side stream stores each chunk into the peers' buffers (csrc/exchange.cu)."""
        self._check(self.lib.snrf_set_replication_mode(self.h, {"stores": 0, "dma": 1, "push": 2}[mode]))

    def set_feature_dtype(self, dtype) -> None:
        """Element type of the ``sam`` / ``clipseg`` rows the render calls write: ``torch.float32`` (default, as the
        reference) or ``torch.float16`` (tinycudann's own output precision; halves the bytes of the tile exchange and
        of the device-to-host copy).  See ``snrf_set_feature_dtype``."""
        dtype = {"f32": torch.float32, "f16": torch.float16}.get(dtype, dtype)
        assert dtype in (torch.float32, torch.float16), dtype
        self._check(self.lib.snrf_set_feature_dtype(self.h, int(dtype == torch.float16)))
        self.feature_dtype = dtype

    def set_march_first(self, enable: bool) -> None:
        """Frame calls march the whole tile in one launch, then run the feature kernels per chunk (``snrf_set_march_first``)."""
        self._check(self.lib.snrf_set_march_first(self.h, int(bool(enable))))

    def set_pipeline(self, mode: int) -> None:
        """Chunk pipelining of ``render_frame`` over the library's internal streams: 0 off, 1 auto (default: only
        when outputs are replicated to other ranks), 2 always."""
        self._check(self.lib.snrf_set_pipeline(self.h, int(mode)))

    def set_timing(self, enable: bool) -> None:
        self._check(self.lib.snrf_set_timing(self.h, int(enable)))

    def kernel_times(self):
        """{kernel: (total ms, launches)} since the previous call, from CUDA events on the launching stream."""
        ms = (C.c_double * 3)()
        cnt = (C.c_int64 * 3)()
        self._check(self.lib.snrf_kernel_times(self.h, ms, cnt))
        return {k: (ms[i], cnt[i]) for i, k in enumerate(("march", "feature", "tapgemm"))}

    def set_early_termination(self, eps: float) -> None:
        """Opt-in: skip the second half of a ray's nerfacto samples once the transmittance is below ``eps``
        (0 = exact, the default).  See ``snrf_set_early_termination`` for the error bounds."""
        self._check(self.lib.snrf_set_early_termination(self.h, float(eps)))

    def set_brick_budget(self, gigabytes: float):
        """HBM budget for the cell-major "brick" copies of the leading proposal / nerfacto grid levels
        (``snrf_set_brick_budget``; default 4 GiB, 0 = off; results are bit-identical with any budget).
        Returns ``(proposal_levels, field_levels)`` bricked."""
        a, b = C.c_int(0), C.c_int(0)
        self._check(self.lib.snrf_set_brick_budget(self.h, int(gigabytes * (1 << 30)), C.byref(a), C.byref(b)))
        return a.value, b.value

    def set_anneal(self, anneal: float) -> None:
        """Proposal-weight annealing exponent for training-mode sampling (``snrf_set_anneal``; 1 = off)."""
        self._check(self.lib.snrf_set_anneal(self.h, float(anneal)))

    def set_feature_cutoff(self, cutoff: float) -> None:
        """Bucketed feature kernel: only the leading slots of a ray whose sharpened weight is >= ``cutoff`` (in buckets
        of 1 / 2 / 4 / 8 / 16) are evaluated.  Library default ``2**-24`` = drop what is below one fp32 ulp of the
        accumulated feature; ``0`` = drop exact zeros only; ``< 0`` = every slot of every ray through the un-bucketed
        kernel.  See ``snrf_set_feature_cutoff``."""
        self._check(self.lib.snrf_set_feature_cutoff(self.h, float(cutoff)))

    def feature_slot_stats(self, reset: bool = True):
        """Rays per bucket (1 / 2 / 4 / 8 / 16 slots) of the bucketed feature kernel since the last reset, and the
        number of (ray, sample) slots that were actually gathered."""
        a = (C.c_int64 * 5)()
        self._check(self.lib.snrf_feature_slot_stats(self.h, a, int(reset)))
        rays = [int(v) for v in a]
        return rays, sum(r << b for b, r in enumerate(rays))

    def set_engine(self, engine: str) -> None:
        """``tcgen05`` (default) or ``mma_sync`` (the recompiled-legacy comparison path)."""
        self.engine = engine
        self._check(self.lib.snrf_set_engine(self.h, {"tcgen05": 1, "mma_sync": 0}[engine]))

    # ------------------------------------------------------------------------------------------
    def load_params(self, params: Dict[str, torch.Tensor]) -> None:
        """Upload a reference-layout ``state_dict`` subset (keys as in ``synthetic.make_synthetic_params``)."""
        cfg, lib, s = self.cfg, self.lib, self.stream

        def flat(name):
            t = params[name].detach().to(torch.float32).contiguous().view(-1)
            return t, t.data_ptr(), t.numel()

        for name in self.DENSITY_PARAMS:
            self.upload_density_params(name, params[name])
        for which, (enc, net, n_out) in enumerate(
            (("sam_field.clip_encs", "sam_field.sam_net", cfg.sam_out),
             ("sam_field.clipseg_encs", "sam_field.clipseg_net", cfg.clipseg_out))
        ):
            if f"{net}.params" not in params:
                continue
            for i, g in enumerate(cfg.sam_grids):
                d = grid_desc(g)
                t, p, n = flat(f"{enc}.{i}.params")
                self._check(lib.snrf_upload_feature_grid(self.h, which, i, p, n, C.byref(d), s))
            t, p, n = flat(f"{net}.params")
            self._check(lib.snrf_upload_feature_net(self.h, which, p, n, n_out, s))
            if which == 0:
                self.have_sam = True
            else:
                self.have_clipseg = True
        if "conv_head.0.weight" in params:
            self.upload_conv_head(*[params[k] for k in ("conv_head.0.weight", "conv_head.0.bias", "conv_head.2.weight",
                                                        "conv_head.2.bias")])

    DENSITY_PARAMS = ("proposal_networks.0.mlp_base.params", "field.mlp_base.params", "field.mlp_head.params")

    def upload_density_params(self, name: str, t: torch.Tensor) -> None:
        """(Re-)upload one of the proposal / nerfacto flat parameter tensors (host or device, fp32 -> packed fp16)."""
        cfg, lib, s = self.cfg, self.lib, self.stream
        t = t.detach().to(torch.float32).contiguous().view(-1)
        if name == "proposal_networks.0.mlp_base.params":
            d = grid_desc(cfg.proposal_grid)
            self._check(lib.snrf_upload_proposal(self.h, t.data_ptr(), t.numel(), C.byref(d), s))
        elif name == "field.mlp_base.params":
            d = grid_desc(cfg.field_grid)
            self._check(lib.snrf_upload_field_base(self.h, t.data_ptr(), t.numel(), C.byref(d), s))
        elif name == "field.mlp_head.params":
            self._check(lib.snrf_upload_field_head(self.h, t.data_ptr(), t.numel(), s))
        else:
            raise KeyError(name)

    def upload_conv_head(self, w0: torch.Tensor, b0: torch.Tensor, w2: torch.Tensor, b2: torch.Tensor) -> None:
        """``conv_head.{0,2}.{weight,bias}`` (sam_model.py:202-208), host or device tensors."""
        ts = [t.detach().to(torch.float32).contiguous() for t in (w0, b0, w2, b2)]
        self._check(self.lib.snrf_upload_conv_head(self.h, *[x.data_ptr() for x in ts], self.stream))
        self.have_conv = True

    # ------------------------------------------------------------------------------------------
    def _opts(self, background=None) -> L.RenderOpts:
        cfg = self.cfg
        o = L.RenderOpts()
        o.near_plane, o.far_plane, o.hist_padding = cfg.near_plane_eval, cfg.far_plane, cfg.histogram_padding
        o.k_sam, o.sharpen, o.patch_size = cfg.num_sam_samples, cfg.sharpening_temperature, cfg.patch_size
        if background is None:
            o.bg_mode = L.BG_LAST_SAMPLE
        else:
            o.bg_mode = L.BG_FIXED
            for i in range(3):
                o.bg[i] = float(background[i])
        return o

    def _prep(self, t: Optional[torch.Tensor], cols: int) -> Optional[torch.Tensor]:
        if t is None:
            return None
        t = t.to(device=self.device, dtype=torch.float32).reshape(-1, cols).contiguous()
        return t

    def render(
        self,
        origins: torch.Tensor,
        directions: torch.Tensor,
        nears: Optional[torch.Tensor] = None,
        fars: Optional[torch.Tensor] = None,
        get_feature: Sequence[str] = ("sam",),
        patch: bool = False,
        fast: bool = False,
        background=None,
        debug: bool = False,
        out: Optional[Dict[str, torch.Tensor]] = None,
        picks: bool = False,
        jitter: Optional[torch.Tensor] = None,
    ) -> Dict[str, torch.Tensor]:
        """One chunk of rays: ``SAMModel.forward`` in eval mode (samnerf/sam_model.py:226-314).
        ``out``: optional preallocated CUDA output tensors (e.g. row slices of frame buffers) to write into."""
        o, d = self._prep(origins, 3), self._prep(directions, 3)
        nr, fr = self._prep(nears, 1), self._prep(fars, 1)
        n = o.shape[0]
        dev = self.device
        given = out if out is not None else {}

        def buf(name, *shape, dtype=torch.float32):
            t = given.get(name)
            if t is None:
                return torch.empty(*shape, device=dev, dtype=dtype)
            assert t.is_cuda and t.is_contiguous() and t.dtype == dtype and tuple(t.shape) == shape, (name, t.dtype, dtype)
            return t

        out = {"rgb": buf("rgb", n, 3), "depth": buf("depth", n, 1)}
        if not fast:
            out["accumulation"] = buf("accumulation", n, 1)
            out["prop_depth_0"] = buf("prop_depth_0", n, 1)
        flags = 0
        cfg = self.cfg
        if cfg.distill_sam and "sam" in get_feature:
            flags |= L.WANT_SAM
            if patch:
                flags |= L.PATCH
                out["sam"] = buf("sam", n // (cfg.patch_size**2), cfg.sam_out)
            else:
                out["sam"] = buf("sam", n, cfg.sam_out, dtype=self.feature_dtype)
        if cfg.distill_sam and cfg.use_clipseg_feature and "clipseg" in get_feature:
            flags |= L.WANT_CLIPSEG
            out["clipseg"] = buf("clipseg", n, cfg.clipseg_out, dtype=torch.float32 if patch else self.feature_dtype)
        dbg = None
        if debug:
            k = cfg.num_sam_samples
            out["_prop_weights"] = torch.empty(n, 64, device=dev)
            out["_edges"] = torch.empty(n, 33, device=dev)
            out["_weights"] = torch.empty(n, 32, device=dev)
            out["_density"] = torch.empty(n, 32, device=dev)
            out["_rgb_samples"] = torch.empty(n, 32, 3, device=dev)
            out["_sam_t"] = torch.empty(n, k, device=dev)
            out["_sam_w"] = torch.empty(n, k, device=dev)
            dbg = L.DebugOut()
            dbg.prop_weights, dbg.edges = out["_prop_weights"].data_ptr(), out["_edges"].data_ptr()
            dbg.weights, dbg.density = out["_weights"].data_ptr(), out["_density"].data_ptr()
            dbg.rgb_samples = out["_rgb_samples"].data_ptr()
            dbg.sam_t, dbg.sam_w = out["_sam_t"].data_ptr(), out["_sam_w"].data_ptr()
            if flags & L.WANT_SAM:
                out["_sam_feat"] = torch.empty(n, k, cfg.sam_in, device=dev, dtype=torch.float16)
                dbg.sam_feat = out["_sam_feat"].data_ptr()
        if picks and not debug:
            # the top-k picks only (what the training side of the feature branch needs): 2 x midpoint t and the
            # sharpened, renormalised weights of the k samples (sam_model.py:244-255)
            k = cfg.num_sam_samples
            out["_sam_t"] = torch.empty(n, k, device=dev)
            out["_sam_w"] = torch.empty(n, k, device=dev)
            dbg = L.DebugOut()
            dbg.sam_t, dbg.sam_w = out["_sam_t"].data_ptr(), out["_sam_w"].data_ptr()
        opts = self._opts(background)
        jit = self._set_jitter(jitter, n)
        rc = self.lib.snrf_render(
            self.h, o.data_ptr(), d.data_ptr(), _ptr(nr), _ptr(fr), n, flags, C.byref(opts),
            out["rgb"].data_ptr(), out["depth"].data_ptr(), _ptr(out.get("accumulation")),
            _ptr(out.get("prop_depth_0")), _ptr(out.get("sam")), _ptr(out.get("clipseg")),
            C.byref(dbg) if dbg is not None else None, self.stream,
        )
        self._check(rc)
        return out

    def render_frame(self, origins, directions, get_feature: Sequence[str] = ("sam",), fast: bool = False,
                     chunk: Optional[int] = None, out: Optional[Dict[str, torch.Tensor]] = None) -> Dict[str, torch.Tensor]:
        """All rays of a frame (or of this rank's tile of it) in ``eval_num_rays_per_chunk`` chunks, every ray with
        its features (SURVEY 8 d config 3), written straight into frame-sized buffers (no concatenation)."""
        o, d = self._prep(origins, 3), self._prep(directions, 3)
        n = o.shape[0]
        chunk = chunk or self.cfg.eval_num_rays_per_chunk
        cfg, dev = self.cfg, self.device
        if out is None:
            out = {"rgb": torch.empty(n, 3, device=dev), "depth": torch.empty(n, 1, device=dev)}
            if not fast:
                out["accumulation"] = torch.empty(n, 1, device=dev)
                out["prop_depth_0"] = torch.empty(n, 1, device=dev)
            if cfg.distill_sam and "sam" in get_feature:
                out["sam"] = torch.empty(n, cfg.sam_out, device=dev, dtype=self.feature_dtype)
            if cfg.distill_sam and cfg.use_clipseg_feature and "clipseg" in get_feature:
                out["clipseg"] = torch.empty(n, cfg.clipseg_out, device=dev, dtype=self.feature_dtype)
        flags = 0
        if "sam" in out:
            flags |= L.WANT_SAM
        if "clipseg" in out:
            flags |= L.WANT_CLIPSEG
        for k, v in out.items():
            want = self.feature_dtype if k in ("sam", "clipseg") else torch.float32
            assert v.is_cuda and v.is_contiguous() and v.dtype == want and v.shape[0] == n, (k, v.dtype, want)
        opts = self._opts(None)
        self._check(self.lib.snrf_render_frame(
            self.h, o.data_ptr(), d.data_ptr(), None, None, n, chunk, flags, C.byref(opts), out["rgb"].data_ptr(),
            out["depth"].data_ptr(), _ptr(out.get("accumulation")), _ptr(out.get("prop_depth_0")), _ptr(out.get("sam")),
            _ptr(out.get("clipseg")), self.stream))
        return out

    # ---- camera in front ----------------------------------------------------------
    def generate_rays(self, cam: Camera, rows=None, cols=None, patch: int = 1, pixel_area: bool = False):
        """``Cameras.generate_rays`` for one camera (cameras.py:312-482,490-726) over the pixel grid ``rows x cols``
        (``None`` = the whole image), row-major or patch-major (``patch`` > 1, sam_model.py:376-379).
        Returns ``origins[n,3], directions[n,3], pixel_area[n,1] | None`` on the device; with a crop box on the camera
        (``cam.aabb``) also ``nears[n,1], fars[n,1]``."""
        r, n_rows = _index_list(rows, cam.height)
        c, n_cols = _index_list(cols, cam.width)
        n = n_rows * n_cols
        o = torch.empty(n, 3, device=self.device)
        d = torch.empty(n, 3, device=self.device)
        a = torch.empty(n, 1, device=self.device) if pixel_area else None
        nr = torch.empty(n, 1, device=self.device) if cam.aabb is not None else None
        fr = torch.empty(n, 1, device=self.device) if cam.aabb is not None else None
        st = cam.as_struct()
        self._check(self.lib.snrf_generate_rays(self.h, C.byref(st), r, n_rows, c, n_cols, int(patch), o.data_ptr(),
                                                d.data_ptr(), _ptr(a), _ptr(nr), _ptr(fr), self.stream))
        return (o, d, a) if cam.aabb is None else (o, d, a, nr, fr)

    def render_camera(self, cam: Camera, rows=None, cols=None, get_feature: Sequence[str] = ("sam",), patch: bool = False,
                      fast: bool = False, chunk: Optional[int] = None,
                      out: Optional[Dict[str, torch.Tensor]] = None) -> Dict[str, torch.Tensor]:
        """One loop of ``SAMModel.get_outputs_for_camera_ray_bundle`` (sam_model.py:354-418) from the camera itself:
        rays are generated on the device and rendered chunk by chunk in the same call."""
        cfg, dev = self.cfg, self.device
        r, n_rows = _index_list(rows, cam.height)
        c, n_cols = _index_list(cols, cam.width)
        n = n_rows * n_cols
        chunk = chunk or cfg.eval_num_rays_per_chunk
        flags = 0
        if out is None:
            out = {"rgb": torch.empty(n, 3, device=dev), "depth": torch.empty(n, 1, device=dev)}
            if not fast:
                out["accumulation"] = torch.empty(n, 1, device=dev)
                out["prop_depth_0"] = torch.empty(n, 1, device=dev)
            fdt = torch.float32 if patch else self.feature_dtype
            if cfg.distill_sam and "sam" in get_feature:
                out["sam"] = torch.empty(n // (cfg.patch_size**2) if patch else n, cfg.sam_out, device=dev, dtype=fdt)
            if cfg.distill_sam and cfg.use_clipseg_feature and "clipseg" in get_feature:
                out["clipseg"] = torch.empty(n, cfg.clipseg_out, device=dev, dtype=fdt)
        if "sam" in out:
            flags |= L.WANT_SAM | (L.PATCH if patch else 0)
        if "clipseg" in out:
            flags |= L.WANT_CLIPSEG
        for k, v in out.items():
            want = (torch.float32 if patch else self.feature_dtype) if k in ("sam", "clipseg") else torch.float32
            assert v.is_cuda and v.is_contiguous() and v.dtype == want, (k, v.dtype, want)
        opts = self._opts(None)
        st = cam.as_struct()
        self._check(self.lib.snrf_render_camera(
            self.h, C.byref(st), r, n_rows, c, n_cols, chunk, flags, C.byref(opts), out["rgb"].data_ptr(),
            out["depth"].data_ptr(), _ptr(out.get("accumulation")), _ptr(out.get("prop_depth_0")), _ptr(out.get("sam")),
            _ptr(out.get("clipseg")), self.stream))
        return out

    # ---- training side of the feature-field branch --------------------------------
    FEATURE_PARAMS = {
        "sam": ("sam_field.sam_net.params", "sam_field.clip_encs.0.params", "sam_field.clip_encs.1.params"),
        "clipseg": ("sam_field.clipseg_net.params", "sam_field.clipseg_encs.0.params", "sam_field.clipseg_encs.1.params"),
    }

    def upload_feature_params(self, which: str, net: Optional[torch.Tensor] = None, grid0: Optional[torch.Tensor] = None,
                              grid1: Optional[torch.Tensor] = None) -> None:
        """Re-upload (fp32 -> packed fp16) whichever of the branch's flat parameter tensors changed, e.g. after an
        optimiser step (host or device tensors)."""
        cfg, w, s = self.cfg, {"sam": 0, "clipseg": 1}[which], self.stream
        for i, t in enumerate((grid0, grid1)):
            if t is not None:
                t = t.detach().to(torch.float32).contiguous().view(-1)
                d = grid_desc(cfg.sam_grids[i])
                self._check(self.lib.snrf_upload_feature_grid(self.h, w, i, t.data_ptr(), t.numel(), C.byref(d), s))
        if net is not None:
            t = net.detach().to(torch.float32).contiguous().view(-1)
            n_out = cfg.sam_out if which == "sam" else cfg.clipseg_out
            self._check(self.lib.snrf_upload_feature_net(self.h, w, t.data_ptr(), t.numel(), n_out, s))

    def feature_forward(self, which: str, origins, directions, sam_t, sam_w, save_for_backward: bool = True):
        """``MeanRenderer(SAMField.get_outputs(picked samples))`` from the picks of a render (``_sam_t`` / ``_sam_w``
        of ``render(debug=True)``): returns ``out[N,n_out]`` and the fp16 encoder outputs ``[N,16,192]`` the
        backward pass needs (``None`` when not saved)."""
        o, d = self._prep(origins, 3), self._prep(directions, 3)
        t, w = self._prep(sam_t, 16), self._prep(sam_w, 16)
        n = o.shape[0]
        n_out = self.cfg.sam_out if which == "sam" else self.cfg.clipseg_out
        out = torch.empty(n, n_out, device=self.device)
        enc = torch.empty(n, 16, self.cfg.sam_in, device=self.device, dtype=torch.float16) if save_for_backward else None
        self._check(self.lib.snrf_feature_forward(self.h, {"sam": 0, "clipseg": 1}[which], o.data_ptr(), d.data_ptr(),
                                                  t.data_ptr(), w.data_ptr(), n, out.data_ptr(), _ptr(enc), self.stream))
        return out, enc

    def feature_backward(self, which: str, origins, directions, sam_t, sam_w, enc: torch.Tensor, d_out: torch.Tensor,
                         grads: Optional[Dict[str, torch.Tensor]] = None, want: Sequence[str] = ("net", "grid0", "grid1")):
        """Gradients of the branch's flat fp32 parameters, accumulated (+=) into ``grads`` (``net`` / ``grid0`` /
        ``grid1``; missing ones are created as zeros).  Only the entries in ``want`` are computed and returned."""
        cfg = self.cfg
        o, d = self._prep(origins, 3), self._prep(directions, 3)
        t, w = self._prep(sam_t, 16), self._prep(sam_w, 16)
        n = o.shape[0]
        n_out = cfg.sam_out if which == "sam" else cfg.clipseg_out
        g = d_out.to(device=self.device, dtype=torch.float32).reshape(n, n_out).contiguous()
        assert enc.is_cuda and enc.dtype == torch.float16 and enc.is_contiguous() and enc.numel() == n * 16 * cfg.sam_in
        sizes = {"net": cfg.sam_hidden * cfg.sam_in + n_out * cfg.sam_hidden,
                 "grid0": cfg.sam_grids[0].n_params, "grid1": cfg.sam_grids[1].n_params}
        grads = {} if grads is None else grads
        res = {}
        for k in want:
            if k not in grads:
                grads[k] = torch.zeros(sizes[k], device=self.device)
            gk = grads[k]
            assert gk.is_cuda and gk.dtype == torch.float32 and gk.is_contiguous() and gk.numel() == sizes[k], k
            res[k] = gk
        self._check(self.lib.snrf_feature_backward(
            self.h, {"sam": 0, "clipseg": 1}[which], o.data_ptr(), d.data_ptr(), t.data_ptr(), w.data_ptr(), n,
            g.data_ptr(), enc.data_ptr(), _ptr(res.get("net")), _ptr(res.get("grid0")), _ptr(res.get("grid1")), self.stream))
        return res

    CONV_PARAMS = ("conv_head.0.weight", "conv_head.0.bias", "conv_head.2.weight", "conv_head.2.bias")

    def patch_aggregate_backward(self, feat: torch.Tensor, d_out: torch.Tensor,
                                 grads: Optional[Dict[str, torch.Tensor]] = None, want_d_feat: bool = True):
        """Backward of ``patch_aggregate``: returns ``(grads, d_feat)`` with ``grads`` keyed like ``CONV_PARAMS``
        (torch Conv2d layouts, accumulated into the tensors passed in) and ``d_feat[P*16,256]``."""
        p = self.cfg.patch_size
        f = feat.to(device=self.device, dtype=torch.float32).contiguous()
        n_patches = f.shape[0] // (p * p)
        g = d_out.to(device=self.device, dtype=torch.float32).reshape(n_patches, f.shape[1]).contiguous()
        k = self.cfg.kernel_size
        shapes = {"conv_head.0.weight": (256, 256, k, k), "conv_head.0.bias": (256,),
                  "conv_head.2.weight": (256, 256, k, k), "conv_head.2.bias": (256,)}
        grads = {} if grads is None else grads
        for name, shp in shapes.items():
            if name not in grads:
                grads[name] = torch.zeros(*shp, device=self.device)
            t = grads[name]
            assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and tuple(t.shape) == shp, name
        d_feat = torch.empty_like(f) if want_d_feat else None
        self._check(self.lib.snrf_patch_aggregate_backward(
            self.h, f.data_ptr(), n_patches, p, g.data_ptr(), *[grads[n].data_ptr() for n in self.CONV_PARAMS],
            _ptr(d_feat), self.stream))
        return grads, d_feat

    def field_backward(self, which: str, positions: torch.Tensor, directions: Optional[torch.Tensor] = None,
                       d_density: Optional[torch.Tensor] = None, d_rgb: Optional[torch.Tensor] = None,
                       grads: Optional[Dict[str, torch.Tensor]] = None) -> Dict[str, torch.Tensor]:
        """Backward of ``query_density`` / ``query_rgb``: flat fp32 gradients ``base`` (and ``head`` with ``d_rgb``)
        of the ``proposal`` or nerfacto ``field`` parameters, accumulated (+=) into ``grads``."""
        cfg = self.cfg
        x = self._prep(positions, 3)
        n = x.shape[0]
        prep = lambda t, c: None if t is None else t.to(device=self.device, dtype=torch.float32).reshape(n, c).contiguous()
        d = None if directions is None else self._prep(directions.expand(*positions.shape[:-1], 3), 3)
        gd, gr = prep(d_density, 1), prep(d_rgb, 3)
        if which == "proposal":
            sizes = {"base": cfg.proposal_mlp_params + cfg.proposal_grid.n_params}
        else:
            sizes = {"base": cfg.field_mlp_params + cfg.field_grid.n_params, "head": cfg.head_mlp_params}
        grads = {} if grads is None else grads
        for k, sz in sizes.items():
            if k == "head" and gr is None:
                continue
            if k not in grads:
                grads[k] = torch.zeros(sz, device=self.device)
            assert grads[k].is_cuda and grads[k].dtype == torch.float32 and grads[k].is_contiguous() and grads[k].numel() == sz, k
        self._check(self.lib.snrf_field_backward(
            self.h, 0 if which == "proposal" else 1, x.data_ptr(), _ptr(d), n, _ptr(gd), _ptr(gr),
            grads["base"].data_ptr(), _ptr(grads.get("head")), self.stream))
        return grads

    def pick_samples(self, weights: torch.Tensor, starts: torch.Tensor, ends: torch.Tensor):
        """Top-k + sharpen of the feature samples (sam_model.py:244-255) from ``weights / starts / ends [N,S(,1)]``:
        returns ``(sam_t[N,k], sam_w[N,k])`` in the layout ``feature_forward`` takes."""
        cfg = self.cfg
        n, s = weights.shape[0], weights.shape[1]
        w, a, b = (t.detach().to(device=self.device, dtype=torch.float32).reshape(n, s).contiguous() for t in (weights, starts, ends))
        k = cfg.num_sam_samples
        sam_t, sam_w = torch.empty(n, k, device=self.device), torch.empty(n, k, device=self.device)
        self._check(self.lib.snrf_pick_samples(self.h, w.data_ptr(), a.data_ptr(), b.data_ptr(), n, s, k,
                                               float(cfg.sharpening_temperature), sam_t.data_ptr(), sam_w.data_ptr(), self.stream))
        return sam_t, sam_w

    def ray_op_backward(self, mode: int, a, b, g, background=None):
        """Backward of ``ray_op`` mode 0 (get_weights: returns ``d_densities``) or mode 3 (RGB composite: returns
        ``(d_rgb_samples, d_weights)``)."""
        a = a.to(device=self.device, dtype=torch.float32).contiguous()
        b = b.to(device=self.device, dtype=torch.float32).contiguous()
        g = g.to(device=self.device, dtype=torch.float32).contiguous()
        n, s = b.shape[0], b.shape[1]
        out_a = torch.empty_like(a)
        out_b = torch.empty(n, s, device=self.device) if mode == 3 else None
        bg = (C.c_float * 3)(*[float(v) for v in background]) if background is not None else None
        self._check(self.lib.snrf_ray_op_backward(self.h, mode, a.data_ptr(), b.data_ptr(), g.data_ptr(), out_a.data_ptr(),
                                                  _ptr(out_b), n, s, L.BG_FIXED if bg is not None else L.BG_LAST_SAMPLE,
                                                  C.cast(bg, C.c_void_p) if bg is not None else None, self.stream))
        return out_a if mode == 0 else (out_a, out_b)

    def _set_jitter(self, jitter: Optional[torch.Tensor], n: int):
        """Training-mode single-jitter draws ``[n,2]`` for the next render / sample call (``snrf_set_jitter``).
        Returns the device tensor so that the caller keeps it alive across the launch."""
        if jitter is None:
            return None
        j = jitter.to(device=self.device, dtype=torch.float32).reshape(-1, 2).contiguous()
        # the library compares the row count with the ray count of the call that consumes the draws and refuses a
        # mismatch (SNRF_E_INVALID, "snrf_set_jitter was given ..."), so hand over what was really given
        self._check(self.lib.snrf_set_jitter(self.h, j.data_ptr(), j.shape[0]))
        return j

    def sample(self, origins, directions, nears=None, fars=None, jitter: Optional[torch.Tensor] = None):
        """Proposal weights ``[N,64]``, nerf bin edges ``[N,33]`` and proposal median depth ``[N,1]``;
        ``jitter[N,2]`` switches to training-mode stratified sampling (ray_samplers.py:104-112,314-322)."""
        o, d = self._prep(origins, 3), self._prep(directions, 3)
        nr, fr = self._prep(nears, 1), self._prep(fars, 1)
        n = o.shape[0]
        jit = self._set_jitter(jitter, n)  # noqa: F841  (kept alive until the launch is enqueued)
        w0 = torch.empty(n, 64, device=self.device)
        edges = torch.empty(n, 33, device=self.device)
        pd = torch.empty(n, 1, device=self.device)
        opts = self._opts()
        self._check(self.lib.snrf_sample(self.h, o.data_ptr(), d.data_ptr(), _ptr(nr), _ptr(fr), n, C.byref(opts),
                                         w0.data_ptr(), edges.data_ptr(), pd.data_ptr(), self.stream))
        return w0, edges, pd

    # ---- component-level queries ---------------------------------------------------------------
    def query_density(self, which: str, positions: torch.Tensor):
        """``which``: "proposal" or "field".  Returns ``density[...,1]`` and (field only) ``geo[...,15]`` fp16."""
        shp = positions.shape[:-1]
        x = self._prep(positions, 3)
        n = x.shape[0]
        dens = torch.empty(n, device=self.device)
        geo = torch.empty(n, 15, device=self.device, dtype=torch.float16) if which == "field" else None
        self._check(self.lib.snrf_query_density(self.h, 0 if which == "proposal" else 1, x.data_ptr(), n,
                                                dens.data_ptr(), _ptr(geo), self.stream))
        return dens.view(*shp, 1), (geo.view(*shp, 15) if geo is not None else None)

    def query_rgb(self, directions: torch.Tensor, geo: torch.Tensor) -> torch.Tensor:
        shp = geo.shape[:-1]
        d = self._prep(directions.expand(*shp, 3), 3)
        g = geo.to(device=self.device, dtype=torch.float16).reshape(-1, 15).contiguous()
        rgb = torch.empty(d.shape[0], 3, device=self.device)
        self._check(self.lib.snrf_query_rgb(self.h, d.data_ptr(), g.data_ptr(), d.shape[0], rgb.data_ptr(), self.stream))
        return rgb.view(*shp, 3)

    def query_features(self, which: str, positions: torch.Tensor):
        """``SAMField.get_outputs`` per sample: returns ``(hashgrid[...,192] fp16, out[...,n_out] fp32)``."""
        shp = positions.shape[:-1]
        x = self._prep(positions, 3)
        n = x.shape[0]
        n_out = self.cfg.sam_out if which == "sam" else self.cfg.clipseg_out
        hg = torch.empty(n, self.cfg.sam_in, device=self.device, dtype=torch.float16)
        out = torch.empty(n, n_out, device=self.device)
        self._check(self.lib.snrf_query_features(self.h, 0 if which == "sam" else 1, x.data_ptr(), n, hg.data_ptr(),
                                                 out.data_ptr(), self.stream))
        return hg.view(*shp, -1), out.view(*shp, n_out)

    def ray_op(self, mode: int, a, b=None, c=None, n_channels: int = 0, background=None) -> torch.Tensor:
        a = a.to(device=self.device, dtype=torch.float32).contiguous()
        b = None if b is None else b.to(device=self.device, dtype=torch.float32).contiguous()
        c = None if c is None else c.to(device=self.device, dtype=torch.float32).contiguous()
        n, s = a.shape[0], a.shape[1]
        shape = {0: (n, s), 1: (n, 1), 2: (n, 1), 3: (n, 3), 4: (n, n_channels)}[mode]
        out = torch.empty(*shape, device=self.device)
        bg = (C.c_float * 3)(*[float(v) for v in background]) if background is not None else None
        self._check(self.lib.snrf_ray_op(self.h, mode, a.data_ptr(), _ptr(b), _ptr(c), out.data_ptr(), n, s,
                                         n_channels, L.BG_FIXED if bg is not None else L.BG_LAST_SAMPLE,
                                         C.cast(bg, C.c_void_p) if bg is not None else None, self.stream))
        return out

    def patch_aggregate(self, feat: torch.Tensor) -> torch.Tensor:
        """Conv head on patch-major rows ``[P*p*p, 256]`` -> ``[P, 256]`` (samnerf/sam_model.py:260-265)."""
        p = self.cfg.patch_size
        f = feat.to(device=self.device, dtype=torch.float32).contiguous()
        n_patches = f.shape[0] // (p * p)
        out = torch.empty(n_patches, f.shape[1], device=self.device)
        self._check(self.lib.snrf_patch_aggregate(self.h, f.data_ptr(), n_patches, p, out.data_ptr(), self.stream))
        return out
