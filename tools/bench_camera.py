"""What fusing ray generation in front buys: one 800x800 RGB + SAM frame rendered (a) from rays that
start in pinned host memory (15.4 MB of H2D per frame, the bench.py e2e path without the D2H half) and (b) from the
104-byte camera through snrf_render_camera; plus the ray-generation kernel on its own against its 24 B/ray roofline.
Device-resident outputs in both cases, CUDA-event timed, L2 flushed between frames.  Needs a B200.

    python tools/bench_camera.py [--steps 10]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from samnerf_b200 import SAMNeRFConfig, make_synthetic_params  # noqa: E402
from samnerf_b200.renderer import Camera, Renderer  # noqa: E402
from samnerf_b200.synthetic import look_at, orbit_rays  # noqa: E402


def timed(fn, steps, warmup=3):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(steps):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        tot += s.elapsed_time(e)
    return tot / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    args = ap.parse_args()
    cfg = SAMNeRFConfig.distill(clipseg=False, patch_size=1)
    r = Renderer(cfg)
    r.load_params(make_synthetic_params(cfg, "scene", 0))
    H = W = 800
    cam = Camera(800.0, 800.0, W / 2.0, H / 2.0, W, H, look_at((1.2, 0.0, 0.4))[:3, :4])
    o, d = orbit_rays(H, W, 800.0)
    o_host, d_host = o.reshape(-1, 3).contiguous().pin_memory(), d.reshape(-1, 3).contiguous().pin_memory()
    n = H * W
    out = {"rgb": torch.empty(n, 3, device="cuda"), "depth": torch.empty(n, 1, device="cuda"),
           "accumulation": torch.empty(n, 1, device="cuda"), "prop_depth_0": torch.empty(n, 1, device="cuda"),
           "sam": torch.empty(n, cfg.sam_out, device="cuda")}
    o_dev, d_dev = torch.empty(n, 3, device="cuda"), torch.empty(n, 3, device="cuda")

    def from_host_rays():
        o_dev.copy_(o_host, non_blocking=True)
        d_dev.copy_(d_host, non_blocking=True)
        r.render_frame(o_dev, d_dev, get_feature=("sam",), out=out)

    def from_camera():
        r.render_camera(cam, get_feature=("sam",), out=out)

    a = timed(from_host_rays, args.steps)
    ref = out["sam"].clone()
    b = timed(from_camera, args.steps)
    same = bool(torch.equal(torch.nan_to_num(ref), torch.nan_to_num(out["sam"])))
    g = timed(lambda: r.generate_rays(cam), args.steps)
    print(json.dumps({
        "frame_from_pinned_host_rays_ms": a, "frame_from_camera_ms": b, "h2d_bytes_saved_per_frame": 2 * n * 12,
        "same_features": same, "raygen_kernel_ms": g, "raygen_GBps_written": n * 24 / (g * 1e-3) / 1e9,
        "note": "orbit camera of bench.py; features identical only if the camera path reproduces synthetic.orbit_rays bit for bit",
    }))


if __name__ == "__main__":
    main()
