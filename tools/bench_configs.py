"""Throughput of the other BASELINE.json configs (2: RGB-only 800x800, 3: +SAM, 4: 1600x1060 + ClipSeg + patch head),
device-resident, CUDA-event timed, L2 flushed between frames.  Not the headline bench (bench.py); fills BASELINE.md section 5."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from samnerf_b200 import SAMNeRFConfig, make_synthetic_params  # noqa: E402
from samnerf_b200.config import get_feature_size  # noqa: E402
from samnerf_b200.nerfstudio_api import RayBundle, SAMModel  # noqa: E402
from samnerf_b200.renderer import Renderer  # noqa: E402
from samnerf_b200.synthetic import orbit_rays  # noqa: E402


def timed(fn, steps=5, warmup=3):
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
    res = {}
    # config 2 / 3: 800x800, every ray
    cfg = SAMNeRFConfig.distill(clipseg=False, patch_size=1)
    params = make_synthetic_params(cfg, "scene", 0)
    r = Renderer(cfg)
    r.load_params(params)
    o, d = orbit_rays(800, 800, 800.0)
    o, d = o.reshape(-1, 3).cuda(), d.reshape(-1, 3).cuda()
    ms = timed(lambda: r.render_frame(o, d, get_feature=()))
    res["config2_rgb_only_800x800"] = {"rays": 640000, "ms": ms, "mrays_s": 640000 / ms / 1e3}
    ms = timed(lambda: r.render_frame(o, d, get_feature=("sam",)))
    res["config3_rgb_sam_800x800"] = {"rays": 640000, "ms": ms, "mrays_s": 640000 / ms / 1e3}
    r.close()
    del r
    # config 4: the reference-faithful frame at 1600x1060 with ClipSeg and the patch head (loops A, B, C)
    cfg = SAMNeRFConfig.distill(clipseg=True, patch_size=4)
    params = make_synthetic_params(cfg, "scene", 0)
    m = SAMModel(cfg)
    m.load_state_dict(params)
    H, W = 1060, 1600
    o, d = orbit_rays(H, W, 1600.0)
    bundle = RayBundle(origins=o.cuda(), directions=d.cuda(), pixel_area=torch.ones(H, W, 1, device="cuda"),
                       camera_indices=torch.zeros(H, W, 1, dtype=torch.long, device="cuda"))
    fh, fw = get_feature_size(H, W)
    n_rays = H * W + fh * 4 * fw * 4 + 1024
    ms = timed(lambda: m.get_outputs_for_camera_ray_bundle(bundle), steps=3, warmup=2)
    res["config4_1600x1060_clipseg_patch"] = {"rays": n_rays, "ms": ms, "mrays_s": n_rays / ms / 1e3,
                                              "sam_map": [fh, fw, 256], "clipseg_map": [32, 32, 192]}
    # patch aggregation kernel alone on [2752*16, 256]
    feat = torch.randn(fh * fw * 16, 256, device="cuda") * 0.3
    ms = timed(lambda: m.renderer.patch_aggregate(feat))
    res["patch_head_2752_patches"] = {"patches": fh * fw, "ms": ms, "tflops": fh * fw * 37.75e6 / ms / 1e9}
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
