"""Training-step throughput through the shim: `samnerf_distill` shapes, a reference-sized batch of
4096 rays (train_num_rays_per_batch, samconfigs.py:108) made of 256 4x4 patches, every parameter group trained
(proposal network, nerfacto field, sam_field, conv head), Adam like samconfigs.py:144-161.  Not the headline bench
(bench.py); prints one JSON line with ms per phase so that the simple first-version backward kernels can be judged and
then replaced one at a time.  Needs a B200.

    python tools/bench_train.py [--rays 4096] [--steps 10] [--tiny]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from samnerf_b200 import SAMNeRFConfig, make_synthetic_params  # noqa: E402
from samnerf_b200.nerfstudio_api import RayBundle, SAMModel  # noqa: E402
from samnerf_b200.synthetic import orbit_rays  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rays", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--tiny", action="store_true", help="small hash tables (plumbing check)")
    args = ap.parse_args()
    cfg = SAMNeRFConfig.tiny(clipseg=False, patch_size=4) if args.tiny else SAMNeRFConfig.distill(clipseg=False, patch_size=4)
    params = make_synthetic_params(cfg, "scene", 0)
    m = SAMModel(cfg)
    m.load_state_dict(params)
    m.train()
    groups = m.get_param_groups()
    opt = torch.optim.Adam([{"params": g} for g in groups.values()], lr=5e-4, eps=1e-15)
    # 4x4 patches of adjacent pixels (training patches, pixel_samplers.py:279-294) spread over the 800x800 frame
    o, d = orbit_rays(800, 800, 800.0)
    n_patches = args.rays // 16
    g = torch.Generator().manual_seed(0)
    py = torch.randint(0, 800 - 4, (n_patches,), generator=g)
    px = torch.randint(0, 800 - 4, (n_patches,), generator=g)
    yy = (py[:, None, None] + torch.arange(4)[None, :, None]).expand(-1, 4, 4).reshape(-1)
    xx = (px[:, None, None] + torch.arange(4)[None, None, :]).expand(-1, 4, 4).reshape(-1)
    bundle = RayBundle(origins=o[yy, xx].cuda(), directions=d[yy, xx].cuda())
    image = torch.rand(args.rays, 3, generator=g).cuda()
    feat = (torch.randn(n_patches, 256, generator=g) * 0.1).cuda()

    def ev():
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    tot = {"forward": 0.0, "backward": 0.0, "optimizer": 0.0}
    losses = []
    for step in range(args.warmup + args.steps):
        opt.zero_grad(set_to_none=False)
        e0 = ev()
        m.before_train_iteration(step)
        out = m(bundle, get_feature=["sam"])
        batch = {"image": image, "sam": feat}
        loss = sum(m.get_loss_dict(out, batch, m.get_metrics_dict(out, batch)).values())  # rgb + interlevel + distortion + sam
        e1 = ev()
        loss.backward()
        e2 = ev()
        opt.step()
        m.after_train_iteration(step)
        e3 = ev()
        torch.cuda.synchronize()
        losses.append(float(loss.detach()))
        if step >= args.warmup:
            tot["forward"] += e0.elapsed_time(e1)
            tot["backward"] += e1.elapsed_time(e2)
            tot["optimizer"] += e2.elapsed_time(e3)
    ms = {k: v / args.steps for k, v in tot.items()}
    step_ms = sum(ms.values())
    print(json.dumps({
        "metric": "training rays/s through the shim (all parameter groups, simple first-version backward kernels)",
        "value": args.rays / step_ms * 1e3, "unit": "rays/s", "rays_per_batch": args.rays, "ms_per_step": step_ms,
        "ms": ms, "loss_first": losses[0], "loss_last": losses[-1], "config": "tiny" if args.tiny else "samnerf_distill p=4",
        "note": "the optimizer step includes the re-upload (fp32 -> packed fp16) of every changed table on the next forward",
    }))


if __name__ == "__main__":
    main()
