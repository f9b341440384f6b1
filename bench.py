#!/usr/bin/env python
"""Headline benchmark: Mrays/s rendering RGB + 256-d SAM features at 800x800 (BASELINE.json), 1/2/4/8 B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--config sam|rgb|clipseg_patch]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \\
        bench.py --gpus N --steps K --warmup W

One "step" = one frame.  Default `--config sam` is BASELINE.json configs[2] (the configuration the metric is quoted
on; SURVEY.md 8 d config 3): the `samnerf_distill` model, 800x800 = 640 000 rays, every ray rendered to rgb / median
depth / accumulation / proposal depth AND the 256-d SAM feature (k = 16 top samples, sharpening 10, patch 1) through
`libsnrf` (C ABI).  `--config rgb` is configs[1] (RGB-only), `--config clipseg_patch` is configs[3] (1600x1060 frame +
patch-aggregated SAM map + ClipSeg map through SAMModel.get_outputs_for_camera_ray_bundle).  With N > 1 the frame is cut
into N contiguous row blocks ("screen tiles"), one per rank, and the rendered tiles are exchanged into every rank's
frame buffer (symmetric memory): by copy engines per chunk, by the kernels' own multicast / peer stores, or by an NCCL
all-gather (--gather).  Total work is fixed, so scaling is "strong".  Synthetic scene-like parameters (seed 0),
synthetic orbit camera.

`--impl reference` times the reference's own CPU path: the reference is Python + the CUDA-only tinycudann, so its CPU
implementation is the oracle port (oracle/samnerf_oracle.py, validated against the reference's own modules by
oracle/make_golden.py), all host threads, on a bounded sample of the same frame per step.  `cpu_baseline` of the
native line is the same code on the same sample definition.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

# algorithmic bytes per ray of each kernel (SURVEY.md 8 d / BASELINE.md section 3):
#   march:   64 proposal samples x 5 levels x 8 corners x 4 B + 32 nerf samples x 16 levels x 8 corners x 4 B
#   feature: 16 picked samples x 24 levels x 8 corners x 16 B (every slot), or 3 072 B per evaluated slot (bucketed)
#   tapgemm: 512 B of fp16 hidden sums read + 1 024 B of fp32 features written per ray
BYTES_PER_RAY = {"march": 10240 + 16384, "feature": 49152, "tapgemm": 1536}
CPU_SAMPLE_RAYS = 16384  # one definition for `cpu_baseline` and `--impl reference`
# HBM the bench spends on cell-major brick copies of the nerfacto / proposal grid levels (GiB).  Measured on the headline
# frame (march kernel, ms / frame, gpurun call 14): 10 GiB 4.29 -> 24 GiB 4.17 -> 62 GiB 4.08; a B200 has 180 GB and the
# whole model is 160 MB, so the bench takes the fastest setting.  Results are bit-identical at every budget.
BRICK_GB_DEFAULT = 62.0

CONFIGS = {
    "sam": dict(metric="Mrays/s rendering RGB+256-d SAM features at 800x800", H=800, W=800, focal=800.0,
                workload="samnerf_distill 800x800 RGB+256-d SAM (k=16, p=1): 640000 rays", features=("sam",)),
    "rgb": dict(metric="Mrays/s rendering RGB only at 800x800 (BASELINE.json configs[1])", H=800, W=800, focal=800.0,
                workload="samnerf_no_distill 800x800 RGB-only volumetric render: 640000 rays", features=()),
    "clipseg_patch": dict(metric="Mrays/s rendering the 1600x1060 frame + patch-aggregated SAM map + ClipSeg map (BASELINE.json configs[3])",
                          H=1060, W=1600, focal=1600.0,
                          workload="samnerf_distill + ClipSeg head, 1600x1060 through get_outputs_for_camera_ray_bundle "
                                   "(loops A/B/C: 1696000 + 44032 + 1024 rays, p=4)", features=("sam", "clipseg")),
}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def source_digest() -> str:
    """Digest of the CUDA sources: ties committed ncu captures to the build being benchmarked."""
    h = hashlib.sha256()
    d = os.path.join(ROOT, "segment-anything-in-nerf_b200", "csrc")
    for name in sorted(os.listdir(d)):
        if name.endswith((".cu", ".cuh")):
            with open(os.path.join(d, name), "rb") as f:
                h.update(name.encode() + b"\0" + f.read())
    return h.hexdigest()[:16]


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of each hot kernel from the newest committed ncu
    capture (profiles/r*_traffic.json, written by tools/ncu_traffic.py from `ncu --set full` reports of bench.py).
    Returns ({kernel: bytes per ray}, note)."""
    d = os.path.join(ROOT, "profiles")
    files = sorted(f for f in os.listdir(d) if f.endswith("_traffic.json")) if os.path.isdir(d) else []
    if not files:
        return {}, "no ncu capture committed"
    with open(os.path.join(d, files[-1])) as f:
        t = json.load(f)
    per_ray = {}
    for k in ("march", "feature", "tapgemm"):
        e = t.get(k) or t.get(k + "_kernel") or (t.get("sam_kernel") if k == "feature" else None)
        if e:
            per_ray[k] = (e["dram_bytes_read"] + e["dram_bytes_write"]) / e["rays_per_launch"]
    same = t.get("source_digest") == source_digest()
    note = (f"profiles/{files[-1]}: ncu --set full, dram bytes per launch of bench.py's workload scaled to this run's rays per launch; "
            + ("captured from exactly these CUDA sources" if same else
               f"STALE: captured from sources {t.get('source_digest', '(round 1, unrecorded)')}, this build is {source_digest()}"))
    return per_ray, note


class ClockSampler:
    """nvidia-smi clock / throttle-reason samples taken DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, reasons, mx = [], set(), None
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    mx = float(f[2])
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm))
        return out


def frame_rays(conf):
    from samnerf_b200.synthetic import orbit_rays

    o, d = orbit_rays(conf["H"], conf["W"], conf["focal"])
    return o.reshape(-1, 3).contiguous(), d.reshape(-1, 3).contiguous()


def model_config(name):
    from samnerf_b200 import SAMNeRFConfig

    if name == "clipseg_patch":
        return SAMNeRFConfig.distill(clipseg=True, patch_size=4)
    return SAMNeRFConfig.distill(clipseg=False, patch_size=1)


# ---- the CPU arm: one sample definition for `cpu_baseline` (native line) and `--impl reference` -----------------------
def cpu_sample(o, d, step: int, n: int = CPU_SAMPLE_RAYS):
    """`n` rays strided over the frame, shifted by `step` so that successive steps see different rays."""
    idx = ((torch.arange(n) * (o.shape[0] // n)) + step) % o.shape[0]
    return o[idx].contiguous(), d[idx].contiguous()


CPU_SAMPLE_TEXT = (f"{CPU_SAMPLE_RAYS} rays strided over the frame per step (bounded sample of the frame; the stride start moves by one "
                   "ray per step), oracle port in torch CPU fp32 on all host threads, mean over the timed steps")


def cpu_rate(conf_name: str, steps: int, warmup: int, regime: str):
    """Mrays/s of the oracle port.  Returns (value, ms per step, cores)."""
    from oracle.samnerf_oracle import Oracle
    from samnerf_b200 import make_synthetic_params

    conf = CONFIGS[conf_name]
    cfg = model_config(conf_name)
    params = make_synthetic_params(cfg, regime, 0)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    orc = Oracle(cfg, params)
    o, d = frame_rays(conf)
    feats = tuple(conf["features"])
    times = []
    for step in range(warmup + steps):
        oo, dd = cpu_sample(o, d, step)
        t0 = time.perf_counter()
        with torch.no_grad():
            orc.render_rays(oo, dd, get_feature=feats)
        if step >= warmup:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    return CPU_SAMPLE_RAYS * len(times) / total / 1e6, 1e3 * total / len(times), cores


def run_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    conf = CONFIGS[args.config]
    value, ms, cores = cpu_rate(args.config, args.steps, args.warmup, args.regime)
    line = {
        "impl": "reference", "metric": conf["metric"], "value": value, "unit": "Mrays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": conf["workload"], "regime": args.regime,
                   "note": "reference CPU path = oracle port (the reference's tinycudann dependency is CUDA-only)"},
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": cores, "kind": "port", "sample": CPU_SAMPLE_TEXT},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---- config 4: the reference-faithful whole-image call through the model shim -----------------------------------------
def run_clipseg_patch(args, dev):
    from samnerf_b200 import make_synthetic_params
    from samnerf_b200.config import get_feature_size
    from samnerf_b200.nerfstudio_api import RayBundle, SAMModel

    conf = CONFIGS["clipseg_patch"]
    cfg = model_config("clipseg_patch")
    m = SAMModel(cfg)
    m.load_state_dict(make_synthetic_params(cfg, args.regime, 0))
    m.renderer.set_brick_budget(args.brick_gb if args.brick_gb >= 0 else BRICK_GB_DEFAULT)
    H, W = conf["H"], conf["W"]
    o, d = frame_rays(conf)
    o_host, d_host = o.view(H, W, 3).pin_memory(), d.view(H, W, 3).pin_memory()
    area = torch.ones(H, W, 1, device=dev)
    cam = torch.zeros(H, W, 1, dtype=torch.long, device=dev)
    bundle = RayBundle(origins=o_host.to(dev), directions=d_host.to(dev), pixel_area=area, camera_indices=cam)
    fh, fw = get_feature_size(H, W)
    n_rays = H * W + fh * 4 * fw * 4 + 1024
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(max(args.warmup, 3)):
        outs = m.get_outputs_for_camera_ray_bundle(bundle)
    torch.cuda.synchronize()
    launches0 = m.renderer.launch_count
    clocks = ClockSampler(dev.index)
    clocks.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for s, e in ev:
        flush.zero_()
        s.record()
        outs = m.get_outputs_for_camera_ray_bundle(bundle)
        e.record()
    torch.cuda.synchronize()
    launches = m.renderer.launch_count - launches0
    ms = sum(s.elapsed_time(e) for s, e in ev) / args.steps
    # end to end: pinned host rays in, every output tensor of the call back on the host, inside the timed region
    keys = [k for k, v in outs.items() if torch.is_tensor(v)]
    host = {k: torch.empty(outs[k].shape, dtype=outs[k].dtype).pin_memory() for k in keys}
    e2e_steps = max(3, min(args.steps, 5))
    ev2 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(e2e_steps)]
    for i in range(2 + e2e_steps):
        if i >= 2:
            ev2[i - 2][0].record()
        b = RayBundle(origins=o_host.to(dev, non_blocking=True), directions=d_host.to(dev, non_blocking=True),
                      pixel_area=area, camera_indices=cam)
        res = m.get_outputs_for_camera_ray_bundle(b)
        for k in keys:
            host[k].copy_(res[k], non_blocking=True)
        if i >= 2:
            ev2[i - 2][1].record()
    torch.cuda.synchronize()
    e2e_ms = sum(s.elapsed_time(e) for s, e in ev2) / e2e_steps
    clk = clocks.stop()
    line = {
        "metric": conf["metric"], "value": n_rays / ms / 1e3, "unit": "Mrays/s", "n_gpus": 1, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f16 operands / f32 accumulate", "data": "synthetic",
        "config": {"workload": conf["workload"], "regime": args.regime, "rays_per_frame": n_rays,
                   "sam_map": [fh, fw, 256], "clipseg_map": [32, 32, 192],
                   "l2": "256 MiB buffer zeroed between timed frames (untimed)"},
        "roofline": None,
        "e2e": {"value": n_rays / e2e_ms / 1e3, "unit": "Mrays/s", "h2d_bytes_per_step": 2 * H * W * 3 * 4,
                "d2h_bytes_per_step": int(sum(host[k].numel() * host[k].element_size() for k in keys)), "steps": e2e_steps,
                "api": "SAMModel.get_outputs_for_camera_ray_bundle, pinned host ray bundle in, every output tensor copied back"},
        "gpu_launches": launches, "clocks": clk,
    }
    if not args.no_cpu_baseline:
        v, _, cores = cpu_rate("clipseg_patch", 2, 1, args.regime)
        line["cpu_baseline"] = {"value": v, "unit": "Mrays/s", "cores": cores, "kind": "port", "sample": CPU_SAMPLE_TEXT}
    print(json.dumps(line), flush=True)


def run_native(args):
    import torch.distributed as dist

    from samnerf_b200 import make_synthetic_params
    from samnerf_b200.renderer import Renderer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit("--gpus N > 1 must be launched with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if args.config == "clipseg_patch":
        if world > 1:
            raise SystemExit("--config clipseg_patch is a single-GPU configuration (BASELINE.json configs[3])")
        return run_clipseg_patch(args, dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    conf = CONFIGS[args.config]
    H, W = conf["H"], conf["W"]
    feats = conf["features"]
    cfg = model_config(args.config)
    params = make_synthetic_params(cfg, args.regime, 0)
    r = Renderer(cfg, device=local, engine=args.engine)
    r.load_params(params)
    brick_gb = args.brick_gb if args.brick_gb >= 0 else BRICK_GB_DEFAULT
    brick_levels = r.set_brick_budget(brick_gb)
    if args.early_termination > 0:
        r.set_early_termination(args.early_termination)  # opt-in, not the reference's exact arithmetic: see config
    r.set_feature_cutoff(args.feature_cutoff)
    # Feature rows: fp32 on one GPU (the dtype of the reference's outputs["sam"]); with N > 1 the rows are exchanged AND
    # stored as fp16 on every rank (--feature-dtype auto): tinycudann's own per-sample output precision (sam_field.py:51-61),
    # one rounding of the fp32 result (2^-11 relative; the stated feature tolerance is 2e-2).  It halves the 585 MB each of
    # 8 ranks has to receive per frame, which at 770 GB/s per direction would otherwise take 0.76 of the 0.83 ms a rank has.
    if args.feature_dtype == "auto":
        args.feature_dtype = "f16" if world > 1 else "f32"
    fdt = torch.float16 if args.feature_dtype == "f16" else torch.float32
    r.set_feature_dtype(fdt)
    if args.march_first == -1:
        # measured at N = 8 (gpurun calls 17 / 24): multicast push 596 Mrays/s chunk-interleaved vs 557 march-first;
        # peer-pointer push 560 vs 573; at N = 4 (call 12) and N = 2 the interleaved frame is the faster one as well
        args.march_first = 0
    r.set_march_first(bool(args.march_first))
    o_all, d_all = frame_rays(conf)
    n_all = o_all.shape[0]
    assert H % world == 0
    n_loc = n_all // world
    lo = rank * n_loc
    o_host = o_all[lo:lo + n_loc].contiguous().pin_memory()
    d_host = d_all[lo:lo + n_loc].contiguous().pin_memory()
    o_dev, d_dev = o_host.to(dev), d_host.to(dev)
    # Rays per launch.  The reference's eval_num_rays_per_chunk (32768, samconfigs.py:79) caps the [N,S,C] intermediates
    # its PyTorch path materialises; this path has none (128 B + 512 B of scratch per ray), so a launch covers 131072 rays
    # on one GPU (fewer launch tails; measured +8 %), and the tile is cut as finely with N ranks (131072 / N) so that the
    # exchange of chunk c overlaps the compute of chunk c+1 (the chunk size is caller-tunable in the reference too,
    # eval_utils.py:90-91)
    chunk = args.chunk or max(16384, 131072 // world)  # five launches per tile at every N

    names = {"rgb": 3, "depth": 1, "accumulation": 1, "prop_depth_0": 1}
    if "sam" in feats:
        names["sam"] = cfg.sam_out
    dtypes = {k: (fdt if k == "sam" else torch.float32) for k in names}
    esz = {k: (2 if dtypes[k] == torch.float16 else 4) for k in names}
    offs, tot = {}, 0
    for k, c in names.items():  # byte offsets of the outputs inside one frame-sized allocation (256-byte aligned)
        offs[k] = tot
        tot += (n_all * c * esz[k] + 255) // 256 * 256
    # With N > 1 the rendered tiles are exchanged into every rank's frame buffer over NVLink, overlapped with the next
    # chunk's compute; a symmetric-memory barrier ends the frame.  Fallback / comparison: NCCL all-gather (--gather nccl).
    gather_mode, symm, full = "single", None, {}
    if args.gather == "auto":
        # measured on 8 B200 (profiles/r02_multi_gpu.txt): push through the multicast alias 596 Mrays/s, push through peer
        # pointers 573, copy engines 496, NCCL 306 (older commit).  Without a multicast alias pushmc is the peer-pointer push.
        args.gather = "pushmc"
    if world > 1:
        gather_mode = "nccl"
        if args.gather != "nccl":
            try:
                import torch.distributed._symmetric_memory as symm_mem

                big = symm_mem.empty(tot, dtype=torch.uint8, device=dev)
                symm = symm_mem.rendezvous(big, dist.group.WORLD.group_name)
                has_mc = bool(int(getattr(symm, "multicast_ptr", 0) or 0)) and args.gather == "mc"
                for k, c in names.items():
                    full[k] = big[offs[k]:offs[k] + n_all * c * esz[k]].view(dtypes[k]).view(n_all, c)
                gather_mode = ("fused multimem.st (NVSwitch multicast)" if has_mc else
                               "copy engines (cudaMemcpyAsync to peer buffers per chunk)" if args.gather == "dma" else
                               "push kernel (peer stores from 64 CTAs of 128 threads on a side stream, csrc/exchange.cu)" if args.gather == "push" else
                               "push kernel through the NVSwitch multicast alias (multimem.st from 64 CTAs of 128 threads on a side stream)" if args.gather == "pushmc" else
                               "fused peer stores (NVLink P2P)")
            except Exception as e:  # no symmetric memory on this box: say so and use NCCL
                if rank == 0:
                    print(f"[bench] symmetric memory unavailable ({type(e).__name__}: {e}); using NCCL all-gather", file=sys.stderr)
                symm, gather_mode, full = None, "nccl", {}
    for k, c in names.items():
        if k not in full:
            full[k] = torch.empty(n_all, c, device=dev, dtype=dtypes[k])
    mine = {k: v[lo:lo + n_loc] for k, v in full.items()}

    def point_replication_at_peers():
        if symm is None:
            return
        mc = int(getattr(symm, "multicast_ptr", 0) or 0) if args.gather in ("mc", "pushmc") else 0
        r.set_replication_mode("push" if args.gather == "pushmc" else args.gather if args.gather in ("dma", "push") else "stores")
        for k, c in names.items():
            peers = [int(symm.buffer_ptrs[p]) + offs[k] for p in range(world) if p != rank]
            if args.gather == "pushmc":  # feature rows through the multicast alias, the small outputs through peer copies
                r.set_replication(k, full[k], peers, mc + offs[k] if (mc and k == "sam") else 0)
            else:
                r.set_replication(k, full[k], () if mc else peers, mc + offs[k] if mc else 0)

    point_replication_at_peers()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    from samnerf_b200.tiles import all_gather_tiles, ray_block

    assert ray_block(rank, world, H, W) == (lo, lo + n_loc)
    r.set_pipeline(args.pipeline)

    def frame():
        r.render_frame(o_dev, d_dev, get_feature=feats, chunk=chunk, out=mine)
        if world > 1:
            if symm is not None:
                symm.barrier()  # every rank's stores have landed in every frame buffer
            else:
                all_gather_tiles(full, H, W)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        frame()
    barrier()
    exchange_check = None
    if world > 1:
        # (1) every rank must hold the same complete frame: per-tile checksums across ranks; (2) rank 0 renders the whole
        # frame alone into a private buffer and compares it with the gathered one bit for bit (both untimed)
        sums = torch.stack([full[k][p * n_loc:(p + 1) * n_loc].double().nan_to_num().abs().sum() for p in range(world)
                            for k in names])
        allsums = [torch.empty_like(sums) for _ in range(world)]
        dist.all_gather(allsums, sums)
        for p in range(1, world):
            if not torch.equal(allsums[0], allsums[p]) or float(allsums[0].min()) <= 0.0:
                raise SystemExit(f"tile exchange ({gather_mode}) is wrong: rank 0 {allsums[0].tolist()} vs rank {p} {allsums[p].tolist()}")
        if rank == 0:
            for k in names:
                r.set_replication(k, None)
            solo = r.render_frame(o_all.to(dev), d_all.to(dev), get_feature=feats, chunk=chunk)
            torch.cuda.synchronize()
            bad = [k for k in names if not torch.equal(torch.nan_to_num(solo[k]), torch.nan_to_num(full[k]))]
            if bad:
                raise SystemExit(f"gathered frame differs from the single-GPU render of the same rays in {bad}")
            exchange_check = ("gathered frame on rank 0 is bit-identical to a single-GPU render of all rays; "
                              "tile checksums agree on every rank")
            del solo
            point_replication_at_peers()
        barrier()
    launches0 = r.launch_count
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for s, e in ev:
        flush.zero_()  # evict L2 between timed frames (not timed)
        s.record()
        frame()
        e.record()
    barrier()
    launches = r.launch_count - launches0
    # Per-kernel durations for the roofline: the same K steps again with CUDA events around every launch of the
    # three hot kernels (on the launching stream).  Kept out of the region that produces `value` because the event
    # pairs cost ~2 % of it; clocks are sampled across both passes.
    r.kernel_times()
    r.set_timing(True)
    r.set_pipeline(0)  # serial launches, so an event pair brackets exactly one kernel
    for _ in range(args.steps):
        flush.zero_()
        frame()
    barrier()
    r.set_timing(False)
    r.set_pipeline(args.pipeline)
    ktimes = r.kernel_times()
    slot_stats = None
    if args.feature_cutoff >= 0 and "sam" in feats and args.engine == "tcgen05":
        r.feature_slot_stats(reset=True)
        frame()  # one untimed frame to count the slots the bucketed kernel really evaluates
        barrier()
        slot_stats = r.feature_slot_stats(reset=True)
    total_ms = sum(s.elapsed_time(e) for s, e in ev)
    t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    value = n_all * args.steps / (total_ms * 1e-3) / 1e6

    # ---- end to end: pinned host rays in, host images + features out, copies inside the timed region ----------
    # The library's copy-engine exchange takes any UVA destination, so the pinned host buffer is registered as the one
    # "peer": each chunk's feature rows leave for host memory as soon as its output layer has finished, overlapped with
    # the next chunk's render; the caller's stream resumes when the last copy has landed.  Every rank feeds and drains
    # its own tile over its own PCIe link (no NVLink exchange here: the frame is assembled in host memory).
    out_host = {k: torch.empty(n_loc, c, dtype=dtypes[k]).pin_memory() for k, c in names.items()}
    o_stage, d_stage = torch.empty_like(o_dev), torch.empty_like(d_dev)
    r.set_replication_mode("dma")
    for k, c in names.items():
        r.set_replication(k, full[k], [out_host[k].data_ptr() - lo * c * esz[k]], 0)  # offset-preserving alias of this tile
    e2e_chunk = args.chunk or 32768

    def frame_e2e():
        o_stage.copy_(o_host, non_blocking=True)
        d_stage.copy_(d_host, non_blocking=True)
        r.render_frame(o_stage, d_stage, get_feature=feats, chunk=e2e_chunk, out=mine)

    for _ in range(2):
        frame_e2e()
    barrier()
    e2e_ok = all(torch.equal(torch.nan_to_num(out_host[k]), torch.nan_to_num(mine[k].cpu())) for k in names)
    e2e_steps = max(3, min(args.steps, 5))
    ev2 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(e2e_steps)]
    for s, e in ev2:
        s.record()
        frame_e2e()
        e.record()
    barrier()
    clk = clocks.stop() if rank == 0 else None
    for k in names:
        r.set_replication(k, None)
    t2 = torch.tensor([sum(s.elapsed_time(e) for s, e in ev2)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    e2e_value = n_all * e2e_steps / (float(t2.item()) * 1e-3) / 1e6
    h2d = 2 * n_all * 3 * 4
    d2h = n_all * sum(c * esz[k] for k, c in names.items())
    if not e2e_ok:
        raise SystemExit("end-to-end path: the host copy differs from the device frame")

    if rank == 0:
        peak, peak_src = measured_peaks()
        traffic, traffic_note = ncu_traffic()
        kernels = {}
        for k in ("march", "feature", "tapgemm"):
            ms, cnt = ktimes[k]
            if cnt == 0:
                continue
            rays_per_launch = n_loc * args.steps / cnt
            bpr = BYTES_PER_RAY[k]
            if k == "tapgemm" and args.feature_dtype == "f16":
                bpr = 512 + 512  # fp16 hidden sums in, fp16 rows out
            extra = {}
            if k == "feature" and slot_stats is not None:
                slots_per_ray = slot_stats[1] / max(n_loc, 1)
                bpr = 3072.0 * slots_per_ray
                extra = {"feature_slots": {"rays_per_bucket_1_2_4_8_16": slot_stats[0], "slots_per_ray": slots_per_ray,
                                           "note": "bucketed kernel: achieved / algorithmic bytes count the evaluated slots only "
                                                   "(3072 B each); all 16 slots of every ray would be 49152 B/ray"}}
            ach = bpr * rays_per_launch / (ms / cnt * 1e-3) / 1e9
            kernels[k] = {"ms_per_step": ms / args.steps, "avg_launch_ms": ms / cnt, "launches_timed": cnt,
                          "algorithmic_bytes_per_ray": bpr, "algorithmic_bytes_per_launch": bpr * rays_per_launch,
                          "achieved": ach, "frac": ach / peak,
                          "traffic": traffic[k] * rays_per_launch if k in traffic else None, **extra}
        top = max(kernels, key=lambda k: kernels[k]["ms_per_step"])
        names_long = {"march": "march_kernel (proposal sampling + nerfacto field + compositing + top-k)",
                      "feature": ("sam_bucket_kernel" if slot_stats is not None else "sam_kernel") +
                                 " (feature-field gather + MLP layer 1 + weighted sum)",
                      "tapgemm": "tapgemm_kernel (feature MLP output layer)"}
        path_bytes = sum(kernels[k]["algorithmic_bytes_per_ray"] for k in kernels)
        line = {
            "metric": conf["metric"], "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f16 operands / f32 accumulate", "data": "synthetic",
            "config": {"workload": conf["workload"], "regime": args.regime, "engine": args.engine,
                       "l2": "256 MiB buffer zeroed between timed frames (untimed); tables + bricks + outputs > 126 MB L2",
                       "tiles": (f"{world} row blocks of {H // world} rows; outputs exchanged by {gather_mode}, "
                                 "frame ends with a symmetric-memory barrier") if world > 1 else "single GPU",
                       "exchange_check": exchange_check,
                       "early_termination": args.early_termination or None,
                       "feature_cutoff": args.feature_cutoff if args.feature_cutoff >= 0 else None,
                       "bricks": {"budget_gib": brick_gb, "proposal_levels": brick_levels[0], "field_levels": brick_levels[1]},
                       "feature_dtype": {"rows": args.feature_dtype,
                                         "note": None if args.feature_dtype == "f32" else
                                         "feature rows are exchanged and stored as fp16 on every rank: one rounding of the fp32 result "
                                         "(2^-11 relative, tinycudann's own output precision; stated feature tolerance 2e-2); N = 1 writes fp32"},
                       "march_first": bool(args.march_first),
                       "chunk": chunk, "e2e_chunk": e2e_chunk,
                       "pipeline": "chunks pipelined over 3 streams" if (args.pipeline == 2 or (args.pipeline == 1 and world > 1 and symm is not None and args.gather in ("mc", "peer"))) else "sequential"},
            "roofline": {"bound": "hbm", "kernel": names_long[top],
                         "achieved": kernels[top]["achieved"], "peak": peak, "unit": "GB/s", "frac": kernels[top]["frac"],
                         "traffic": kernels[top]["traffic"], "peak_source": peak_src, "traffic_note": traffic_note,
                         "algorithmic_bytes_per_launch": kernels[top]["algorithmic_bytes_per_launch"],
                         "avg_launch_ms": kernels[top]["avg_launch_ms"], "launches_timed": kernels[top]["launches_timed"],
                         "why_this_kernel": "largest share of the step (kernel_share_ms_per_step); every kernel's own figures are under `kernels`",
                         "timing_note": "CUDA events around each launch on the launching stream, instrumented repeat of the K timed steps (same inputs, L2 flushed)",
                         "kernels": {names_long[k].split(" ")[0]: v for k, v in kernels.items()},
                         "path": {"bytes_per_ray": path_bytes, "achieved": value * 1e6 * path_bytes / 1e9 / world,
                                  "frac": value * 1e6 * path_bytes / 1e9 / world / peak},
                         "kernel_share_ms_per_step": {k: v["ms_per_step"] for k, v in kernels.items()}},
            "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps,
                    "api": "Renderer.render_frame on pinned host rays; the library's copy engines push each chunk's rows to pinned host "
                           "memory while the next chunk renders (per rank: its own tile over its own PCIe link)"},
            "gpu_launches": launches,
            "clocks": clk,
        }
        if world == 1 and not args.no_cpu_baseline:
            v, _, cores = cpu_rate(args.config, 3, 1, args.regime)
            line["cpu_baseline"] = {"value": v, "unit": "Mrays/s", "cores": cores, "kind": "port", "sample": CPU_SAMPLE_TEXT}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["native", "reference"], default="native")
    ap.add_argument("--config", choices=sorted(CONFIGS), default="sam",
                    help="sam = BASELINE.json configs[2] (headline, default); rgb = configs[1]; clipseg_patch = configs[3]")
    ap.add_argument("--regime", choices=["scene", "init"], default="scene")
    ap.add_argument("--engine", choices=["tcgen05", "mma_sync"], default="tcgen05")
    ap.add_argument("--gather", choices=["auto", "mc", "peer", "dma", "push", "pushmc", "nccl"], default="auto",
                    help="N > 1: how the tiles are exchanged (auto = push kernel; copy engines; fused multicast / peer stores; NCCL)")
    ap.add_argument("--feature-dtype", choices=["auto", "f32", "f16"], default="auto",
                    help="element type of the 256-d feature rows: auto = f32 on one GPU, f16 with N > 1 (wire and storage)")
    ap.add_argument("--march-first", type=int, default=-1, choices=[-1, 0, 1],
                    help="frame = one march launch over the tile, then the feature chunks (-1 = auto: off)")
    ap.add_argument("--chunk", type=int, default=0, help="rays per launch (default: 131072 on one GPU, finer with N > 1)")
    ap.add_argument("--pipeline", type=int, default=1, choices=[0, 1, 2],
                    help="chunk pipelining over 3 streams: 0 off, 1 auto (only with replicated outputs, N > 1), 2 always")
    ap.add_argument("--early-termination", type=float, default=0.0,
                    help="opt-in transmittance threshold below which the MLPs of a ray's last 16 nerf samples are skipped "
                         "(0 = exact, the default and the headline configuration)")
    ap.add_argument("--feature-cutoff", type=float, default=2.0 ** -24,
                    help="bucketed feature kernel: evaluate only the leading picked samples of a ray whose sharpened weight is "
                         ">= this (library default 2^-24 = below one fp32 ulp of the sum; 0 = drop exact zeros); "
                         "< 0 = every sample of every ray through the un-bucketed kernel")
    ap.add_argument("--brick-gb", type=float, default=-1.0,
                    help="HBM budget (GiB) for the cell-major brick copies of the leading grid levels (bench default 62 = 14 of "
                         "the 16 nerfacto levels, 59.3 GiB of the 180 GB; 24 = 13 levels, 10 = 12 levels; library default 4 = 11 "
                         "levels; 0 = off; a pure re-layout, results are bit-identical)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
