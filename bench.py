#!/usr/bin/env python
"""Headline benchmark: Mrays/s rendering RGB + 256-d SAM features at 800x800 (BASELINE.json), 1/2/4/8 B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \\
        bench.py --gpus N --steps K --warmup W

One "step" = one 800x800 frame of the `samnerf_distill` model (SURVEY.md 8 d config 3): 640 000 rays, every ray
rendered to rgb / median depth / accumulation / proposal depth AND the 256-d SAM feature (k = 16 top samples,
sharpening 10, patch 1), in the reference's chunks of 32 768 rays, through `libsnrf` (C ABI).  With N > 1 the frame is
cut into N contiguous row blocks ("screen tiles"), one per rank, and the rendered tiles are exchanged into every
rank's frame buffer (symmetric memory): by copy engines per chunk (default), by the kernels' own multicast / peer
stores, or by an NCCL all-gather (--gather).  Total work is fixed, so scaling is "strong".  Synthetic scene-like
parameters (seed 0), synthetic orbit camera.

`--impl reference` times the reference's own CPU path: the reference is Python + the CUDA-only tinycudann, so its CPU
implementation is the oracle port (oracle/samnerf_oracle.py, validated against the reference's own modules by
oracle/make_golden.py), all host threads, on a bounded sample of the same frame per step.
"""
from __future__ import annotations

import argparse
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

H = W = 800
BYTES_PER_RAY = {"proposal": 10240, "field": 16384, "sam": 49152}  # SURVEY.md 8 d / BASELINE.md section 3
METRIC = "Mrays/s rendering RGB+256-d SAM features at 800x800"
WORKLOAD = "samnerf_distill 800x800 RGB+256-d SAM (k=16, p=1): 640000 rays in chunks of 32768"


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(rays_per_launch: float):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel from the committed ncu capture, scaled
    to this run's rays per launch (None when no capture is committed)."""
    path = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if not os.path.exists(path):
        return None
    with open(path) as f:
        t = json.load(f)["sam_kernel"]
    return (t["dram_bytes_read"] + t["dram_bytes_write"]) * rays_per_launch / t["rays_per_launch"]


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


def frame_rays():
    from samnerf_b200.synthetic import orbit_rays

    o, d = orbit_rays(H, W, 800.0)
    return o.reshape(-1, 3).contiguous(), d.reshape(-1, 3).contiguous()


def cpu_reference_rate(n_rays: int, repeats: int, cfg, params):
    """Oracle port on a strided sample of the frame, all host threads.  Returns (Mrays/s best, cores, sample text)."""
    from oracle.samnerf_oracle import Oracle

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    orc = Oracle(cfg, params)
    o, d = frame_rays()
    idx = (torch.arange(n_rays) * (o.shape[0] // n_rays)).long()
    o, d = o[idx].contiguous(), d[idx].contiguous()
    best = float("inf")
    for _ in range(repeats):
        t0 = time.perf_counter()
        with torch.no_grad():
            orc.render_rays(o, d, get_feature=("sam",))
        best = min(best, time.perf_counter() - t0)
    return n_rays / best / 1e6, cores, f"{n_rays} rays strided over the 800x800 frame, best of {repeats}, torch CPU fp32"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from samnerf_b200 import SAMNeRFConfig, make_synthetic_params
    from oracle.samnerf_oracle import Oracle

    cfg = SAMNeRFConfig.distill(clipseg=False, patch_size=1)
    params = make_synthetic_params(cfg, args.regime, 0)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    orc = Oracle(cfg, params)
    o, d = frame_rays()
    n = args.ref_rays
    times = []
    for step in range(args.warmup + args.steps):
        idx = ((torch.arange(n) * (o.shape[0] // n)) + step) % o.shape[0]
        oo, dd = o[idx].contiguous(), d[idx].contiguous()
        t0 = time.perf_counter()
        with torch.no_grad():
            orc.render_rays(oo, dd, get_feature=("sam",))
        dt = time.perf_counter() - t0
        if step >= args.warmup:
            times.append(dt)
    total = sum(times)
    value = n * len(times) / total / 1e6
    sample = f"{n} rays strided over the 800x800 frame per step (bounded sample of the 640000-ray frame)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "regime": args.regime, "note": "reference CPU path = oracle port (tinycudann is CUDA-only)"},
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_native(args):
    import torch.distributed as dist

    from samnerf_b200 import SAMNeRFConfig, make_synthetic_params
    from samnerf_b200.renderer import Renderer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    cfg = SAMNeRFConfig.distill(clipseg=False, patch_size=1)
    params = make_synthetic_params(cfg, args.regime, 0)
    r = Renderer(cfg, device=local, engine=args.engine)
    r.load_params(params)
    brick_levels = r.set_brick_budget(args.brick_gb) if args.brick_gb >= 0 else None
    if args.early_termination > 0:
        r.set_early_termination(args.early_termination)  # opt-in, not the reference's exact arithmetic: see config
    if args.feature_cutoff >= 0:
        r.set_feature_cutoff(args.feature_cutoff)  # opt-in bucketed feature kernel: see config
    o_all, d_all = frame_rays()
    n_all = o_all.shape[0]
    assert H % world == 0
    n_loc = n_all // world
    lo = rank * n_loc
    o_host = o_all[lo:lo + n_loc].contiguous().pin_memory()
    d_host = d_all[lo:lo + n_loc].contiguous().pin_memory()
    o_dev, d_dev = o_host.to(dev), d_host.to(dev)
    # reference chunk size on one GPU; with N ranks the tile is cut finer so that the NVLink exchange of chunk c
    # overlaps the compute of chunk c+1 (the chunk size is caller-tunable in the reference too, eval_utils.py:90-91)
    # measured: 32768 is best up to 2 ranks, 16384 at 4 and 8 (profiles/r01_multi_gpu.txt)
    chunk = args.chunk or (cfg.eval_num_rays_per_chunk if world <= 2 else 16384)

    # frame-sized outputs; with N > 1 each rank renders into its row block of the gathered frame (in-place all-gather)
    names = {"rgb": 3, "depth": 1, "accumulation": 1, "prop_depth_0": 1, "sam": cfg.sam_out}
    # With N > 1 the rendered tiles are exchanged by the kernels themselves: every output row is stored into every
    # rank's frame buffer over NVLink (one multimem.st through the NVSwitch multicast alias when available, else
    # peer-mapped pointers), overlapped with the next chunk's compute; a symmetric-memory barrier ends the frame.
    # Fallback / comparison: NCCL all-gather of the rendered tiles (--gather nccl).
    gather_mode, symm, full = "single", None, {}
    if args.gather == "auto":
        # what was measured: copy engines at N = 2 and 4, the fused multicast stores at N = 8 (278.9 Mrays/s,
        # profiles/r01_multi_gpu.txt; the copy-engine mode has not been run on 8 GPUs yet)
        args.gather = "mc" if world >= 8 else "dma"
    if world > 1:
        gather_mode = "nccl"
        if args.gather != "nccl":
            try:
                import torch.distributed._symmetric_memory as symm_mem

                big = symm_mem.empty(n_all * sum(names.values()), dtype=torch.float32, device=dev)
                symm = symm_mem.rendezvous(big, dist.group.WORLD.group_name)
                mc = int(getattr(symm, "multicast_ptr", 0) or 0) if args.gather == "mc" else 0
                if args.gather == "dma":
                    r.set_replication_mode("dma")
                off = 0
                for k, c in names.items():
                    full[k] = big[off:off + n_all * c].view(n_all, c)
                    peers = [int(symm.buffer_ptrs[p]) + off * 4 for p in range(world) if p != rank]
                    r.set_replication(k, full[k], () if mc else peers, mc + off * 4 if mc else 0)
                    off += n_all * c
                gather_mode = ("fused multimem.st (NVSwitch multicast)" if mc else
                               "copy engines (cudaMemcpyAsync to peer buffers per chunk)" if args.gather == "dma" else
                               "fused peer stores (NVLink P2P)")
            except Exception as e:  # no symmetric memory on this box: say so and use NCCL
                if rank == 0:
                    print(f"[bench] symmetric memory unavailable ({type(e).__name__}: {e}); using NCCL all-gather", file=sys.stderr)
                symm, gather_mode, full = None, "nccl", {}
    for k, c in names.items():
        if k not in full:
            full[k] = torch.empty(n_all, c, device=dev)
    mine = {k: v[lo:lo + n_loc] for k, v in full.items()}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    from samnerf_b200.tiles import all_gather_tiles, ray_block

    assert ray_block(rank, world, H, W) == (lo, lo + n_loc)
    r.set_pipeline(args.pipeline)

    def frame():
        r.render_frame(o_dev, d_dev, get_feature=("sam",), chunk=chunk, out=mine)
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
    if world > 1:
        # every rank must now hold the same complete frame: compare per-tile checksums across ranks (untimed)
        sums = torch.stack([full["sam"][p * n_loc:(p + 1) * n_loc].double().nan_to_num().abs().sum() for p in range(world)]
                           + [full["rgb"].double().sum()])
        allsums = [torch.empty_like(sums) for _ in range(world)]
        dist.all_gather(allsums, sums)
        for p in range(1, world):
            if not torch.equal(allsums[0], allsums[p]) or float(allsums[0].min()) <= 0.0:
                raise SystemExit(f"tile exchange ({gather_mode}) is wrong: rank 0 {allsums[0].tolist()} vs rank {p} {allsums[p].tolist()}")
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
    # three hot kernels (on the launching stream).  Kept out of the region that produces `value` because 60 event
    # pairs per frame cost ~2 % of it; clocks are sampled across both passes.
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
    if args.feature_cutoff >= 0:
        r.feature_slot_stats(reset=True)
        frame()  # one untimed frame to count the slots the bucketed kernel really evaluates
        barrier()
        slot_stats = r.feature_slot_stats(reset=True)
    clk = clocks.stop() if rank == 0 else None
    total_ms = sum(s.elapsed_time(e) for s, e in ev)
    t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    value = n_all * args.steps / (total_ms * 1e-3) / 1e6

    # ---- end to end: pinned host rays in, host images + features out, copies inside the timed region ----------
    out_host = {k: torch.empty(n_loc, c).pin_memory() for k, c in names.items()}
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    o_stage, d_stage = torch.empty_like(o_dev), torch.empty_like(d_dev)

    def frame_e2e():
        cur = torch.cuda.current_stream(dev)
        done = []
        for i in range(0, n_loc, chunk):
            sl = slice(i, min(i + chunk, n_loc))
            with torch.cuda.stream(s_in):
                o_stage[sl].copy_(o_host[sl], non_blocking=True)
                d_stage[sl].copy_(d_host[sl], non_blocking=True)
                ready = torch.cuda.Event()
                ready.record(s_in)
            cur.wait_event(ready)
            r.render(o_stage[sl], d_stage[sl], get_feature=("sam",), out={k: v[sl] for k, v in mine.items()})
            rendered = torch.cuda.Event()
            rendered.record(cur)
            last = i + chunk >= n_loc
            with torch.cuda.stream(s_out):
                s_out.wait_event(rendered)
                # the 256-d features (98 % of the bytes) leave chunk by chunk, overlapped with the next chunk's
                # render; the four small per-ray outputs go once, after the last chunk (4 copies instead of 4 per chunk)
                out_host["sam"][sl].copy_(mine["sam"][sl], non_blocking=True)
                if last:
                    for k in names:
                        if k != "sam":
                            out_host[k].copy_(mine[k], non_blocking=True)
                fin = torch.cuda.Event()
                fin.record(s_out)
            done.append(fin)
        for f in done:
            cur.wait_event(f)

    for _ in range(2):
        frame_e2e()
    barrier()
    e2e_steps = max(3, min(args.steps, 5))
    ev2 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(e2e_steps)]
    for s, e in ev2:
        s.record()
        frame_e2e()
        e.record()
    barrier()
    t2 = torch.tensor([sum(s.elapsed_time(e) for s, e in ev2)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    e2e_value = n_all * e2e_steps / (float(t2.item()) * 1e-3) / 1e6
    h2d = 2 * n_all * 3 * 4
    d2h = n_all * sum(names.values()) * 4

    if rank == 0:
        peak, peak_src = measured_peaks()
        f_ms, f_cnt = ktimes["feature"]
        m_ms, m_cnt = ktimes["march"]
        g_ms, g_cnt = ktimes["tapgemm"]
        rays_per_launch = n_loc * args.steps / max(f_cnt, 1)
        ach = BYTES_PER_RAY["sam"] * rays_per_launch / (f_ms / max(f_cnt, 1) * 1e-3) / 1e9 if f_ms > 0 else None
        if slot_stats is not None and f_ms > 0:
            # bucketed kernel: state the roofline on the slots it gathers (3 072 B each), not on all 16 per ray
            slots_per_ray = slot_stats[1] / max(n_loc, 1)
            ach = 3072.0 * slots_per_ray * rays_per_launch / (f_ms / max(f_cnt, 1) * 1e-3) / 1e9
        path_bytes = sum(BYTES_PER_RAY.values())
        line = {
            "metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f16 operands / f32 accumulate", "data": "synthetic",
            "config": {"workload": WORKLOAD, "regime": args.regime, "engine": args.engine,
                       "l2": "256 MiB buffer zeroed between timed frames (untimed); tables 158 MB + outputs 668 MB > 126 MB L2",
                       "tiles": (f"{world} row blocks of {H // world} rows; 256-d features exchanged by {gather_mode}, "
                                 "frame ends with a symmetric-memory barrier") if world > 1 else "single GPU",
                       "early_termination": args.early_termination or None,
                       "feature_cutoff": args.feature_cutoff if args.feature_cutoff >= 0 else None,
                       "bricks": None if brick_levels is None else {"budget_gib": args.brick_gb, "proposal_levels": brick_levels[0], "field_levels": brick_levels[1]},
                       "chunk": chunk, "pipeline": "chunks pipelined over 3 streams" if (args.pipeline == 2 or (args.pipeline == 1 and world > 1 and symm is not None and args.gather in ("mc", "peer"))) else "sequential"},
            "roofline": {"bound": "hbm", "kernel": "sam_kernel (feature-field gather + MLP layer 1 + weighted sum)",
                         "achieved": ach, "peak": peak, "unit": "GB/s", "frac": (ach / peak) if ach else None,
                         "traffic": ncu_traffic(rays_per_launch), "peak_source": peak_src,
                         "traffic_note": "ncu dram bytes of one launch (profiles/r01_traffic.json): far below the "
                                         "algorithmic bytes - the launch's table footprint is L2/L1-resident; the kernel "
                                         "is bound by the L1 tag stage and issue, not HBM (DESIGN.md section 4 B)",
                         "algorithmic_bytes_per_launch": (BYTES_PER_RAY["sam"] if slot_stats is None else
                                                          3072.0 * slot_stats[1] / max(n_loc, 1)) * rays_per_launch,
                         "feature_slots": None if slot_stats is None else {
                             "rays_per_bucket_1_2_4_8_16": slot_stats[0], "slots_per_ray": slot_stats[1] / max(n_loc, 1),
                             "note": "bucketed kernel: achieved / algorithmic bytes count the evaluated slots only; "
                                     "all 16 slots would be 49152 B/ray"},
                         "avg_launch_ms": f_ms / max(f_cnt, 1), "launches_timed": f_cnt,
                         "timing_note": "CUDA events around each launch, instrumented repeat of the K timed steps (same inputs, L2 flushed)",
                         "path": {"bytes_per_ray": path_bytes, "achieved": value * 1e6 * path_bytes / 1e9 / world,
                                  "frac": value * 1e6 * path_bytes / 1e9 / world / peak},
                         "kernel_share_ms_per_step": {"march": m_ms / args.steps, "feature": f_ms / args.steps,
                                                      "tapgemm": g_ms / args.steps}},
            "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "api": "Renderer.render per chunk, pinned host buffers, 3-stream pipeline"},
            "gpu_launches": launches,
            "clocks": clk,
        }
        if world == 1 and not args.no_cpu_baseline:
            v, cores, sample = cpu_reference_rate(args.cpu_rays, 2, cfg, params)
            line["cpu_baseline"] = {"value": v, "unit": "Mrays/s", "cores": cores, "kind": "port", "sample": sample}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["native", "reference"], default="native")
    ap.add_argument("--regime", choices=["scene", "init"], default="scene")
    ap.add_argument("--engine", choices=["tcgen05", "mma_sync"], default="tcgen05")
    ap.add_argument("--gather", choices=["auto", "mc", "peer", "dma", "nccl"], default="auto",
                    help="N > 1: how the feature tiles are exchanged (fused multicast / peer stores, or NCCL)")
    ap.add_argument("--chunk", type=int, default=0, help="rays per chunk (default: 32768, finer with N > 1)")
    ap.add_argument("--pipeline", type=int, default=1, choices=[0, 1, 2],
                    help="chunk pipelining over 3 streams: 0 off, 1 auto (only with replicated outputs, N > 1), 2 always")
    ap.add_argument("--early-termination", type=float, default=0.0,
                    help="opt-in transmittance threshold below which a ray's last 16 nerf samples are skipped "
                         "(0 = exact, the default and the headline configuration)")
    ap.add_argument("--feature-cutoff", type=float, default=-1.0,
                    help="opt-in bucketed feature kernel: evaluate only the leading picked samples of a ray whose "
                         "sharpened weight is >= this (0 = drop exact zeros, 5.96e-8 = below one fp32 ulp of the sum); "
                         "< 0 = every sample (default and headline configuration)")
    ap.add_argument("--brick-gb", type=float, default=-1.0,
                    help="HBM budget (GiB) for the cell-major brick copies of the leading grid levels (library default 4; "
                         "0 = off; a pure re-layout, results are bit-identical)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-rays", type=int, default=32768,
                    help="rays of the bounded CPU-baseline sample (one reference chunk, timed twice: ~4 s of host work at "
                         "the ~0.017 Mrays/s measured on the GPU box, ~45 s on a slow 8-core host)")
    ap.add_argument("--ref-rays", type=int, default=4096)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
