"""Two-rank render on two GPUs: tile all-gather (multicast stores, peer stores, copy engines) == single-GPU frame.
Needs >= 2 CUDA devices (skipped on the 1-GPU test box; run with `gpurun --gpus 2 -- pytest tests/test_multi_gpu.py -m gpu`)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["SNRF_ROOT"]); sys.path.insert(0, os.path.join(os.environ["SNRF_ROOT"], "tests"))
from helpers import make_renderer, model_pair, test_rays
from samnerf_b200.tiles import ray_block
import torch.distributed._symmetric_memory as symm_mem
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
cfg, params, _ = model_pair("tiny", "scene", 11, False, 1)
from samnerf_b200.renderer import Renderer
r = Renderer(cfg, device=rank); r.load_params(params)
H, W = 64, 96
o, d = test_rays(H * W, seed=3)
names = {"rgb": 3, "depth": 1, "accumulation": 1, "prop_depth_0": 1, "sam": 256}
single = r.render_frame(o.to(dev), d.to(dev), get_feature=("sam",), chunk=1024)      # whole frame on this GPU
mode = os.environ["SNRF_GATHER"]
f16 = mode.endswith("_f16")          # fp16 feature rows (snrf_set_feature_dtype), exchanged and stored as such
march_first = "_mf" in mode          # one march launch over the tile, then the feature chunks
mode = mode.split("_")[0]
if f16:
    r.set_feature_dtype(torch.float16)
    single = r.render_frame(o.to(dev), d.to(dev), get_feature=("sam",), chunk=1024)
    ref32 = r.feature_dtype
esz = {k: (2 if (f16 and k == "sam") else 4) for k in names}
offs, tot = {}, 0
for k, c in names.items():
    offs[k] = tot
    tot += (H * W * c * esz[k] + 255) // 256 * 256
big = symm_mem.empty(tot, dtype=torch.uint8, device=dev)
big.fill_(0x5A)
hdl = symm_mem.rendezvous(big, dist.group.WORLD.group_name)
mc = int(hdl.multicast_ptr or 0) if mode == "mc" else 0
if mode in ("dma", "push"):
    r.set_replication_mode(mode)
r.set_march_first(march_first)
if mode == "mc" and not mc:
    print("SKIP no multicast"); dist.destroy_process_group(); sys.exit(0)
full = {}
for k, c in names.items():
    dt = torch.float16 if esz[k] == 2 else torch.float32
    full[k] = big[offs[k]:offs[k] + H * W * c * esz[k]].view(dt).view(H * W, c)
    peers = [int(hdl.buffer_ptrs[p]) + offs[k] for p in range(world) if p != rank]
    r.set_replication(k, full[k], () if mc else peers, mc + offs[k] if mc else 0)
lo, hi = ray_block(rank, world, H, W)
hdl.barrier()
r.render_frame(o[lo:hi].to(dev), d[lo:hi].to(dev), get_feature=("sam",), chunk=1024, out={k: v[lo:hi] for k, v in full.items()})
hdl.barrier()
torch.cuda.synchronize()
for k in names:
    assert torch.equal(full[k], single[k]), (k, rank, mode)
print("OK", rank, mode)
dist.destroy_process_group()
'''


@pytest.mark.parametrize("mode", ["peer", "mc", "dma", "push", "dma_f16", "push_mf_f16"])
def test_fused_tile_all_gather(mode, tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, SNRF_ROOT=ROOT, SNRF_GATHER=mode)
    out = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
         "--master-port", "29533", str(script)], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert out.stdout.count("OK") == 2 or "SKIP" in out.stdout, out.stdout[-2000:]
