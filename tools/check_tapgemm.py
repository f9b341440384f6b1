"""Output layer: the TMA-fed kernel (gemm_tma.cu) against the round-1 kernel (gemm.cu, SNRF_TAPGEMM=v1) on the same inputs,
several row counts including tails; run each variant in its own process (the switch is read once per process)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = r'''
import os, sys, torch
sys.path.insert(0, os.environ["SNRF_ROOT"]); sys.path.insert(0, os.path.join(os.environ["SNRF_ROOT"], "tests"))
from helpers import make_renderer, model_pair
cfg, params, orc = model_pair("tiny", "scene", 11, True, 1)
r = make_renderer(cfg, params)
r.set_feature_cutoff(-1.0)
outs = {}
for which in ("sam", "clipseg"):
    for n in (1, 127, 128, 129, 1000, 40000):
        g = torch.Generator().manual_seed(n)
        o = torch.randn(n, 3, generator=g) * 0.3
        d = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1)
        t = torch.rand(n, 16, generator=g) * 4
        w = torch.softmax(torch.randn(n, 16, generator=g) * 3, -1)
        out, _ = r.feature_forward(which, o, d, t, w)
        torch.cuda.synchronize()
        outs[f"{which}_{n}"] = out.cpu()
torch.save(outs, os.environ["SNRF_OUT"])
print("done", len(outs))
'''


def main():
    import torch

    res = {}
    for tag in ("v1", "tma"):
        env = dict(os.environ, SNRF_ROOT=ROOT, SNRF_OUT=f"/tmp/tapgemm_{tag}.pt")
        if tag == "v1":
            env["SNRF_TAPGEMM"] = "v1"
        p = subprocess.run([sys.executable, "-c", WORKER], env=env, capture_output=True, text=True, timeout=240)
        print(tag, p.stdout.strip()[-200:], p.stderr.strip()[-800:] if p.returncode else "")
        if p.returncode:
            raise SystemExit(f"{tag} failed")
        res[tag] = torch.load(f"/tmp/tapgemm_{tag}.pt")
    bad = 0
    for k in res["v1"]:
        a, b = res["v1"][k], res["tma"][k]
        same = torch.equal(a, b)
        err = float((a - b).abs().max())
        print(f"{k:16s} identical={same} max|diff|={err:.3e} finite={bool(torch.isfinite(b).all())}")
        bad += (not same) and err > 1e-5 * float(a.abs().max())
    print("TAPGEMM CHECK", "OK" if not bad else f"FAILED ({bad})")


if __name__ == "__main__":
    main()
