"""One-off hardware diagnostics for the round-2 GPU test failures (not part of the product or the test suite).

    python tools/diag_round2.py            # prints findings; never raises on a numerical difference
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402


def full_training_step():
    from helpers import model_pair, test_rays
    from samnerf_b200.nerfstudio_api import RayBundle, SAMModel

    cfg, params, _ = model_pair("tiny", "scene", 25, False, 1)
    m = SAMModel(cfg)
    m.load_state_dict(params)
    m.train()
    m.proposal_sampler.train_stratified = False
    o, d = test_rays(1024, seed=6)
    bundle = RayBundle(origins=o.cuda(), directions=d.cuda())
    gen = torch.Generator().manual_seed(3)
    image = torch.rand(1024, 3, generator=gen).cuda()
    feat = (torch.randn(1024, 256, generator=gen) * 0.1).cuda()
    groups = m.get_param_groups()
    out = m(bundle, get_feature=["sam"])
    print("[train] sam rows with NaN:", int(torch.isnan(out["sam"]).any(-1).sum()), "of", out["sam"].shape[0],
          "| rgb finite:", bool(torch.isfinite(out["rgb"]).all()))
    rgb_loss = torch.nn.functional.mse_loss(out["rgb"], image)
    sam_loss = torch.nn.functional.mse_loss(out["sam"], feat, reduction="none").mean(dim=-1).nanmean()
    (rgb_loss + sam_loss).backward()
    for gname, g in groups.items():
        for i, q in enumerate(g):
            if q.grad is None:
                print(f"[train] {gname}[{i}] shape {tuple(q.shape)}: grad is None")
            else:
                bad = int((~torch.isfinite(q.grad)).sum())
                print(f"[train] {gname}[{i}] shape {tuple(q.shape)}: non-finite {bad} of {q.grad.numel()}, max|g| "
                      f"{float(q.grad.nan_to_num().abs().max()):.3e}")


def feature_backward_errors():
    import test_backward as tb
    from helpers import make_renderer, model_pair

    for which, clipseg in (("sam", False), ("clipseg", True)):
        cfg, params, orc0 = model_pair("tiny", "scene", 21, clipseg, 1)
        r = make_renderer(cfg, params)
        o, d, sam_t, sam_w = tb._branch_inputs(cfg, orc0, 700, seed=9, which=which)
        ok = torch.isfinite(sam_w).all(-1)
        o, d, sam_t, sam_w = o[ok], d[ok], sam_t[ok], sam_w[ok]
        sam_w, order = sam_w.sort(dim=-1, descending=True)
        sam_t = torch.gather(sam_t, 1, order)
        got_out, enc = r.feature_forward(which, o, d, sam_t, sam_w)
        g_out = torch.randn(got_out.shape, generator=torch.Generator().manual_seed(1))
        grads = r.feature_backward(which, o, d, sam_t, sam_w, enc, g_out)
        torch.cuda.synchronize()
        net_name = f"sam_field.{which}_net.params"
        for label in ("oracle's own encoder outputs", "device-saved encoder outputs"):
            orc, p = tb._fresh_oracle(cfg, params)
            if label.startswith("oracle"):
                out, f = tb.oracle_branch(orc, which, o, d, sam_t, sam_w)
            else:
                out = tb.oracle_branch_at(orc, which, o, d, sam_t, sam_w, enc.cpu())
            (out * g_out).sum().backward()
            want = p[net_name].grad
            got = grads["net"].cpu()
            err = (got - want).abs()
            n1 = 256 * 192
            i = int(err.argmax())
            where = f"W1[{i // 192},{i % 192}]" if i < n1 else f"W2[{(i - n1) // 256},{(i - n1) % 256}]"
            rows = err[:n1].view(256, 192).max(dim=1).values
            print(f"[feat-bwd {which}] vs {label}: max|err| {float(err.max()):.3e} at {where} (max|grad| {float(want.abs().max()):.3e}); "
                  f"dW1 rows with err > 1e-2: {int((rows > 1e-2).sum())}; dW2 max err {float(err[n1:].max()):.3e}")
            if not label.startswith("oracle"):
                x = enc.cpu().float().reshape(-1, 192)
                x0 = f["hashgrid"].reshape(-1, 192) if which == "sam" else None
                if x0 is not None:
                    print(f"[feat-bwd {which}] encoder outputs differing from the oracle: {float((x != x0.detach()).float().mean()):.4f} of elements, "
                          f"max diff {float((x - x0.detach()).abs().max()):.3e}")


def march_new_vs_v1():
    """Where do the new march kernel's per-sample densities leave the oracle's band, and what do the edges look like there?"""
    import subprocess
    import numpy as np
    from helpers import make_renderer, model_pair, test_rays

    cfg, params, orc = model_pair("full", "scene", 0, False, 1)
    o, d = test_rays(1024, seed=5)
    ref = orc.render_rays(o, d, get_feature=("sam",), return_intermediates=True)
    outs = {}
    for tag in ("v1", "new"):
        os.environ["SNRF_MARCH"] = tag
        r = make_renderer(cfg, params)
        outs[tag] = {k: v.cpu() for k, v in r.render(o, d, get_feature=("sam",), debug=True).items()}
        torch.cuda.synchronize()
    os.environ.pop("SNRF_MARCH", None)
    for tag, out in outs.items():
        e, er = out["_edges"], ref["_eu1"]
        rel = ((e - er).abs() / er.abs().clamp_min(1e-6))
        dn, dr = out["_density"], ref["_density"]
        bad = (dn - dr).abs() > 1e-3 + 3e-2 * dr.abs()
        print(f"[march {tag}] edges rel err: median {float(rel.median()):.2e} p99 {float(rel.flatten().kthvalue(int(0.99 * rel.numel())).values):.2e} max {float(rel.max()):.2e};"
              f" density outside band: {float(bad.float().mean()):.4f}; by sample index (x32): {[int(v) for v in bad.float().sum(0).tolist()]}")
        mid_rel = 0.5 * (rel[:, :-1] + rel[:, 1:])
        if bad.any():
            print(f"[march {tag}]   edge rel err at bad samples: median {float(mid_rel[bad].median()):.2e}; at good samples: median {float(mid_rel[~bad].median()):.2e};"
                  f" bad samples with weight > 1e-3: {int((bad & (ref['_weights'] > 1e-3)).sum())} of {int(bad.sum())}")
    e1, e2 = outs["v1"]["_edges"], outs["new"]["_edges"]
    print(f"[march new vs v1] edges: max rel diff {float(((e1 - e2).abs() / e1.abs().clamp_min(1e-6)).max()):.2e}; rgb max diff {float((outs['v1']['rgb'] - outs['new']['rgb']).abs().max()):.2e};"
          f" rays with |rgb diff| > 2/255: {int(((outs['v1']['rgb'] - outs['new']['rgb']).abs().max(-1).values > 2 / 255).sum())} of 1024")


if __name__ == "__main__":
    fns = [globals()[n] for n in sys.argv[1:]] or [march_new_vs_v1]
    for fn in fns:
        try:
            fn()
        except Exception as e:  # keep going: this is a diagnostic
            import traceback

            traceback.print_exc()
            print(f"[diag] {fn.__name__} raised {type(e).__name__}: {e}")
