"""Generate ``tests/golden/sampler_training.npz``: the REFERENCE'S OWN ``ProposalNetworkSampler`` run in training
mode (stratified single jitter, ray_samplers.py:104-112,314-322) with its ``torch.rand`` draws recorded, so that the
oracle's and the kernel's training-mode sampling can be pinned to it.

TEST INFRASTRUCTURE ONLY.  Runs only in the build container (needs /root/reference):

    python -m oracle.make_jitter_golden
"""
from __future__ import annotations

import os

import numpy as np
import torch

from oracle.make_golden import GOLDEN, build_reference_model, load_reference, make_bundle, params_checksum


def main():
    from samnerf_b200 import SAMNeRFConfig, make_synthetic_params
    from samnerf_b200.synthetic import plumbing_rays

    ref = load_reference()
    cfg = SAMNeRFConfig.tiny(clipseg=False, patch_size=1)
    params = make_synthetic_params(cfg, "scene", 8)
    model = build_reference_model(ref, cfg, params)
    o, d = plumbing_rays()
    idx = torch.arange(160) * 51 + 3
    o, d = o[idx].contiguous(), d[idx].contiguous()
    bundle = model.collider(make_bundle(ref, o, d))  # eval collider: near 0 (the model object is in eval mode)
    sampler = model.proposal_sampler
    sampler.train()                  # stratified sampling on (train_stratified defaults to True)
    sampler._step, sampler._steps_since_update = 0, 0  # "updated" branch: density with grad enabled (step < 10)
    draws = []
    real_rand = torch.rand

    def recording_rand(*a, **k):
        t = real_rand(*a, **k)
        draws.append(t.clone())
        return t

    torch.manual_seed(123)
    torch.rand = recording_rand
    try:
        ray_samples, weights_list, ray_samples_list = sampler(bundle, density_fns=model.density_fns)
    finally:
        torch.rand = real_rand
    assert len(draws) == 2 and draws[0].shape == (160, 1) and draws[1].shape == (160, 1), [x.shape for x in draws]
    rs0 = ray_samples_list[0]
    out = {
        "_origins": o.numpy(), "_directions": d.numpy(), "_params_checksum": np.array(params_checksum(params)),
        "jitter": torch.cat(draws, dim=-1).numpy(),
        "edges0": torch.cat([rs0.frustums.starts[..., 0], rs0.frustums.ends[:, -1:, 0]], -1).detach().numpy(),
        "w0": weights_list[0][..., 0].detach().numpy(),
        "edges1": torch.cat([ray_samples.frustums.starts[..., 0], ray_samples.frustums.ends[:, -1:, 0]], -1).detach().numpy(),
        "spacing1": torch.cat([ray_samples.spacing_starts[..., 0], ray_samples.spacing_ends[:, -1:, 0]], -1).detach().numpy(),
    }
    # second pass with annealed proposal weights (set_anneal, ray_samplers.py:546-548,583; nerfacto.py:248-256)
    draws.clear()
    sampler.set_anneal(0.37)
    torch.rand = recording_rand
    try:
        rs2, wl2, _ = sampler(bundle, density_fns=model.density_fns)
    finally:
        torch.rand = real_rand
    out["anneal"] = np.array(0.37)
    out["jitter_anneal"] = torch.cat(draws, dim=-1).numpy()
    out["edges1_anneal"] = torch.cat([rs2.frustums.starts[..., 0], rs2.frustums.ends[:, -1:, 0]], -1).detach().numpy()
    out["w0_anneal"] = wl2[0][..., 0].detach().numpy()
    path = os.path.join(GOLDEN, "sampler_training.npz")
    np.savez_compressed(path, **out)
    print({k: v.shape for k, v in out.items()}, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
