"""Generate ``tests/golden/state_dict_layout.json``: the key names and shapes of the REFERENCE'S OWN modules'
``state_dict()`` for the hot path, wrapped the way the reference's trainer writes them into a checkpoint.

TEST INFRASTRUCTURE ONLY.  Runs only in the build container (needs /root/reference):

    python -m oracle.make_ckpt_golden

What executes, unmodified, from /root/reference: ``TCNNNerfactoField``, ``HashMLPDensityField``, ``SAMField`` (their
constructors register the ``params`` tensors whose names a checkpoint carries), wired like
``NerfactoModel.populate_modules`` / ``SAMModel.populate_modules`` by ``oracle.make_golden.build_reference_model``.
The container layout (``{"step", "pipeline", "optimizers", "scalers"}``, ``_model.`` prefix, ``module.`` under DDP)
is restated from ``nerfstudio/engine/trainer.py:389-400`` and ``nerfstudio/pipelines/base_pipeline.py:109-115,373``.
Only names and shapes are stored (the conv head alone is 4.7 MB of fp32); ``tests/test_checkpoint.py`` fills them
with seeded values.
"""
from __future__ import annotations

import json
import os

import torch

from oracle.make_golden import GOLDEN, build_reference_model, load_reference


def main():
    from samnerf_b200 import SAMNeRFConfig, make_synthetic_params

    ref = load_reference()
    layout = {}
    for name, cfg in (
        ("tiny_distill_clipseg_p4", SAMNeRFConfig.tiny(clipseg=True, patch_size=4)),
        ("full_distill_p4", SAMNeRFConfig.distill(clipseg=False, patch_size=4)),
    ):
        params = make_synthetic_params(cfg, "init", 1)
        model = build_reference_model(ref, cfg, params)
        sd = model.state_dict()
        layout[name] = {"_model." + k: list(v.shape) for k, v in sd.items()}
        # what a real pipeline adds around the model (base_pipeline.py:109-115 splits on the "_model." prefix);
        # the camera optimiser is the one non-model tensor a VanillaPipeline state_dict carries
        layout[name]["datamanager.train_camera_optimizer.pose_adjustment"] = [2, 6]
        print(name, len(sd), "model tensors;", sum(v.numel() for v in sd.values()), "values")
    path = os.path.join(GOLDEN, "state_dict_layout.json")
    with open(path, "w") as f:
        json.dump(layout, f, indent=1, sort_keys=True)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    torch.manual_seed(0)
    main()
