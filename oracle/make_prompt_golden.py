"""Generate ``tests/golden/prompts.npz`` with the REFERENCE'S OWN ``project`` (samnerf/sam_model.py:95-123) and
``SamPredictor.set_feature`` (samnerf/segment_anything/predictor.py:100-127), compiled out of the files with ``ast``
(the modules themselves need torchmetrics / torchvision extras / a GPU to import).

TEST INFRASTRUCTURE ONLY.  Runs only in the build container (needs /root/reference):

    python -m oracle.make_prompt_golden
"""
from __future__ import annotations

import ast
import math
import os
from types import SimpleNamespace

import numpy as np
import torch

from oracle.make_golden import GOLDEN, REF


def _function_source(path, name, cls=None):
    src = open(path).read()
    for node in ast.parse(src).body:
        if cls is None and isinstance(node, ast.FunctionDef) and node.name == name:
            return ast.get_source_segment(src, node)
        if cls is not None and isinstance(node, ast.ClassDef) and node.name == cls:
            for sub in node.body:
                if isinstance(sub, ast.FunctionDef) and sub.name == name:
                    import textwrap

                    seg = textwrap.dedent("    " + ast.get_source_segment(src, sub))
                    return seg.replace("@torch.no_grad()\n", "")
    raise KeyError(name)


def main():
    from samnerf_b200.synthetic import look_at

    ns = dict(torch=torch, np=np, math=math)
    exec(_function_source(os.path.join(REF, "samnerf", "sam_model.py"), "project"), ns)
    exec(_function_source(os.path.join(REF, "samnerf", "segment_anything", "predictor.py"), "set_feature", "SamPredictor"), ns)
    g = torch.Generator().manual_seed(0)
    c2w = look_at((1.1, 0.6, 0.45))[:3, :4]
    intrin = torch.tensor([[300.0, 0.0, 160.0], [0.0, 310.0, 120.0], [0.0, 0.0, 1.0]])
    pts = torch.randn(64, 3, generator=g) * 0.6
    out = {"c2w": c2w.numpy(), "intrin": intrin.numpy(), "points": pts.numpy(),
           "project": ns["project"](intrin, c2w, pts).numpy()}
    for name, (fh, fw, size) in {"landscape": (43, 64, (1060, 1600)), "square": (64, 64, (800, 800))}.items():
        feat = torch.randn(fh, fw, 4, generator=g)  # the padding is channel-agnostic; 4 channels keep the fixture small
        me = SimpleNamespace(reset_image=lambda: None,
                             model=SimpleNamespace(device="cpu", image_encoder=SimpleNamespace(img_size=1024)))
        ns["set_feature"](me, feat.permute(2, 0, 1), original_image_size=size)
        out[f"{name}.feat"] = feat.numpy()
        out[f"{name}.padded"] = me.features.numpy()
        out[f"{name}.input_size"] = np.array(me.input_size)
        out[f"{name}.original"] = np.array(size)
    path = os.path.join(GOLDEN, "prompts.npz")
    np.savez_compressed(path, **out)
    print({k: v.shape for k, v in out.items()}, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
