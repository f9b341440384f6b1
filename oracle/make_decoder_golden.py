"""Generate ``tests/golden/sam_decoder.npz`` and ``tests/golden/sam_decoder_layout.json`` with the REFERENCE'S OWN
``PromptEncoder`` / ``TwoWayTransformer`` / ``MaskDecoder`` / ``Sam`` / ``SamPredictor``
(samnerf/segment_anything/modeling/*.py, predictor.py - imported unmodified from /root/reference) and with the click
extraction of ``SAMModel.get_outputs_for_camera_ray_bundle`` (samnerf/sam_model.py:501-512, compiled out of the file).

* ``sam_decoder.npz``: a reduced-width instance of the reference's modules (embedding 32, 8 x 8 map, input frame 128) with
  seeded random weights - every tensor of its ``state_dict()``, a rendered-feature-map stand-in, click prompts, and what
  ``SamPredictor.set_feature`` + ``predict(..., return_torch=True)`` return for them (single- and multi-mask).
* ``sam_decoder_layout.json``: key names and shapes of ``prompt_encoder.*`` / ``mask_decoder.*`` at SAM's real widths
  (``build_sam.py:46-106``), which is what ``SamMaskPredictor.from_sam_checkpoint`` has to accept strictly.

TEST INFRASTRUCTURE ONLY.  Runs only in the build container (needs /root/reference):

    python -m oracle.make_decoder_golden
"""
from __future__ import annotations

import json
import os
import sys
import textwrap

import numpy as np
import torch

from oracle.make_golden import GOLDEN, REF

TINY = dict(embed=32, grid=8, img=128, mask_in_chans=8, depth=2, mlp_dim=64, heads=4, iou_hidden=32)


def reference_modules():
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from samnerf.segment_anything.modeling import MaskDecoder, PromptEncoder, Sam, TwoWayTransformer
    from samnerf.segment_anything.predictor import SamPredictor

    return MaskDecoder, PromptEncoder, Sam, TwoWayTransformer, SamPredictor


def build_reference(embed, grid, img, mask_in_chans, depth, mlp_dim, heads, iou_hidden, seed=0):
    MaskDecoder, PromptEncoder, Sam, TwoWayTransformer, SamPredictor = reference_modules()
    torch.manual_seed(seed)

    class _NoEncoder(torch.nn.Module):  # the NeRF renders the map; only img_size is read (predictor.py:100-127)
        img_size = img

    sam = Sam(image_encoder=_NoEncoder(),
              prompt_encoder=PromptEncoder(embed_dim=embed, image_embedding_size=(grid, grid), input_image_size=(img, img),
                                           mask_in_chans=mask_in_chans),
              mask_decoder=MaskDecoder(num_multimask_outputs=3,
                                       transformer=TwoWayTransformer(depth=depth, embedding_dim=embed, mlp_dim=mlp_dim, num_heads=heads),
                                       transformer_dim=embed, iou_head_depth=3, iou_head_hidden_dim=iou_hidden))
    # default inits leave the embeddings at N(0,1) and the LayerNorms at identity; perturb everything so that no
    # parameter can be dropped or swapped unnoticed
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for p in sam.parameters():
            p.add_(0.05 * torch.randn(p.shape, generator=g))
    return sam.eval(), SamPredictor(sam)


def click_points_reference(heat, image_width, image_height):
    """Lines 501-512 of samnerf/sam_model.py, executed as they stand."""
    from einops import rearrange

    lines = open(os.path.join(REF, "samnerf", "sam_model.py")).read().splitlines()
    first = next(i for i, l in enumerate(lines) if "clip_feature = rearrange(" in l)
    last = next(i for i, l in enumerate(lines) if "clip_points[..., 1] = clip_points[..., 1] / _fh * image_height" in l)
    ns = dict(torch=torch, np=np, rearrange=rearrange, clip_feature=heat, image_width=image_width, image_height=image_height,
              clip_points=np.zeros((0, 2), np.float32))
    exec(textwrap.dedent("\n".join(lines[first:last + 1])), ns)
    return ns["clip_points"]


def build_reference_clipseg(seed=0):
    """The reference's ``CLIPDensePredT(version="ViT-B/16", reduce_dim=64)`` (sam_model.py:216) with OpenAI's ``clip``
    package - absent here - stubbed out: its CLIP towers are only touched by the image / text branches, never by the
    ``inp_feature`` branch with a tensor conditional that is pinned here."""
    import types

    if REF not in sys.path:
        sys.path.insert(0, REF)
    if "clip" not in sys.modules:
        class _NoClip(torch.nn.Module):
            def __init__(self):
                super().__init__()
                self.visual = torch.nn.Identity()

        stub = types.ModuleType("clip")
        stub.load = lambda version, device="cpu", jit=False: (_NoClip(), None)
        stub.tokenize = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("no CLIP text tower in this container"))
        sys.modules["clip"] = stub
    from samnerf.clipseg.models.clipseg import CLIPDensePredT

    model = CLIPDensePredT(version="ViT-B/16", reduce_dim=64)
    model.load_state_dict(seeded_state({k: tuple(v.shape) for k, v in model.state_dict().items()}, seed), strict=True)
    return model.eval()


def seeded_state(shapes, seed: int):
    """A state dict drawn from one seeded generator in sorted-key order (weights of the ClipSeg fixture: 1.1 M values that
    the test regenerates instead of storing; layer norms near identity, everything else small and non-zero)."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for k in sorted(shapes):
        t = 0.08 * torch.randn(shapes[k], generator=g)
        out[k] = t + 1.0 if (k.endswith("norm1.weight") or k.endswith("norm2.weight")) else t
    return out


def clipseg_inputs():
    """Rendered-ClipSeg-map stand-in ``[32,32,192]`` and text-embedding stand-in ``[1,512]`` of the fixture (seeded)."""
    g = torch.Generator().manual_seed(31)
    return 0.3 * torch.randn(32, 32, 192, generator=g), torch.randn(1, 512, generator=g)


def click_heat() -> torch.Tensor:
    """The heat map of the click fixture (the test regenerates it from the same seed instead of storing 1 MB):
    512 / 16 = 32 x 32 blocks - the reference's topk(k=1000) needs at least 1000 of them."""
    g = torch.Generator().manual_seed(11)
    coarse = torch.rand(32, 32, generator=g).repeat_interleave(16, 0).repeat_interleave(16, 1)  # block means on both sides of 0.7
    return 0.8 * coarse + 0.2 * torch.rand(512, 512, generator=g)


def main():
    out = {}
    sam, pred = build_reference(**TINY)
    for k, v in sam.state_dict().items():
        if k.startswith(("prompt_encoder.", "mask_decoder.")):
            out["w." + k] = v.numpy()
    g = torch.Generator().manual_seed(5)
    for name, (fh, fw, size) in {"landscape": (5, 8, (100, 160)), "square": (8, 8, (96, 96))}.items():
        feat = torch.randn(fh, fw, TINY["embed"], generator=g)
        pts = np.array([[30.2, 40.7], [0.0, 0.0], [size[1] - 1.0, size[0] - 1.0], [77.5, 12.25]], np.float64)
        lab = np.array([1, 1, 0, 1])
        pred.set_feature(feat.permute(2, 0, 1), original_image_size=size)
        out[f"{name}.feat"], out[f"{name}.points"], out[f"{name}.labels"], out[f"{name}.size"] = feat.numpy(), pts, lab, np.array(size)
        for multi in (False, True):
            masks, iou, low = pred.predict(point_coords=pts, point_labels=lab, multimask_output=multi, return_logits=True, return_torch=True)
            tag = f"{name}.{'multi' if multi else 'single'}"
            out[tag + ".logits"], out[tag + ".iou"], out[tag + ".low"] = masks.numpy(), iou.numpy(), low.numpy()
    out["clicks.points"] = click_points_reference(click_heat(), 1297, 840)
    # ClipSeg decoder: the reference's module on a rendered-map stand-in, through the call of sam_model.py:487-499
    clipseg = build_reference_clipseg(seed=21)  # the test rebuilds these weights with seeded_state(..., 21)
    cmap, cond = clipseg_inputs()
    acts = []
    for _i in range(3):
        _c = cmap[..., 64 * _i: 64 * (_i + 1)].reshape(-1, 64).unsqueeze(dim=1)
        acts.append(torch.cat([_c.mean(dim=0, keepdim=True), _c], dim=0))
    with torch.no_grad():
        logits = clipseg(None, inp_feature={"activations": acts, "visual_q": None, "transformed_image_size": (32, 32)},
                         conditional=cond)[0][0][0]
    out["clipseg.logits_every_3rd"] = logits[::3, ::3].numpy()  # every token's 16 x 16 block is sampled ~28 times
    path = os.path.join(GOLDEN, "sam_decoder.npz")
    np.savez_compressed(path, **out)
    print(len(out), "arrays,", os.path.getsize(path), "bytes")

    MaskDecoder, PromptEncoder, Sam, TwoWayTransformer, _ = reference_modules()
    full = Sam(image_encoder=torch.nn.Identity(),
               prompt_encoder=PromptEncoder(embed_dim=256, image_embedding_size=(64, 64), input_image_size=(1024, 1024), mask_in_chans=16),
               mask_decoder=MaskDecoder(num_multimask_outputs=3,
                                        transformer=TwoWayTransformer(depth=2, embedding_dim=256, mlp_dim=2048, num_heads=8),
                                        transformer_dim=256, iou_head_depth=3, iou_head_hidden_dim=256))
    layout = {k: list(v.shape) for k, v in full.state_dict().items() if k.startswith(("prompt_encoder.", "mask_decoder."))}
    with open(os.path.join(GOLDEN, "sam_decoder_layout.json"), "w") as f:
        json.dump(layout, f, indent=0, sort_keys=True)
    print(len(layout), "keys,", sum(int(np.prod(s)) for s in layout.values()), "parameters")


if __name__ == "__main__":
    main()
