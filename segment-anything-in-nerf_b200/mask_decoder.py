"""Downstream of the rendered feature map: SAM's prompt encoder + two-way-transformer mask decoder,
fed with the ``sam[fh,fw,256]`` map this package renders instead of SAM's ViT image encoder.

What the reference does with the map (paths under /root/reference):
  ``SAMModel.get_outputs_for_camera_ray_bundle``   samnerf/sam_model.py:485-486,514-527  ``predictor.set_feature`` -> ``generate_masked_img``
  ``generate_masked_img`` / ``get_masked_image``    samnerf/sam_utils.py:28-54
  ``SamPredictor.set_feature / predict / predict_torch``  samnerf/segment_anything/predictor.py:100-127,128-277
  ``ResizeLongestSide.apply_coords``                samnerf/segment_anything/utils/transforms.py:33-43
  ``PromptEncoder`` (points, no boxes, no mask input)  samnerf/segment_anything/modeling/prompt_encoder.py
  ``MaskDecoder`` + ``TwoWayTransformer``            samnerf/segment_anything/modeling/mask_decoder.py, transformer.py
  ``Sam.postprocess_masks``                         samnerf/segment_anything/modeling/sam.py:133-162

This is host-side glue around a 4 M-parameter network that runs once per frame on a 64 x 64 map (a few GFLOP, not the
hot path), written in plain torch: device memory and library kernels, no custom CUDA.  The module tree mirrors the
reference's parameter names so that ``load_state_dict(strict=True)`` accepts the ``prompt_encoder.*`` / ``mask_decoder.*``
entries of a SAM checkpoint (``sam_vit_h_4b8939.pth``) unchanged; pinned by ``tests/golden/sam_decoder.npz`` (a
reduced-width instance of the reference's own modules, weights + inputs + outputs) and, in the build container, against
the reference's modules at full width (``tests/test_mask_decoder.py``).  The prompt encoder implements what the reference
calls it with - point prompts; box and mask prompts raise.

``ClipSegDecoder`` is the part of ``CLIPDensePredT`` (samnerf/clipseg/models/clipseg.py:301-500) that runs behind the
rendered 32 x 32 x 192 ClipSeg map - the reference's own ``inp_feature`` branch (:453-497): FiLM conditioning, three
transformer blocks over the rendered activations, the 16 x 16 transposed convolution - with the same parameter names, so
``rd64-uni.pth`` loads.  The text prompt enters as its CLIP embedding (``cond [1,512]``, which the reference's
``get_cond_vec`` accepts as well, :238-239); the CLIP text tower itself is OpenAI's ``clip`` package and not part of this.
"""
from __future__ import annotations

import math
from typing import Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from .prompts import pad_feature_map, predictor_input_size


class _LayerNorm2d(nn.Module):
    """Channel-wise layer norm of an NCHW tensor (modeling/common.py:31-43); parameters ``weight`` / ``bias``."""

    def __init__(self, channels: int, eps: float = 1e-6):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(channels))
        self.bias = nn.Parameter(torch.zeros(channels))
        self.eps = eps

    def forward(self, x):
        mu = x.mean(1, keepdim=True)
        var = (x - mu).pow(2).mean(1, keepdim=True)
        return self.weight[:, None, None] * ((x - mu) / torch.sqrt(var + self.eps)) + self.bias[:, None, None]


class _RandomFourierPE(nn.Module):
    """Positional encoding with random spatial frequencies (prompt_encoder.py:171-214): ``[sin, cos](2 pi (2 c - 1) G)``."""

    def __init__(self, num_pos_feats: int, scale: float = 1.0):
        super().__init__()
        self.register_buffer("positional_encoding_gaussian_matrix", scale * torch.randn(2, num_pos_feats))

    def encode(self, unit_coords: torch.Tensor) -> torch.Tensor:  # coords in [0,1], last dim (x, y)
        c = (2.0 * unit_coords - 1.0) @ self.positional_encoding_gaussian_matrix
        c = 2.0 * math.pi * c
        return torch.cat([torch.sin(c), torch.cos(c)], dim=-1)

    def grid(self, h: int, w: int) -> torch.Tensor:  # [C, h, w], cell centres
        dev = self.positional_encoding_gaussian_matrix.device
        ys = (torch.arange(h, device=dev, dtype=torch.float32) + 0.5) / h
        xs = (torch.arange(w, device=dev, dtype=torch.float32) + 0.5) / w
        yy, xx = torch.meshgrid(ys, xs, indexing="ij")
        return self.encode(torch.stack([xx, yy], dim=-1)).permute(2, 0, 1)


class PointPromptEncoder(nn.Module):
    """``PromptEncoder`` for point prompts.  The mask-input branch (``mask_downscaling``) is constructed so that a SAM
    checkpoint loads strictly, but the reference never feeds it on this path and neither does this class."""

    def __init__(self, embed_dim: int = 256, image_embedding_size: Tuple[int, int] = (64, 64),
                 input_image_size: Tuple[int, int] = (1024, 1024), mask_in_chans: int = 16):
        super().__init__()
        self.embed_dim, self.image_embedding_size, self.input_image_size = embed_dim, image_embedding_size, input_image_size
        self.pe_layer = _RandomFourierPE(embed_dim // 2)
        self.point_embeddings = nn.ModuleList([nn.Embedding(1, embed_dim) for _ in range(4)])  # neg, pos, 2 box corners
        self.not_a_point_embed = nn.Embedding(1, embed_dim)
        self.mask_downscaling = nn.Sequential(
            nn.Conv2d(1, mask_in_chans // 4, kernel_size=2, stride=2), _LayerNorm2d(mask_in_chans // 4), nn.GELU(),
            nn.Conv2d(mask_in_chans // 4, mask_in_chans, kernel_size=2, stride=2), _LayerNorm2d(mask_in_chans), nn.GELU(),
            nn.Conv2d(mask_in_chans, embed_dim, kernel_size=1))
        self.no_mask_embed = nn.Embedding(1, embed_dim)

    def get_dense_pe(self) -> torch.Tensor:
        return self.pe_layer.grid(*self.image_embedding_size).unsqueeze(0)

    def forward(self, points: Tuple[torch.Tensor, torch.Tensor], boxes=None, masks=None):
        """``points = (coords [B,N,2] in input-frame pixels, labels [B,N] in {1 fg, 0 bg})`` ->
        (sparse ``[B,N+1,C]``, dense ``[B,C,H,W]``).  One padding point (label -1) is appended, as the reference does
        whenever no box is given (prompt_encoder.py:77-84)."""
        if boxes is not None or masks is not None:
            raise NotImplementedError("box / mask prompts are not on the reference's NeRF path (sam_utils.py:45-51)")
        coords, labels = points
        b = coords.shape[0]
        coords = torch.cat([coords + 0.5, torch.zeros(b, 1, 2, device=coords.device, dtype=coords.dtype)], dim=1)
        labels = torch.cat([labels, -torch.ones(b, 1, device=labels.device, dtype=labels.dtype)], dim=1)
        size_xy = torch.tensor([self.input_image_size[1], self.input_image_size[0]], device=coords.device, dtype=torch.float32)
        emb = self.pe_layer.encode(coords.to(torch.float32) / size_xy)
        pad, neg, pos = (labels == -1)[..., None], (labels == 0)[..., None], (labels == 1)[..., None]
        emb = torch.where(pad, torch.zeros_like(emb), emb)
        emb = emb + pad * self.not_a_point_embed.weight + neg * self.point_embeddings[0].weight + pos * self.point_embeddings[1].weight
        h, w = self.image_embedding_size
        dense = self.no_mask_embed.weight.reshape(1, -1, 1, 1).expand(b, -1, h, w)
        return emb, dense


class _Attention(nn.Module):
    """Multi-head attention whose projections may shrink the width by ``downsample_rate`` (transformer.py:185-240)."""

    def __init__(self, dim: int, heads: int, downsample_rate: int = 1):
        super().__init__()
        inner = dim // downsample_rate
        assert inner % heads == 0
        self.heads = heads
        self.q_proj, self.k_proj, self.v_proj = nn.Linear(dim, inner), nn.Linear(dim, inner), nn.Linear(dim, inner)
        self.out_proj = nn.Linear(inner, dim)

    def forward(self, q, k, v):
        def split(x):
            b, n, c = x.shape
            return x.reshape(b, n, self.heads, c // self.heads).transpose(1, 2)

        q, k, v = split(self.q_proj(q)), split(self.k_proj(k)), split(self.v_proj(v))
        a = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(q.shape[-1]), dim=-1)
        o = (a @ v).transpose(1, 2)
        return self.out_proj(o.reshape(o.shape[0], o.shape[1], -1))


class _Mlp(nn.Module):
    def __init__(self, dim: int, hidden: int):
        super().__init__()
        self.lin1, self.lin2 = nn.Linear(dim, hidden), nn.Linear(hidden, dim)

    def forward(self, x):
        return self.lin2(F.relu(self.lin1(x)))


class _TwoWayBlock(nn.Module):
    """Token self-attention, tokens -> image, token MLP, image -> tokens (transformer.py:109-183)."""

    def __init__(self, dim: int, heads: int, mlp_dim: int, downsample_rate: int, skip_first_layer_pe: bool):
        super().__init__()
        self.self_attn = _Attention(dim, heads)
        self.norm1 = nn.LayerNorm(dim)
        self.cross_attn_token_to_image = _Attention(dim, heads, downsample_rate)
        self.norm2 = nn.LayerNorm(dim)
        self.mlp = _Mlp(dim, mlp_dim)
        self.norm3 = nn.LayerNorm(dim)
        self.norm4 = nn.LayerNorm(dim)
        self.cross_attn_image_to_token = _Attention(dim, heads, downsample_rate)
        self.skip_first_layer_pe = skip_first_layer_pe

    def forward(self, tokens, image, token_pe, image_pe):
        if self.skip_first_layer_pe:
            tokens = self.self_attn(tokens, tokens, tokens)
        else:
            q = tokens + token_pe
            tokens = tokens + self.self_attn(q, q, tokens)
        tokens = self.norm1(tokens)
        tokens = self.norm2(tokens + self.cross_attn_token_to_image(tokens + token_pe, image + image_pe, image))
        tokens = self.norm3(tokens + self.mlp(tokens))
        image = self.norm4(image + self.cross_attn_image_to_token(image + image_pe, tokens + token_pe, tokens))
        return tokens, image


class TwoWayTransformer(nn.Module):
    def __init__(self, depth: int = 2, embedding_dim: int = 256, num_heads: int = 8, mlp_dim: int = 2048,
                 attention_downsample_rate: int = 2):
        super().__init__()
        self.layers = nn.ModuleList([_TwoWayBlock(embedding_dim, num_heads, mlp_dim, attention_downsample_rate, i == 0)
                                     for i in range(depth)])
        self.final_attn_token_to_image = _Attention(embedding_dim, num_heads, attention_downsample_rate)
        self.norm_final_attn = nn.LayerNorm(embedding_dim)

    def forward(self, image_embedding, image_pe, point_embedding):
        image = image_embedding.flatten(2).permute(0, 2, 1)  # [B, HW, C]
        pe = image_pe.flatten(2).permute(0, 2, 1)
        tokens = point_embedding
        for layer in self.layers:
            tokens, image = layer(tokens, image, point_embedding, pe)
        tokens = self.norm_final_attn(tokens + self.final_attn_token_to_image(tokens + point_embedding, image + pe, image))
        return tokens, image


class _Head(nn.Module):
    """``MLP`` of mask_decoder.py:154-176: ReLU between ``layers``, none after the last."""

    def __init__(self, d_in: int, hidden: int, d_out: int, n_layers: int):
        super().__init__()
        dims = [d_in] + [hidden] * (n_layers - 1) + [d_out]
        self.layers = nn.ModuleList([nn.Linear(a, b) for a, b in zip(dims[:-1], dims[1:])])

    def forward(self, x):
        for i, layer in enumerate(self.layers):
            x = layer(x) if i == len(self.layers) - 1 else F.relu(layer(x))
        return x


class MaskDecoder(nn.Module):
    def __init__(self, transformer_dim: int = 256, transformer: Optional[nn.Module] = None, num_multimask_outputs: int = 3,
                 iou_head_depth: int = 3, iou_head_hidden_dim: int = 256):
        super().__init__()
        d = transformer_dim
        self.transformer = transformer if transformer is not None else TwoWayTransformer(embedding_dim=d)
        self.num_mask_tokens = num_multimask_outputs + 1
        self.iou_token = nn.Embedding(1, d)
        self.mask_tokens = nn.Embedding(self.num_mask_tokens, d)
        self.output_upscaling = nn.Sequential(nn.ConvTranspose2d(d, d // 4, kernel_size=2, stride=2), _LayerNorm2d(d // 4), nn.GELU(),
                                              nn.ConvTranspose2d(d // 4, d // 8, kernel_size=2, stride=2), nn.GELU())
        self.output_hypernetworks_mlps = nn.ModuleList([_Head(d, d, d // 8, 3) for _ in range(self.num_mask_tokens)])
        self.iou_prediction_head = _Head(d, iou_head_hidden_dim, self.num_mask_tokens, iou_head_depth)

    def forward(self, image_embeddings, image_pe, sparse_prompt_embeddings, dense_prompt_embeddings, multimask_output: bool):
        """-> (low-resolution mask logits ``[B, 1 or 3, 4H, 4W]``, predicted IoU ``[B, 1 or 3]``)."""
        b = sparse_prompt_embeddings.shape[0]
        out_tokens = torch.cat([self.iou_token.weight, self.mask_tokens.weight], dim=0)
        tokens = torch.cat([out_tokens.unsqueeze(0).expand(b, -1, -1), sparse_prompt_embeddings], dim=1)
        src = torch.repeat_interleave(image_embeddings, b, dim=0) + dense_prompt_embeddings
        pos = torch.repeat_interleave(image_pe, b, dim=0)
        _, c, h, w = src.shape
        hs, src = self.transformer(src, pos, tokens)
        iou_out, mask_out = hs[:, 0], hs[:, 1:1 + self.num_mask_tokens]
        up = self.output_upscaling(src.transpose(1, 2).reshape(b, c, h, w))
        hyper = torch.stack([mlp(mask_out[:, i]) for i, mlp in enumerate(self.output_hypernetworks_mlps)], dim=1)
        masks = (hyper @ up.flatten(2)).reshape(b, -1, up.shape[-2], up.shape[-1])
        iou = self.iou_prediction_head(iou_out)
        keep = slice(1, None) if multimask_output else slice(0, 1)
        return masks[:, keep], iou[:, keep]


class SamMaskPredictor(nn.Module):
    """``SamPredictor`` with a rendered feature map in place of the image encoder: ``set_feature`` then ``predict``.

    ``img_size`` is SAM's input frame (1024); the rendered map's long side must equal the decoder's embedding size
    (``config.get_feature_size``: 64)."""

    mask_threshold = 0.0

    def __init__(self, prompt_encoder: Optional[PointPromptEncoder] = None, mask_decoder: Optional[MaskDecoder] = None,
                 img_size: int = 1024):
        super().__init__()
        self.prompt_encoder = prompt_encoder if prompt_encoder is not None else PointPromptEncoder()
        self.mask_decoder = mask_decoder if mask_decoder is not None else MaskDecoder()
        self.img_size = img_size
        self.features: Optional[torch.Tensor] = None
        self.original_size = self.input_size = None

    @classmethod
    def from_sam_checkpoint(cls, path_or_state, device="cpu") -> "SamMaskPredictor":
        """Load the ``prompt_encoder.*`` / ``mask_decoder.*`` entries of a SAM checkpoint (``build_sam.py:103-106``);
        the image-encoder weights in the file are ignored - the NeRF renders that map."""
        state = torch.load(path_or_state, map_location="cpu") if isinstance(path_or_state, (str, bytes)) else path_or_state
        me = cls()
        own = {k: v for k, v in state.items() if k.startswith(("prompt_encoder.", "mask_decoder."))}
        me.load_state_dict(own, strict=True)
        return me.to(device).eval()

    @property
    def device(self):
        return self.mask_decoder.iou_token.weight.device

    @torch.no_grad()
    def set_feature(self, feature: torch.Tensor, original_image_size: Tuple[int, int]) -> None:
        """``feature``: the rendered map, ``[fh,fw,C]`` as this package returns it or ``[C,fh,fw]`` as the reference's
        call site passes it (sam_model.py:486) - told apart by which end holds the decoder's channel count."""
        c = self.prompt_encoder.embed_dim
        if feature.shape[0] == c and feature.shape[-1] != c:
            feature = feature.permute(1, 2, 0)
        self.original_size = tuple(int(v) for v in original_image_size)
        self.input_size = predictor_input_size(self.original_size, self.img_size)
        self.features = pad_feature_map(feature.to(self.device, torch.float32))

    @torch.no_grad()
    def predict(self, point_coords, point_labels, multimask_output: bool = False, return_logits: bool = False):
        """``point_coords [N,2]`` (x, y) in pixels of the original image, ``point_labels [N]`` ->
        (masks ``[1,K,H,W]`` bool (or logits), iou ``[1,K]``, low-res logits ``[1,K,4S,4S]``) - ``return_torch=True``
        of the reference's ``predict``."""
        if self.features is None:
            raise RuntimeError("set_feature(...) must be called before predict")
        pts = torch.as_tensor(np.asarray(point_coords, dtype=np.float64) if not torch.is_tensor(point_coords) else point_coords,
                              dtype=torch.float64).reshape(-1, 2).clone()
        (oh, ow), (nh, nw) = self.original_size, _preprocess_shape(self.original_size, self.img_size)
        pts[:, 0] *= nw / ow  # ResizeLongestSide.apply_coords: float64 arithmetic, then float32
        pts[:, 1] *= nh / oh
        coords = pts.to(self.device, torch.float32)[None]
        labels = torch.as_tensor(point_labels, dtype=torch.int32, device=self.device).reshape(1, -1)
        sparse, dense = self.prompt_encoder((coords, labels))
        low, iou = self.mask_decoder(self.features, self.prompt_encoder.get_dense_pe(), sparse, dense, multimask_output)
        masks = F.interpolate(low, (self.img_size, self.img_size), mode="bilinear", align_corners=False)
        masks = masks[..., : self.input_size[0], : self.input_size[1]]
        masks = F.interpolate(masks, self.original_size, mode="bilinear", align_corners=False)
        return (masks if return_logits else masks > self.mask_threshold), iou, low


def _preprocess_shape(original_size: Tuple[int, int], long_side: int) -> Tuple[int, int]:
    """``ResizeLongestSide.get_preprocess_shape`` (transforms.py:86-92): round-half-up of the scaled sides."""
    h, w = original_size
    scale = long_side * 1.0 / max(h, w)
    return int(h * scale + 0.5), int(w * scale + 0.5)


MASK_RGBA = (30 / 255, 144 / 255, 255 / 255, 0.6)  # the fixed colour of sam_utils.show_mask_tensor


def masked_image(mask: torch.Tensor, image: torch.Tensor, rgba: Sequence[float] = MASK_RGBA) -> torch.Tensor:
    """``get_masked_image`` (sam_utils.py:37-42): blend ``rgba`` over ``image [H,W,3]`` where ``mask [H,W]`` is set.
    The reference draws a random colour per call; the fixed one is the default here so that frames are reproducible."""
    col = torch.tensor(rgba, device=image.device, dtype=image.dtype)
    m = mask.reshape(*mask.shape[-2:], 1).to(image.dtype)
    return (m * col[:3]) * (m * col[3]) + image * (1 - m * col[3])


def generate_masked_img(predictor: SamMaskPredictor, points, labels, image: torch.Tensor, rgba: Sequence[float] = MASK_RGBA):
    """``generate_masked_img`` (sam_utils.py:45-54): one mask for all prompts, blended over the rendered rgb."""
    masks, _, _ = predictor.predict(points, labels, multimask_output=False)
    return masked_image(masks[0, 0], image, rgba)


class ClipSegDecoder(nn.Module):
    """``CLIPDensePredT(version="ViT-B/16", reduce_dim=64)`` (sam_model.py:216) without its CLIP towers: what
    ``self.clipseg(None, inp_feature=..., conditional=...)`` computes (sam_model.py:487-499).  ``reduce`` / ``reduces`` are
    part of the checkpoint layout (the rendered activations are already reduced, so they are not applied here either)."""

    def __init__(self, reduce_dim: int = 64, n_heads: int = 4, depth: int = 3, cond_dim: int = 512, clip_width: int = 768,
                 patch: int = 16, cond_layer: int = 0):
        super().__init__()
        self.film_mul, self.film_add = nn.Linear(cond_dim, reduce_dim), nn.Linear(cond_dim, reduce_dim)
        self.reduce = nn.Linear(clip_width, reduce_dim)
        self.reduces = nn.ModuleList([nn.Linear(clip_width, reduce_dim) for _ in range(depth)])
        self.blocks = nn.ModuleList([nn.TransformerEncoderLayer(d_model=reduce_dim, nhead=n_heads) for _ in range(depth)])
        self.trans_conv = nn.ConvTranspose2d(reduce_dim, 1, (patch, patch), stride=(patch, patch))
        self.cond_layer = cond_layer

    def forward(self, activations: Sequence[torch.Tensor], cond: torch.Tensor) -> torch.Tensor:
        """``activations``: three ``[1 + T, 1, reduce_dim]`` token tensors (``prompts.clipseg_activations`` of the rendered
        map, CLS slot first); ``cond [1, cond_dim]``.  Returns the ``[sqrt(T) * patch, sqrt(T) * patch]`` logit map."""
        a = None
        for i, (act, block) in enumerate(zip(activations, self.blocks)):
            a = act if a is None else act + a
            if i == self.cond_layer:
                a = self.film_mul(cond) * a + self.film_add(cond)
            a = block(a)
        a = a[1:].permute(1, 2, 0)  # drop the CLS token -> [1, C, T]
        side = int(math.sqrt(a.shape[2]))
        return self.trans_conv(a.reshape(1, a.shape[1], side, side))[0, 0]

    @classmethod
    def from_checkpoint(cls, path_or_state, device="cpu") -> "ClipSegDecoder":
        """``rd64-uni.pth`` (sam_model.py:217-220, loaded non-strictly over the CLIP-carrying model in the reference): here
        every decoder key must be present, and keys of the CLIP towers, if any, are ignored."""
        state = torch.load(path_or_state, map_location="cpu") if isinstance(path_or_state, (str, bytes)) else path_or_state
        me = cls()
        own = set(me.state_dict())
        me.load_state_dict({k: v for k, v in state.items() if k in own}, strict=True)
        return me.to(device).eval()


def clipseg_heat_and_clicks(decoder: ClipSegDecoder, clipseg_map: torch.Tensor, cond: torch.Tensor, image_width: int,
                            image_height: int):
    """Rendered ``clipseg[32,32,192]`` + text embedding -> (``clipseg_feature [512,512,1]``, click prompts ``[n,2]``):
    sam_model.py:487-512."""
    from .prompts import clipseg_activations

    dev = decoder.trans_conv.weight.device
    with torch.no_grad():
        heat = decoder(clipseg_activations(clipseg_map.to(dev, torch.float32)), cond.to(dev, torch.float32)).sigmoid()
    return heat.unsqueeze(dim=-1), clipseg_click_points(heat, image_width, image_height)


def clipseg_click_points(heat: torch.Tensor, image_width: int, image_height: int, k: int = 1000, thresh: float = 0.7,
                         cell: int = 16) -> np.ndarray:
    """ClipSeg heat map ``[H,W]`` (after the sigmoid) -> click prompts (sam_model.py:496-512): average over ``cell x cell``
    blocks, take the ``k`` hottest blocks, keep those above ``thresh``, scale block indices to image pixels.  Returns
    ``[n,2]`` float32 (x, y); empty when nothing passes."""
    h, w = heat.shape
    blocks = heat.reshape(h // cell, cell, w // cell, cell).permute(0, 2, 1, 3).reshape(h // cell, w // cell, -1).mean(dim=-1)
    fh, fw = blocks.shape
    top = blocks.flatten().topk(k=min(k, fh * fw))[1]
    xs, ys = (top % fw).long(), (top // fw).long()
    keep = blocks[ys, xs] > thresh
    pts = torch.stack([xs, ys], dim=1)[keep].cpu().numpy().astype(np.float32)
    pts[..., 0] = pts[..., 0] / fw * image_width
    pts[..., 1] = pts[..., 1] / fh * image_height
    return pts
