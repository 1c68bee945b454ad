"""TEST INFRASTRUCTURE ONLY — CPU restatement (plain PyTorch, fp32 or fp64) of the reference's neck / pixel decoder /
mask head arithmetic (row a8).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.

Pinned against outputs of the reference itself: tests/golden/seg_small.npz is produced by tests/golden/make_golden_seg.py,
which imports the real sam3/model/necks.py and sam3/model/maskformer_segmentation.py (tests/test_seg_oracle.py).

Parameters are a flat dict with the reference's state-dict names:
  neck           convs.{i}.{dconv_2x2_0|dconv_2x2_1|dconv_2x2|conv_1x1|conv_3x3}.{weight,bias}
  pixel decoder  conv_layers.{k}.{weight,bias}, norms.{k}.{weight,bias}
  heads          instance_seg_head.{weight,bias}, semantic_seg_head.{weight,bias}, mask_predictor.mask_embed.layers.{j}.{weight,bias}
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import torch
import torch.nn.functional as F


def _q(x: torch.Tensor, dt):
    """Operand rounding of the CUDA path's precision contract (16-bit GEMM operands, fp32 accumulation), straight-through
    for autograd.  dt=None (the default everywhere) is the reference's exact fp32 arithmetic; tests pass torch.float16 to
    separate ReLU-mask flips caused by operand rounding from arithmetic errors (tests/test_seg_gpu.py)."""
    return x if dt is None else x + (x.to(dt).to(x.dtype) - x).detach()


def neck_branch(x: torch.Tensor, p: Dict[str, torch.Tensor], i: int, scale: float, operand_dtype=None) -> torch.Tensor:
    """One SimpleFPN branch (sam3/model/necks.py:40-92 builds it, :113 runs it)."""
    pre = f"convs.{i}."
    dt = operand_dtype
    w = lambda name: _q(p[pre + name], dt)   # noqa: E731
    x = _q(x, dt)                           # (decides the max-pool arg-max when two candidates round to the same 16-bit value)
    if scale == 4.0:
        x = _q(F.conv_transpose2d(x, w("dconv_2x2_0.weight"), p[pre + "dconv_2x2_0.bias"], stride=2), dt)
        x = _q(F.gelu(x), dt)                                               # nn.GELU() default = exact erf
        x = _q(F.conv_transpose2d(x, w("dconv_2x2_1.weight"), p[pre + "dconv_2x2_1.bias"], stride=2), dt)
    elif scale == 2.0:
        x = _q(F.conv_transpose2d(x, w("dconv_2x2.weight"), p[pre + "dconv_2x2.bias"], stride=2), dt)
    elif scale == 0.5:
        x = F.max_pool2d(x, kernel_size=2, stride=2)
    elif scale != 1.0:
        raise NotImplementedError(scale)
    x = _q(F.conv2d(x, w("conv_1x1.weight"), p[pre + "conv_1x1.bias"]), dt)
    return F.conv2d(x, w("conv_3x3.weight"), p[pre + "conv_3x3.bias"], padding=1)


def neck(x: torch.Tensor, p: Dict[str, torch.Tensor], scale_factors: Sequence[float], operand_dtype=None) -> List[torch.Tensor]:
    return [neck_branch(x, p, i, float(s), operand_dtype) for i, s in enumerate(scale_factors)]


def pixel_decoder(feats: Sequence[torch.Tensor], p: Dict[str, torch.Tensor], shared_conv: bool = False, groups: int = 8,
                  eps: float = 1e-5, prefix: str = "", operand_dtype=None) -> torch.Tensor:
    """maskformer_segmentation.py:203-219: coarse-to-fine, prev = relu(GN(conv3x3(curr + nearest_up(prev))))."""
    dt = operand_dtype
    prev = _q(feats[-1], dt)
    fpn = list(feats[:-1])[::-1]
    for li, cur in enumerate(fpn):
        k = 0 if shared_conv else li
        prev = _q(_q(cur, dt) + F.interpolate(prev, size=cur.shape[-2:], mode="nearest"), dt)
        prev = F.conv2d(prev, _q(p[f"{prefix}conv_layers.{k}.weight"], dt), p[f"{prefix}conv_layers.{k}.bias"], padding=1)
        prev = F.relu(F.group_norm(prev, groups, p[f"{prefix}norms.{k}.weight"], p[f"{prefix}norms.{k}.bias"], eps))
        if li < len(fpn) - 1:
            prev = _q(prev, dt)
    return prev


def mlp(x: torch.Tensor, p: Dict[str, torch.Tensor], prefix: str, num_layers: int, operand_dtype=None) -> torch.Tensor:
    """model_misc.py:188-195 without dropout / residual / out_norm (the mask-embedding MLP uses none)."""
    for j in range(num_layers):
        x = F.linear(_q(x, operand_dtype), _q(p[f"{prefix}layers.{j}.weight"], operand_dtype), p[f"{prefix}layers.{j}.bias"])
        if j < num_layers - 1:
            x = F.relu(x)
    return x


def mask_predictor(obj_queries: torch.Tensor, pixel_embed: torch.Tensor, p: Dict[str, torch.Tensor],
                   prefix: str = "mask_predictor.", operand_dtype=None) -> torch.Tensor:
    """maskformer_segmentation.py:28-51."""
    me = _q(mlp(obj_queries, p, prefix + "mask_embed.", 3, operand_dtype), operand_dtype)
    pixel_embed = _q(pixel_embed, operand_dtype)
    if obj_queries.dim() == 3:
        return torch.einsum("bqc,bchw->bqhw", me, pixel_embed) if pixel_embed.dim() == 4 else torch.einsum("bqc,chw->bqhw", me, pixel_embed)
    return torch.einsum("lbqc,bchw->lbqhw", me, pixel_embed) if pixel_embed.dim() == 4 else torch.einsum("lbqc,chw->lbqhw", me, pixel_embed)


def seg_head(feats: Sequence[torch.Tensor], obj_queries: torch.Tensor, p: Dict[str, torch.Tensor], operand_dtype=None):
    """Pixel path of UniversalSegmentationHead.forward (maskformer_segmentation.py:314-336) for already-gathered per-query
    feature maps: pixel decoder -> instance 1x1 -> mask einsum, plus the semantic 1x1 head."""
    dt = operand_dtype
    pix = pixel_decoder(feats, p, prefix="pixel_decoder.", operand_dtype=dt)
    inst = F.conv2d(_q(pix, dt), _q(p["instance_seg_head.weight"], dt), p["instance_seg_head.bias"])
    masks = mask_predictor(obj_queries, inst, p, operand_dtype=dt)
    sem = F.conv2d(_q(pix, dt), _q(p["semantic_seg_head.weight"], dt), p["semantic_seg_head.bias"])
    return masks, sem


def make_neck_params(dim: int, d_model: int, scale_factors: Sequence[float], seed: int = 0, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)

    def rnd(*shape, std):
        return (torch.randn(*shape, generator=g) * std).to(dtype)

    p: Dict[str, torch.Tensor] = {}
    for i, s in enumerate(scale_factors):
        pre = f"convs.{i}."
        width = dim
        if s == 4.0:
            p[pre + "dconv_2x2_0.weight"] = rnd(dim, dim // 2, 2, 2, std=(dim) ** -0.5)
            p[pre + "dconv_2x2_0.bias"] = rnd(dim // 2, std=0.1)
            p[pre + "dconv_2x2_1.weight"] = rnd(dim // 2, dim // 4, 2, 2, std=(dim // 2) ** -0.5)
            p[pre + "dconv_2x2_1.bias"] = rnd(dim // 4, std=0.1)
            width = dim // 4
        elif s == 2.0:
            p[pre + "dconv_2x2.weight"] = rnd(dim, dim // 2, 2, 2, std=(dim) ** -0.5)
            p[pre + "dconv_2x2.bias"] = rnd(dim // 2, std=0.1)
            width = dim // 2
        p[pre + "conv_1x1.weight"] = rnd(d_model, width, 1, 1, std=width ** -0.5)
        p[pre + "conv_1x1.bias"] = rnd(d_model, std=0.1)
        p[pre + "conv_3x3.weight"] = rnd(d_model, d_model, 3, 3, std=(9 * d_model) ** -0.5)
        p[pre + "conv_3x3.bias"] = rnd(d_model, std=0.1)
    return p


def make_seg_params(d_model: int, stages: int, seed: int = 0, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)

    def rnd(*shape, std):
        return (torch.randn(*shape, generator=g) * std).to(dtype)

    p: Dict[str, torch.Tensor] = {}
    for k in range(stages):
        p[f"pixel_decoder.conv_layers.{k}.weight"] = rnd(d_model, d_model, 3, 3, std=(9 * d_model) ** -0.5 * 1.5)
        p[f"pixel_decoder.conv_layers.{k}.bias"] = rnd(d_model, std=0.1)
        p[f"pixel_decoder.norms.{k}.weight"] = 1.0 + rnd(d_model, std=0.2)
        p[f"pixel_decoder.norms.{k}.bias"] = rnd(d_model, std=0.2)
    p["instance_seg_head.weight"] = rnd(d_model, d_model, 1, 1, std=d_model ** -0.5)
    p["instance_seg_head.bias"] = rnd(d_model, std=0.1)
    p["semantic_seg_head.weight"] = rnd(1, d_model, 1, 1, std=d_model ** -0.5)
    p["semantic_seg_head.bias"] = rnd(1, std=0.1)
    for j in range(3):
        p[f"mask_predictor.mask_embed.layers.{j}.weight"] = rnd(d_model, d_model, std=d_model ** -0.5 * 1.4)
        p[f"mask_predictor.mask_embed.layers.{j}.bias"] = rnd(d_model, std=0.1)
    return p
