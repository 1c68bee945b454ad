"""Generates tests/golden/seg_small.npz by running the REAL reference neck / pixel decoder / mask predictor.

Run in the build container only (needs /root/reference; it does not exist on the GPU box):

    python tests/golden/make_golden_seg.py

Executed from the reference, unmodified: sam3/model/necks.py `Sam3DualViTDetNeck` (convs built by its own __init__),
sam3/model/maskformer_segmentation.py `PixelDecoder`, `MaskPredictor` (+ sam3/model/model_misc.py `MLP`), and the two 1x1
heads exactly as `UniversalSegmentationHead.forward` applies them (:322-335).  Reduced sizes (same code path): trunk width 64,
d_model 32, 8x8 trunk grid, batch 2, 8 queries.  Stored: parameters under the reference's state-dict names, inputs,
outputs, and input gradients for a fixed random output cotangent.
"""
from __future__ import annotations

import sys
import types
from pathlib import Path

import numpy as np
import torch
import torch.nn as nn

ROOT = Path(__file__).resolve().parents[2]
REF = Path("/root/reference")
sys.path.insert(0, str(ROOT))


def import_reference():
    for name, path in (("sam3", REF / "sam3"), ("sam3.model", REF / "sam3" / "model")):
        if name not in sys.modules:
            pkg = types.ModuleType(name)
            pkg.__path__ = [str(path)]
            sys.modules[name] = pkg
    import importlib

    necks = importlib.import_module("sam3.model.necks")
    seg = importlib.import_module("sam3.model.maskformer_segmentation")
    return necks, seg


class _Trunk(nn.Module):
    channel_list = [64]

    def forward(self, x):
        return [x]


class _NoPos(nn.Module):
    def forward(self, x):
        return torch.zeros_like(x)


def main():
    necks, seg = import_reference()
    torch.manual_seed(0)
    out = {}
    # ---------------- neck ----------------
    scales = (4.0, 2.0, 1.0, 0.5)
    neck = necks.Sam3DualViTDetNeck(_Trunk(), _NoPos(), d_model=32, scale_factors=scales)
    g = torch.Generator().manual_seed(3)
    for n, prm in neck.named_parameters():
        with torch.no_grad():
            prm.copy_(torch.randn(prm.shape, generator=g) * (0.1 if prm.dim() == 1 else (prm[0].numel()) ** -0.5))
        out["neck.param." + n] = prm.detach().numpy().copy()
    x = torch.randn(2, 64, 8, 8, generator=g, requires_grad=True)
    feats, _, _, _ = neck(x)
    cots = [torch.randn(f.shape, generator=g) for f in feats]
    (sum((f * c).sum() for f, c in zip(feats, cots))).backward()
    out["neck.x"] = x.detach().numpy()
    out["neck.dx"] = x.grad.numpy().copy()
    for i, (f, c) in enumerate(zip(feats, cots)):
        out[f"neck.out{i}"] = f.detach().numpy()
        out[f"neck.cot{i}"] = c.numpy()
    # per-branch input gradients as well (one backward each)
    for i in range(len(scales)):
        x2 = x.detach().clone().requires_grad_(True)
        f = neck.convs[i](x2)
        (f * cots[i]).sum().backward()
        out[f"neck.dx{i}"] = x2.grad.numpy().copy()

    # ---------------- pixel decoder + heads + mask predictor ----------------
    d = 32
    pd = seg.PixelDecoder(d, 2)
    mp = seg.MaskPredictor(d, mask_dim=d)
    inst_head = nn.Conv2d(d, d, kernel_size=1)
    sem_head = nn.Conv2d(d, 1, kernel_size=1)
    mods = {"pixel_decoder.": pd, "mask_predictor.": mp, "instance_seg_head.": inst_head, "semantic_seg_head.": sem_head}
    for pre, m in mods.items():
        for n, prm in m.named_parameters():
            with torch.no_grad():
                if "norms" in n:
                    prm.copy_((1.0 if n.endswith("weight") else 0.0) + torch.randn(prm.shape, generator=g) * 0.2)
                else:
                    prm.copy_(torch.randn(prm.shape, generator=g) * (0.1 if prm.dim() == 1 else 1.4 * (prm[0].numel()) ** -0.5))
            out["seg.param." + pre + n] = prm.detach().numpy().copy()
    fs = [torch.randn(2, d, 32, 32, generator=g, requires_grad=True), torch.randn(2, d, 16, 16, generator=g, requires_grad=True),
          torch.randn(2, d, 8, 8, generator=g, requires_grad=True)]
    q = torch.randn(2, 8, d, generator=g, requires_grad=True)
    pix = pd(fs)
    inst = inst_head(pix)
    masks = mp(q, inst)
    sem = sem_head(pix)
    cm, cs = torch.randn(masks.shape, generator=g), torch.randn(sem.shape, generator=g)
    ((masks * cm).sum() + (sem * cs).sum()).backward()
    for i, f in enumerate(fs):
        out[f"seg.feat{i}"] = f.detach().numpy()
        out[f"seg.dfeat{i}"] = f.grad.numpy().copy()
    out["seg.queries"], out["seg.dqueries"] = q.detach().numpy(), q.grad.numpy().copy()
    out["seg.pixel_embed"] = pix.detach().numpy()
    out["seg.masks"], out["seg.semantic"] = masks.detach().numpy(), sem.detach().numpy()
    out["seg.cot_masks"], out["seg.cot_semantic"] = cm.numpy(), cs.numpy()
    # aux-mask form (decoder-layer axis) of the mask predictor, forward only
    ql = torch.randn(3, 2, 8, d, generator=g)
    out["seg.queries_layers"] = ql.numpy()
    out["seg.masks_layers"] = mp(ql, inst).detach().numpy()

    path = ROOT / "tests" / "golden" / "seg_small.npz"
    np.savez_compressed(path, **{k: np.asarray(v, dtype=np.float32) for k, v in out.items()})
    print("wrote", path, f"{path.stat().st_size / 1e6:.2f} MB", len(out), "arrays")


if __name__ == "__main__":
    main()
