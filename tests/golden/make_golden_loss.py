"""Generates tests/golden/loss_small.npz with the reference's own loss functions (build container only):

    python tests/golden/make_golden_loss.py

Imported unmodified from /root/reference: sam3/train/loss/loss_fns.py `sigmoid_focal_loss(triton=False)`, `dice_loss`,
and sam3/model/data_misc.py `interpolate`, composed exactly as `Masks.get_loss` composes them (loss_fns.py:684-707).
torchmetrics (imported at module scope by loss_fns.py, unused on this path) is stubbed.
"""
from __future__ import annotations

import sys
import types
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
REF = Path("/root/reference")


def import_reference():
    tm = types.ModuleType("torchmetrics")
    tm.Metric = object
    tmf = types.ModuleType("torchmetrics.functional")
    tm.functional = tmf
    sys.modules.setdefault("torchmetrics", tm)
    sys.modules.setdefault("torchmetrics.functional", tmf)
    for name, path in (("sam3", REF / "sam3"), ("sam3.model", REF / "sam3" / "model"), ("sam3.train", REF / "sam3" / "train"),
                       ("sam3.train.loss", REF / "sam3" / "train" / "loss")):
        if name not in sys.modules:
            pkg = types.ModuleType(name)
            pkg.__path__ = [str(path)]
            sys.modules[name] = pkg
    import importlib

    return importlib.import_module("sam3.train.loss.loss_fns")


def main():
    lf = import_reference()
    g = torch.Generator().manual_seed(0)
    out = {}
    for tag, (N, h, w, H, W) in {"a": (3, 12, 12, 42, 42), "b": (2, 9, 16, 20, 50), "c": (2, 16, 16, 16, 16)}.items():
        src = (torch.randn(N, h, w, generator=g) * 3).requires_grad_(True)
        yy, xx = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
        tgt = torch.stack([((yy - H * 0.4 - 3 * n) ** 2 + (xx - W * 0.5) ** 2 < (0.3 * min(H, W)) ** 2) for n in range(N)]).float()
        num_boxes = 2.5
        s = src[:, None]
        s = lf.interpolate(s, size=tgt.shape[-2:], mode="bilinear", align_corners=False)[:, 0].flatten(1)
        t = tgt.flatten(1)
        lm = lf.sigmoid_focal_loss(s, t, num_boxes, alpha=0.25, gamma=2.0, triton=False)
        ld = lf.dice_loss(s, t, num_boxes)
        (1.3 * lm + 0.7 * ld).backward()
        out[f"{tag}.src"], out[f"{tag}.tgt"] = src.detach().numpy(), tgt.numpy()
        out[f"{tag}.loss_mask"], out[f"{tag}.loss_dice"] = lm.detach().numpy(), ld.detach().numpy()
        out[f"{tag}.dsrc"] = src.grad.numpy().copy()
    path = ROOT / "tests" / "golden" / "loss_small.npz"
    np.savez_compressed(path, **{k: np.asarray(v, dtype=np.float32) for k, v in out.items()})
    print("wrote", path, len(out), "arrays")


if __name__ == "__main__":
    main()
