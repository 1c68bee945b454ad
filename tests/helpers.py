"""Shared helpers for the parity tests."""
from __future__ import annotations

from pathlib import Path

import numpy as np
import torch

GOLDEN = Path(__file__).resolve().parent / "golden"


def load_small_golden():
    """tests/golden/vit_small.npz (made by make_golden.py from the real reference)."""
    from oracle.vit_oracle import LoRASpec, ViTConfig

    z = np.load(GOLDEN / "vit_small.npz")
    cfg = ViTConfig(img_size=224, patch_size=14, embed_dim=128, depth=2, num_heads=2, mlp_hidden=608, window_size=8,
                    global_att_blocks=(1,), pretrain_img_size=112)
    spec = LoRASpec(rank=4, alpha=8.0)
    params = {k[6:]: torch.from_numpy(z[k].astype(np.float32)) for k in z.files if k.startswith("param:")}
    grads = {k[5:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("grad:")}
    return {
        "cfg": cfg, "spec": spec, "params": params, "grads": grads,
        "img": torch.from_numpy(z["img"].astype(np.float32)),
        "gout": torch.from_numpy(z["gout"].astype(np.float32)),
        "out": torch.from_numpy(z["out"]),
        "ln_pre_out": torch.from_numpy(z["ln_pre_out"]),
        "blocks": [torch.from_numpy(z[f"block{i}_out"]) for i in range(cfg.depth)],
        "ref_saved_keys": [str(s) for s in z["ref_saved_keys"]],
        "ref_saved_shapes": [str(s) for s in z["ref_saved_shapes"]],
    }


def rel_l2(got: torch.Tensor, ref: torch.Tensor) -> float:
    got, ref = got.double(), ref.double()
    return ((got - ref).norm() / ref.norm().clamp_min(1e-30)).item()


def rel_max(got: torch.Tensor, ref: torch.Tensor) -> float:
    got, ref = got.double(), ref.double()
    return ((got - ref).abs().max() / ref.abs().max().clamp_min(1e-30)).item()
