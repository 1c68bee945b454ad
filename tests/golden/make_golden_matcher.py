"""Generates tests/golden/matcher_small.npz by running the reference's own BinaryHungarianMatcherV2 on the CPU
(build container only):  python tests/golden/make_golden_matcher.py
Imported unmodified from /root/reference: sam3/train/matcher.py, sam3/model/box_ops.py."""
from __future__ import annotations

import sys
import types
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
REF = Path("/root/reference")

sys.path.insert(0, str(ROOT))
from tests.matcher_cases import CASES  # noqa: E402


def import_reference():
    for name, path in (("sam3", REF / "sam3"), ("sam3.model", REF / "sam3" / "model"), ("sam3.train", REF / "sam3" / "train")):
        if name not in sys.modules:
            pkg = types.ModuleType(name)
            pkg.__path__ = [str(path)]
            sys.modules[name] = pkg
    import importlib

    return importlib.import_module("sam3.train.matcher")


def make_inputs(B, Q, num_boxes, repeat_batch, masks, seed):
    g = torch.Generator().manual_seed(seed)
    Tmax = max(max(num_boxes), 1)

    def boxes(*shape):
        c = torch.rand(*shape, 2, generator=g) * 0.8 + 0.1
        wh = torch.rand(*shape, 2, generator=g) * 0.3 + 0.02
        return torch.cat([c, wh], -1)

    outs = {"pred_logits": torch.randn(B * repeat_batch, Q, 1, generator=g) * 2, "pred_boxes": boxes(B * repeat_batch, Q)}
    tg = {"boxes_padded": boxes(B, Tmax), "num_boxes": torch.tensor(num_boxes, dtype=torch.long)}
    ov = tv = None
    if masks:
        ov = torch.rand(B, Q, generator=g) > 0.2
        tv = torch.rand(B, Tmax, generator=g) > 0.3
    return outs, tg, ov, tv


def main():
    m = import_reference()
    out = {}
    for seed, (name, (kw, B, Q, nb, rep, rb, masks)) in enumerate(CASES.items()):
        outs, tg, ov, tv = make_inputs(B, Q, nb, rb, masks, seed)
        matcher = m.BinaryHungarianMatcherV2(**kw)
        bi, si, ti = matcher(outs, tg, repeats=rep, repeat_batch=rb, out_is_valid=ov, target_is_valid_padded=tv)
        out[f"{name}.logits"] = outs["pred_logits"].numpy()
        out[f"{name}.pred_boxes"] = outs["pred_boxes"].numpy()
        out[f"{name}.boxes_padded"] = tg["boxes_padded"].numpy()
        if masks:
            out[f"{name}.out_valid"], out[f"{name}.tgt_valid"] = ov.numpy(), tv.numpy()
        out[f"{name}.batch_idx"], out[f"{name}.src_idx"] = bi.numpy(), si.numpy()
        out[f"{name}.has_tgt"] = np.array(ti is not None)
        out[f"{name}.tgt_idx"] = ti.numpy() if ti is not None else np.zeros(0, np.int64)
    path = ROOT / "tests" / "golden" / "matcher_small.npz"
    np.savez_compressed(path, **out)
    print("wrote", path, f"{path.stat().st_size / 1e3:.0f} kB")


if __name__ == "__main__":
    main()
