"""The neck / pixel-decoder / mask-head oracle is pinned against tests/golden/seg_small.npz, which
tests/golden/make_golden_seg.py produced by running the reference's own modules (necks.py, maskformer_segmentation.py)."""
import numpy as np
import torch

from oracle import seg_oracle as SO
from tests.helpers import GOLDEN, rel_max

SCALES = (4.0, 2.0, 1.0, 0.5)


def _load():
    z = np.load(GOLDEN / "seg_small.npz")
    return {k: torch.from_numpy(z[k]) for k in z.files}


def test_neck_oracle_matches_reference_forward_and_input_gradients():
    z = _load()
    p = {k[len("neck.param."):]: v for k, v in z.items() if k.startswith("neck.param.")}
    x = z["neck.x"].clone().requires_grad_(True)
    outs = SO.neck(x, p, SCALES)
    for i, o in enumerate(outs):
        assert o.shape == z[f"neck.out{i}"].shape
        assert rel_max(o.detach(), z[f"neck.out{i}"]) < 5e-6
    sum((o * z[f"neck.cot{i}"]).sum() for i, o in enumerate(outs)).backward()
    assert rel_max(x.grad, z["neck.dx"]) < 2e-5
    for i, s in enumerate(SCALES):
        x2 = z["neck.x"].clone().requires_grad_(True)
        (SO.neck_branch(x2, p, i, s) * z[f"neck.cot{i}"]).sum().backward()
        assert rel_max(x2.grad, z[f"neck.dx{i}"]) < 2e-5


def test_seg_head_oracle_matches_reference_forward_and_gradients():
    z = _load()
    p = {k[len("seg.param."):]: v for k, v in z.items() if k.startswith("seg.param.")}
    feats = [z[f"seg.feat{i}"].clone().requires_grad_(True) for i in range(3)]
    q = z["seg.queries"].clone().requires_grad_(True)
    pix = SO.pixel_decoder(feats, p, prefix="pixel_decoder.")
    assert rel_max(pix.detach(), z["seg.pixel_embed"]) < 5e-6
    masks, sem = SO.seg_head(feats, q, p)
    assert rel_max(masks.detach(), z["seg.masks"]) < 5e-6
    assert rel_max(sem.detach(), z["seg.semantic"]) < 5e-6
    ((masks * z["seg.cot_masks"]).sum() + (sem * z["seg.cot_semantic"]).sum()).backward()
    for i in range(3):
        assert rel_max(feats[i].grad, z[f"seg.dfeat{i}"]) < 5e-5
    assert rel_max(q.grad, z["seg.dqueries"]) < 5e-5
    # decoder-layer ("aux masks") form of the einsum
    inst = torch.nn.functional.conv2d(pix.detach(), p["instance_seg_head.weight"], p["instance_seg_head.bias"])
    assert rel_max(SO.mask_predictor(z["seg.queries_layers"], inst, p), z["seg.masks_layers"]) < 5e-6
