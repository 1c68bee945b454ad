"""The mask-loss oracle is pinned against tests/golden/loss_small.npz, produced by the reference's own loss functions
(tests/golden/make_golden_loss.py: sigmoid_focal_loss(triton=False), dice_loss, interpolate as Masks.get_loss composes them)."""
import numpy as np
import pytest
import torch

from oracle import loss_oracle as LO
from tests.helpers import GOLDEN, rel_max


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_mask_loss_oracle_matches_reference(tag):
    z = np.load(GOLDEN / "loss_small.npz")
    src = torch.from_numpy(z[f"{tag}.src"]).requires_grad_(True)
    tgt = torch.from_numpy(z[f"{tag}.tgt"])
    out = LO.mask_losses(src, tgt, 2.5)
    assert abs(out["loss_mask"].item() - float(z[f"{tag}.loss_mask"])) < 1e-6 * max(1.0, abs(float(z[f"{tag}.loss_mask"])))
    assert abs(out["loss_dice"].item() - float(z[f"{tag}.loss_dice"])) < 1e-6
    (1.3 * out["loss_mask"] + 0.7 * out["loss_dice"]).backward()
    assert rel_max(src.grad, torch.from_numpy(z[f"{tag}.dsrc"])) < 1e-5
