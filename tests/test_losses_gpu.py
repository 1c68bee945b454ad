"""f1: fused sigmoid focal loss vs the oracle restatement of loss_fns.py:159-167 (values and gradients)."""
import pytest
import torch

from oracle import loss_oracle as LO
from tests.helpers import rel_max

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape,kw", [((6, 200), dict()), ((3, 5, 37, 41), dict(loss_on_multimask=True)),
                                      ((4, 1001), dict(reduce=False)), ((2, 300), dict(alpha=-1.0, gamma=1.5))])
def test_focal_loss_matches_oracle(shape, kw):
    from sam3_lora_b200.losses import sigmoid_focal_loss

    g = torch.Generator().manual_seed(0)
    x = torch.randn(shape, generator=g) * 3
    y = (torch.rand(shape, generator=g) > 0.7).float()
    xr = x.clone().requires_grad_(True)
    ref = LO.sigmoid_focal_loss(xr, y, 7.0, **kw)
    w = torch.randn(ref.shape, generator=g) if ref.dim() else torch.tensor(1.0)
    (ref * w).sum().backward()
    xc = x.clone().cuda().requires_grad_(True)
    out = sigmoid_focal_loss(xc, y.cuda(), 7.0, **kw)
    (out * w.cuda()).sum().backward()
    assert rel_max(out.detach().cpu(), ref.detach()) < 2e-5
    assert rel_max(xc.grad.cpu(), xr.grad) < 2e-4


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_fused_mask_losses_match_reference_golden(tag):
    """f1: bilinear up-sample + focal + dice in one kernel vs the reference's own functions (tests/golden/loss_small.npz)."""
    import numpy as np

    from sam3_lora_b200.losses import mask_losses
    from tests.helpers import GOLDEN

    z = np.load(GOLDEN / "loss_small.npz")
    for tdt in (torch.float32, torch.bool, torch.uint8):
        src = torch.from_numpy(z[f"{tag}.src"]).cuda().requires_grad_(True)
        tgt = torch.from_numpy(z[f"{tag}.tgt"]).cuda().to(tdt)
        out = mask_losses(src, tgt, 2.5)
        assert abs(out["loss_mask"].item() - float(z[f"{tag}.loss_mask"])) < 2e-5 * max(1.0, abs(float(z[f"{tag}.loss_mask"])))
        assert abs(out["loss_dice"].item() - float(z[f"{tag}.loss_dice"])) < 2e-5
        (1.3 * out["loss_mask"] + 0.7 * out["loss_dice"]).backward()
        assert rel_max(src.grad.cpu(), torch.from_numpy(z[f"{tag}.dsrc"])) < 2e-4


def test_fused_mask_losses_sam3_sizes_against_oracle_and_edge_cases():
    from sam3_lora_b200.losses import dice_loss, mask_losses

    g = torch.Generator().manual_seed(3)
    N, h, H = 5, 288, 1008                                  # SAM3: 288x288 logits against 1008x1008 targets (x3.5)
    src = torch.randn(N, h, h, generator=g) * 4
    tgt = torch.rand(N, 1, 36, 36, generator=g).gt(0.6).float()
    tgt = torch.nn.functional.interpolate(tgt, size=(H, H), mode="nearest")[:, 0].bool()
    sr = src.clone().requires_grad_(True)
    ref = LO.mask_losses(sr, tgt, 3.0)
    (ref["loss_mask"] + ref["loss_dice"]).backward()
    sc = src.cuda().requires_grad_(True)
    out = mask_losses(sc, tgt.cuda(), 3.0)
    (out["loss_mask"] + out["loss_dice"]).backward()
    assert abs(out["loss_mask"].item() - ref["loss_mask"].item()) < 2e-5 * abs(ref["loss_mask"].item())
    assert abs(out["loss_dice"].item() - ref["loss_dice"].item()) < 2e-5
    assert rel_max(sc.grad.cpu(), sr.grad) < 5e-4
    # deterministic: bitwise identical on a second run
    out2 = mask_losses(src.cuda(), tgt.cuda(), 3.0)
    assert out2["loss_mask"].item() == out["loss_mask"].item() and out2["loss_dice"].item() == out["loss_dice"].item()
    # no matched masks: both losses are zero (the reference's empty-selection branch, loss_fns.py:684-687)
    e = mask_losses(torch.zeros(0, h, h, device="cuda"), torch.zeros(0, H, H, device="cuda", dtype=torch.bool), 1.0)
    assert e["loss_mask"].item() == 0.0 and e["loss_dice"].item() == 0.0
    # dice_loss on flat, already-aligned inputs
    x = torch.randn(4, 999, generator=g)
    t = torch.rand(4, 999, generator=g).gt(0.5).float()
    xr = x.clone().requires_grad_(True)
    LO.dice_loss(xr, t, 2.0).backward()
    xc = x.cuda().requires_grad_(True)
    d = dice_loss(xc, t.cuda(), 2.0)
    d.backward()
    assert abs(d.item() - LO.dice_loss(x, t, 2.0).item()) < 1e-5
    assert rel_max(xc.grad.cpu(), xr.grad) < 2e-4
