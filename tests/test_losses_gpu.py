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
