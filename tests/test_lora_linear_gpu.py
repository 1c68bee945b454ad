"""Row a1: the stand-alone fused LoRALinear op vs the reference's LoRALinear golden (lora_layers.py:58-91)
and vs a PyTorch fp32 computation, forward and backward, including adapter dropout semantics."""
import numpy as np
import pytest
import torch

from tests.helpers import GOLDEN, rel_l2

pytestmark = pytest.mark.gpu


def test_lora_linear_matches_reference_golden():
    from sam3_lora_b200.ops import lora_linear

    z = np.load(GOLDEN / "lora_linear.npz")
    t = {k: torch.from_numpy(z[k]).cuda() for k in z.files if k != "scaling"}
    s = float(z["scaling"])
    x = t["x"].clone().requires_grad_(True)
    A = t["A"].clone().requires_grad_(True)
    B = t["B"].clone().requires_grad_(True)
    y = lora_linear(x, t["W"], t["b"], A, B, s)
    assert rel_l2(y.cpu(), t["y"].cpu()) < 1e-3
    (y * t["gy"]).sum().backward()
    assert rel_l2(x.grad.cpu(), t["dx"].cpu()) < 2e-3
    assert rel_l2(A.grad.cpu(), t["dA"].cpu()) < 2e-3
    assert rel_l2(B.grad.cpu(), t["dB"].cpu()) < 2e-3


def test_lora_linear_tiny_and_huge_cotangents():
    """Backward runs on a device-chosen power-of-two multiple of the cotangent (fp16 operands): 1e-10 and 1e7 scale alike."""
    from sam3_lora_b200.ops import lora_linear

    z = np.load(GOLDEN / "lora_linear.npz")
    t = {k: torch.from_numpy(z[k]).cuda() for k in z.files if k != "scaling"}
    s = float(z["scaling"])
    for k in (1.0e-10, 1.0e7):
        x = t["x"].clone().requires_grad_(True)
        A = t["A"].clone().requires_grad_(True)
        B = t["B"].clone().requires_grad_(True)
        (lora_linear(x, t["W"], t["b"], A, B, s) * (t["gy"] * k)).sum().backward()
        assert rel_l2(x.grad.cpu() / k, t["dx"].cpu()) < 2e-3
        assert rel_l2(A.grad.cpu() / k, t["dA"].cpu()) < 2e-3
        assert rel_l2(B.grad.cpu() / k, t["dB"].cpu()) < 2e-3


def test_lora_linear_module_detr_shape_and_small_batch():
    """d=256 Linear as in the DETR / seg-head projections, 3-D input, tiny row count (M < one MMA tile)."""
    from sam3_lora_b200.lora_layers import LoRALinear

    torch.manual_seed(0)
    lin = torch.nn.Linear(256, 768).cuda()
    mod = LoRALinear(lin, rank=8, alpha=16).cuda()
    torch.nn.init.normal_(mod.lora.lora_B, std=0.05)
    x = torch.randn(5, 3, 256, device="cuda", requires_grad=True)
    y = mod(x)
    ref = x @ lin.weight.T + lin.bias + (x @ mod.lora.lora_A @ mod.lora.lora_B) * mod.lora.scaling
    assert y.shape == (5, 3, 768) and rel_l2(y.detach().cpu(), ref.detach().cpu()) < 1e-3
    g = torch.randn_like(y)
    gx, gA, gB = torch.autograd.grad(ref, [x, mod.lora.lora_A, mod.lora.lora_B], g, retain_graph=True)
    y.backward(g)
    assert rel_l2(x.grad.cpu(), gx.cpu()) < 2e-3
    assert rel_l2(mod.lora.lora_A.grad.cpu(), gA.cpu()) < 2e-3
    assert rel_l2(mod.lora.lora_B.grad.cpu(), gB.cpu()) < 2e-3
    assert lin.weight.grad is None     # frozen base


def test_lora_linear_dropout_only_on_adapter_branch():
    from sam3_lora_b200.lora_layers import LoRALinear

    torch.manual_seed(1)
    lin = torch.nn.Linear(128, 128).cuda()
    mod = LoRALinear(lin, rank=4, alpha=8, dropout=0.5).cuda().train()
    torch.nn.init.normal_(mod.lora.lora_B, std=0.1)
    x = torch.randn(64, 128, device="cuda")
    y1, y2 = mod(x), mod(x)
    base = x @ lin.weight.T + lin.bias
    full = base + (x @ mod.lora.lora_A @ mod.lora.lora_B) * mod.lora.scaling
    assert not torch.allclose(y1, y2)                                   # stochastic adapter branch
    assert (y1 - base).abs().max() > 1e-3 and (y1 - full).abs().max() > 1e-3
    mod.eval()
    assert rel_l2(mod(x).cpu(), full.detach().cpu()) < 1e-3            # identity in eval
