"""Host logic of row a8 on the CPU: the frozen-weight packings of conv_ops (what the GEMMs multiply by) are checked against
PyTorch's own convolutions by emulating the device kernels' documented semantics (im2col3x3, pixel shuffle) with torch ops."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from sam3_lora_b200 import conv_ops as CO, ops


def _im2col(x_nhwc):
    """sam3b_im2col3x3 semantics: [B,H,W,C] -> [B*H*W, 9C] with k = (ky, kx, c), zero padding."""
    B, H, W, C = x_nhwc.shape
    xp = F.pad(x_nhwc, (0, 0, 1, 1, 1, 1))
    cols = [xp[:, ky:ky + H, kx:kx + W, :] for ky in range(3) for kx in range(3)]
    return torch.cat(cols, dim=-1).reshape(B * H * W, 9 * C)


def _shuffle(u, B, H, W, C):
    """sam3b_pixel_shuffle2 semantics: [B*H*W, 4C] columns (di, dj, c) -> [B, 2H, 2W, C]."""
    return u.view(B, H, W, 2, 2, C).permute(0, 1, 3, 2, 4, 5).reshape(B, 2 * H, 2 * W, C)


def setup_module(_):
    ops.set_operand_dtype(torch.float16)


def test_conv3x3_packing_forward_and_data_gradient():
    torch.manual_seed(0)
    conv = nn.Conv2d(16, 24, 3, padding=1).requires_grad_(False)
    w9, wg, bias = CO.pack_conv3x3(conv)
    x = torch.randn(2, 16, 6, 5, requires_grad=True)
    conv16 = nn.Conv2d(16, 24, 3, padding=1)
    conv16.weight.data = conv.weight.half().float(); conv16.bias.data = conv.bias.clone()
    y = conv16(x)
    y_pack = _im2col(x.detach().permute(0, 2, 3, 1)) @ w9.float().t() + bias
    assert torch.allclose(y_pack, y.permute(0, 2, 3, 1).reshape(-1, 24), atol=1e-5)
    g = torch.randn_like(y)
    y.backward(g)
    dx_pack = _im2col(g.permute(0, 2, 3, 1)) @ wg.float().t()
    assert torch.allclose(dx_pack, x.grad.permute(0, 2, 3, 1).reshape(-1, 16), atol=1e-5)


def test_deconv2x2_packing_forward_and_data_gradient():
    torch.manual_seed(1)
    dc = nn.ConvTranspose2d(16, 8, kernel_size=2, stride=2).requires_grad_(False)
    wd, wdt, bias = CO.pack_deconv2x2(dc)
    ref = nn.ConvTranspose2d(16, 8, kernel_size=2, stride=2)
    ref.weight.data = dc.weight.half().float(); ref.bias.data = dc.bias.clone()
    x = torch.randn(2, 16, 4, 3, requires_grad=True)
    y = ref(x)
    u = x.detach().permute(0, 2, 3, 1).reshape(-1, 16) @ wd.float().t() + bias
    assert torch.allclose(_shuffle(u, 2, 4, 3, 8), y.permute(0, 2, 3, 1), atol=1e-5)
    g = torch.randn_like(y)
    y.backward(g)
    gu = g.permute(0, 2, 3, 1).view(2, 4, 2, 3, 2, 8).permute(0, 1, 3, 2, 4, 5).reshape(-1, 32)     # pixel_unshuffle2
    assert torch.allclose(gu @ wdt.float().t(), x.grad.permute(0, 2, 3, 1).reshape(-1, 16), atol=1e-5)


def test_conv1x1_packing_pads_output_channels_to_eight():
    torch.manual_seed(2)
    c = nn.Conv2d(16, 1, 1).requires_grad_(False)
    w, wt, bias = CO.pack_conv1x1(c)
    assert w.shape == (8, 16) and wt.shape == (16, 8) and bias.shape == (8,)
    assert torch.equal(w[1:], torch.zeros(7, 16, dtype=w.dtype)) and torch.equal(bias[1:], torch.zeros(7))
    x = torch.randn(5, 16)
    assert torch.allclose((x @ w.float().t() + bias)[:, :1], F.linear(x, c.weight.half().float().view(1, 16), c.bias), atol=1e-6)
    # cache: same tensors -> same packed objects; an in-place update re-packs; trainable convs are refused
    assert CO.pack_conv1x1(c)[0] is w
    c.weight.data.add_(1.0)
    c.weight.add_(0.0)      # bumps the version counter
    assert CO.pack_conv1x1(c)[0] is not w
    import pytest
    from sam3_lora_b200._lib import Sam3bError
    with pytest.raises(Sam3bError):
        CO.pack_conv1x1(nn.Conv2d(4, 8, 1))
