"""Loss kernels without Triton (next-row f1).  `sigmoid_focal_loss` / `dice_loss` keep the reference's signatures
(sam3/train/loss/loss_fns.py:79-123, 126-176); the `triton` flag is accepted and ignored: there is one CUDA path.
`mask_losses` is the fused form of the non-sampled branch of `Masks.get_loss` (loss_fns.py:684-707): bilinear up-sampling
of the matched mask logits to the target size + focal + dice in one pass over the targets, forward and backward."""
from __future__ import annotations

import torch

from . import _lib as L


class _FocalFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, inputs, targets, alpha: float, gamma: float, reduce_sum: bool):
        if not inputs.is_cuda:
            raise L.Sam3bError("sigmoid_focal_loss: inputs are on the CPU; the fused loss has no CPU fallback")
        x = inputs.detach().float().contiguous()
        y = targets.detach().float().contiguous()
        if x.data_ptr() % 16:          # a contiguous slice of a larger tensor: the kernel's float4 loads need 16-byte alignment
            x = x.clone()
        if y.data_ptr() % 16:
            y = y.clone()
        lib = L.load()
        n = x.numel()
        ctx.save_for_backward(x, y)
        ctx.meta = (alpha, gamma, reduce_sum, inputs.dtype, inputs.shape)
        if reduce_sum:
            s = torch.empty(1, device=x.device, dtype=torch.float32)
            L.check(lib.sam3b_focal_loss_fwd(x.data_ptr(), y.data_ptr(), n, alpha, gamma, None, s.data_ptr(), L.current_stream()))
            return s[0]
        out = torch.empty_like(x)
        L.check(lib.sam3b_focal_loss_fwd(x.data_ptr(), y.data_ptr(), n, alpha, gamma, out.data_ptr(), None, L.current_stream()))
        return out.view(inputs.shape)

    @staticmethod
    def backward(ctx, g):
        x, y = ctx.saved_tensors
        alpha, gamma, reduce_sum, dt, shape = ctx.meta
        dx = torch.empty_like(x)
        lib = L.load()
        if reduce_sum:   # the upstream scalar stays on the device (no .item() read-back in the training loop)
            L.check(lib.sam3b_focal_loss_bwd(x.data_ptr(), y.data_ptr(), x.numel(), alpha, gamma, None, 1.0, dx.data_ptr(), L.current_stream()))
            dx = dx * g
        else:
            gg = g.float().contiguous()
            L.check(lib.sam3b_focal_loss_bwd(x.data_ptr(), y.data_ptr(), x.numel(), alpha, gamma, gg.data_ptr(), 1.0, dx.data_ptr(), L.current_stream()))
        return dx.view(shape).to(dt), None, None, None, None


def sigmoid_focal_loss(inputs, targets, num_boxes, alpha: float = 0.25, gamma: float = 2, loss_on_multimask: bool = False,
                       reduce: bool = True, triton: bool = True):
    """Same returns as the reference: reduce=False -> elementwise loss; loss_on_multimask -> [N, M] spatial means / num_boxes;
    otherwise loss.mean(1).sum() / num_boxes."""
    if reduce and not loss_on_multimask:
        total = _FocalFn.apply(inputs, targets, float(alpha), float(gamma), True)
        return total / (num_boxes * inputs.shape[1])
    loss = _FocalFn.apply(inputs, targets, float(alpha), float(gamma), False)
    if not reduce:
        return loss
    assert loss.dim() == 4
    return loss.flatten(2).mean(-1) / num_boxes


class _MaskLossFn(torch.autograd.Function):
    """(loss_mask, loss_dice) as one [2] tensor; see csrc/loss.cu (mask_loss_*_kernel)."""

    @staticmethod
    def forward(ctx, src, tgt, num_boxes: float, alpha: float, gamma: float):
        if not src.is_cuda:
            raise L.Sam3bError("mask_losses: inputs are on the CPU; the fused loss has no CPU fallback")
        x = src.detach().float().contiguous()
        N, h, w = x.shape
        if tgt.dtype == torch.bool:
            t, u8 = tgt.contiguous().view(torch.uint8), 1
        elif tgt.dtype == torch.uint8:
            t, u8 = tgt.contiguous(), 1
        else:
            t, u8 = tgt.detach().float().contiguous(), 0
        H, W = t.shape[-2:]
        strips = (H + 7) // 8
        partial = torch.empty(max(1, N * strips * 4), device=x.device, dtype=torch.float32)
        sums = torch.empty(max(1, N), 4, device=x.device, dtype=torch.float32)
        out = torch.empty(2, device=x.device, dtype=torch.float32)
        L.check(L.load().sam3b_mask_loss_fwd(x.data_ptr(), N, h, w, t.data_ptr(), u8, H, W, alpha, gamma, num_boxes,
                                             partial.data_ptr(), sums.data_ptr(), out.data_ptr(), L.current_stream()))
        ctx.save_for_backward(x, t, sums)
        ctx.meta = (u8, H, W, alpha, gamma, num_boxes, src.dtype)
        return out

    @staticmethod
    def backward(ctx, g):
        x, t, sums = ctx.saved_tensors
        u8, H, W, alpha, gamma, num_boxes, dt = ctx.meta
        N, h, w = x.shape
        dx = torch.empty_like(x)
        gg = g.detach().float().contiguous()
        L.check(L.load().sam3b_mask_loss_bwd(x.data_ptr(), N, h, w, t.data_ptr(), u8, H, W, alpha, gamma, num_boxes,
                                             sums.data_ptr(), gg.data_ptr(), dx.data_ptr(), L.current_stream()))
        return dx.to(dt), None, None, None, None


def mask_losses(src_masks, target_masks, num_boxes, alpha: float = 0.25, gamma: float = 2.0):
    """src_masks [N, h, w] (or [N, 1, h, w]) matched mask logits; target_masks [N, H, W] bool / uint8 / float in {0, 1}.
    Returns {"loss_mask", "loss_dice"} exactly as `Masks.get_loss` does after F.interpolate(..., mode="bilinear",
    align_corners=False) (loss_fns.py:689-707), without materialising the up-sampled logits."""
    if src_masks.dim() == 4:
        src_masks = src_masks[:, 0]
    if target_masks.dim() == 4:
        target_masks = target_masks[:, 0]
    out = _MaskLossFn.apply(src_masks, target_masks, float(num_boxes), float(alpha), float(gamma))
    return {"loss_mask": out[0], "loss_dice": out[1]}


def dice_loss(inputs, targets, num_boxes, loss_on_multimask: bool = False, reduce: bool = True):
    """loss_fns.py:79-123 for the form SAM3's image losses use: inputs / targets [N, P] (already at the same size),
    returns sum_n (1 - (2 sum sig*t + 1) / (sum sig + sum t + 1)) / num_boxes."""
    if loss_on_multimask or not reduce:
        raise NotImplementedError("dice_loss: only the reduced single-mask form is on the fused path")
    N = inputs.shape[0]
    out = _MaskLossFn.apply(inputs.reshape(N, 1, -1), targets.reshape(N, 1, -1), float(num_boxes), 0.25, 2.0)
    return out[1]
