"""Loss kernels without Triton (next-row f1).  `sigmoid_focal_loss` keeps the reference's signature
(sam3/train/loss/loss_fns.py:126-176); the `triton` flag is accepted and ignored: there is one CUDA path."""
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
        if reduce_sum:
            gs = float(g.item()) if g.numel() == 1 and not g.requires_grad else None
            if gs is None:
                L.check(lib.sam3b_focal_loss_bwd(x.data_ptr(), y.data_ptr(), x.numel(), alpha, gamma, None, 1.0, dx.data_ptr(), L.current_stream()))
                dx = dx * g
            else:
                L.check(lib.sam3b_focal_loss_bwd(x.data_ptr(), y.data_ptr(), x.numel(), alpha, gamma, None, gs, dx.data_ptr(), L.current_stream()))
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
