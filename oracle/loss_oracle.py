"""CPU oracle of the sigmoid focal loss.  TEST INFRASTRUCTURE ONLY.  Restates the `triton=False` branch of
sam3/train/loss/loss_fns.py:126-176 (the CPU-able statement of the reference's Triton kernels)."""
import torch
import torch.nn.functional as F


def sigmoid_focal_loss(inputs, targets, num_boxes, alpha: float = 0.25, gamma: float = 2, loss_on_multimask=False, reduce=True):
    prob = inputs.sigmoid()
    ce = F.binary_cross_entropy_with_logits(inputs, targets, reduction="none")
    p_t = prob * targets + (1 - prob) * (1 - targets)
    loss = ce * ((1 - p_t) ** gamma)
    if alpha >= 0:
        loss = (alpha * targets + (1 - alpha) * (1 - targets)) * loss
    if not reduce:
        return loss
    if loss_on_multimask:
        return loss.flatten(2).mean(-1) / num_boxes
    return loss.mean(1).sum() / num_boxes
