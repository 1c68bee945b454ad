"""CPU oracle of the mask losses.  TEST INFRASTRUCTURE ONLY.  Restates the `triton=False` branch of
sam3/train/loss/loss_fns.py:126-176 (the CPU-able statement of the reference's Triton kernels), `_dice_loss` (:105-123) and
the non-sampled branch of `Masks.get_loss` (:684-707).  Pinned by tests/golden/loss_small.npz, which
tests/golden/make_golden_loss.py produces by calling the reference's own functions (tests/test_loss_oracle.py)."""
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


def dice_loss(inputs, targets, num_boxes):
    """loss_fns.py:105-123, single-mask reduced form."""
    p = inputs.sigmoid().flatten(1)
    t = targets.flatten(1)
    numerator = 2 * (p * t).sum(1)
    denominator = p.sum(-1) + t.sum(-1)
    return (1 - (numerator + 1) / (denominator + 1)).sum() / num_boxes


def mask_losses(src_masks, target_masks, num_boxes, alpha: float = 0.25, gamma: float = 2.0):
    """loss_fns.py:684-707: up-sample the matched logits to the target size (bilinear, align_corners=False), then focal + dice."""
    if src_masks.dim() == 3:
        src_masks = src_masks[:, None]
    up = F.interpolate(src_masks.float(), size=target_masks.shape[-2:], mode="bilinear", align_corners=False)[:, 0].flatten(1)
    t = target_masks.to(up).flatten(1)
    return {"loss_mask": sigmoid_focal_loss(up, t, num_boxes, alpha=alpha, gamma=gamma), "loss_dice": dice_loss(up, t, num_boxes)}
