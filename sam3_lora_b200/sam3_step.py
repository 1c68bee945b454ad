"""Row a9, second half: one SAM3 training step (forward -> matcher -> losses -> backward) as the reference's trainer
composes it (train_sam3_lora_native.py:743-793, 892-943), with the native pieces plugged in:

    model            sam3_bridge.build_native_model      (trunk / neck / pixel decoder / every MultiheadAttention native)
    matcher          sam3_lora_b200.matcher.BinaryHungarianMatcherV2    (GPU resident, no SciPy, no host copy of the cost matrix)
    mask losses      sam3_lora_b200.losses.mask_losses   (fused up-sample + focal + dice; replaces Masks.get_loss's tail)
    focal loss       sam3_lora_b200.losses.sigmoid_focal_loss           (replaces the Triton kernels, loss_fns.py:126-176)

The wrapper (`Sam3LossWrapper`, sam3/train/loss/sam3_loss.py), `Boxes`, `IABCEMdetr`, the one-to-many matcher and the
collator are the reference's own classes, imported through sam3_bridge — orchestration, not hot path.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.nn as nn

from . import sam3_bridge as bridge

RESOLUTION = 1008


# ------------------------------------------------------------------------------------------------
# synthetic COCO-shaped samples as the reference's Datapoint objects
# ------------------------------------------------------------------------------------------------
def synthetic_datapoints(n: int, seed: int = 0, resolution: int = RESOLUTION, max_objects: int = 3, prompt: str = "crack"):
    """n Datapoints shaped like `COCOSegmentDataset.__getitem__`'s (train_sam3_lora_native.py:92-232): a normalised
    [3, R, R] image, 1..max_objects boxes (normalised xyxy) with boolean [R, R] segments, one text query."""
    bridge.import_reference()
    from sam3.train.data.sam3_image_dataset import Datapoint, FindQueryLoaded, Image, InferenceMetadata, Object  # noqa: PLC0415

    g = torch.Generator().manual_seed(seed)
    R = resolution
    out = []
    for i in range(n):
        img = (torch.rand(3, R, R, generator=g) - 0.5) / 0.5
        k = int(torch.randint(1, max_objects + 1, (1,), generator=g).item())
        objs = []
        for j in range(k):
            x0, y0 = (torch.rand(2, generator=g) * 0.5).tolist()
            w, h = (0.1 + 0.4 * torch.rand(2, generator=g)).tolist()
            box = torch.tensor([x0, y0, x0 + w, y0 + h], dtype=torch.float32)
            seg = torch.zeros(R, R, dtype=torch.bool)
            yy, xx = torch.meshgrid(torch.arange(R), torch.arange(R), indexing="ij")
            cy, cx, ry, rx = (y0 + h / 2) * R, (x0 + w / 2) * R, h * R / 2, w * R / 2
            seg[((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 <= 1.0] = True      # an ellipse inscribed in the box
            objs.append(Object(bbox=box, area=(box[2] - box[0]) * (box[3] - box[1]), object_id=j, segment=seg))
        q = FindQueryLoaded(query_text=prompt, image_id=0, object_ids_output=list(range(k)), is_exhaustive=True,
                            query_processing_order=0,
                            inference_metadata=InferenceMetadata(coco_image_id=i, original_image_id=i, original_category_id=0,
                                                                 original_size=(R, R), object_id=-1, frame_index=-1))
        out.append(Datapoint(find_queries=[q], images=[Image(data=img, objects=objs, size=(R, R))], raw_images=[None]))
    return out


def collate(datapoints):
    """-> BatchedDatapoint (sam3/train/data/collator.py:136-360, the call of train_sam3_lora_native.py:826-827)."""
    bridge.import_reference()
    from sam3.train.data.collator import collate_fn_api  # noqa: PLC0415

    return collate_fn_api(datapoints, dict_key="input", with_seg_masks=True)["input"]


def move_to_device(obj, device, non_blocking: bool = False):
    """The trainer's recursive mover (train_sam3_lora_native.py:871-885)."""
    if isinstance(obj, torch.Tensor):
        return obj.to(device, non_blocking=non_blocking)
    if isinstance(obj, list):
        return [move_to_device(x, device, non_blocking) for x in obj]
    if isinstance(obj, tuple):
        return tuple(move_to_device(x, device, non_blocking) for x in obj)
    if isinstance(obj, dict):
        return {k: move_to_device(v, device, non_blocking) for k, v in obj.items()}
    if hasattr(obj, "__dataclass_fields__"):
        for f in obj.__dataclass_fields__:
            setattr(obj, f, move_to_device(getattr(obj, f), device, non_blocking))
    return obj


def disable_stochastic(model: nn.Module) -> nn.Module:
    """Parity runs: train() mode (auxiliary outputs, o2m queries, matcher inside the model) with every random op off —
    nn.Dropout, attention dropout of both MHA flavours, DropPath of both trunks."""
    from .mha import MultiheadAttention  # noqa: PLC0415
    from .vit import ViT  # noqa: PLC0415

    for m in model.modules():
        if isinstance(m, nn.Dropout):
            m.p = 0.0
        elif isinstance(m, (nn.MultiheadAttention, MultiheadAttention)):
            m.dropout = 0.0
        elif isinstance(m, ViT):
            m.drop_path_rates = [0.0] * len(m.drop_path_rates)
        elif type(m).__name__ == "DropPath":
            m.drop_prob = 0.0
    return model


# ------------------------------------------------------------------------------------------------
# objective
# ------------------------------------------------------------------------------------------------
def _native_masks_class():
    bridge.import_reference()
    from sam3.train.loss import loss_fns  # noqa: PLC0415

    from .losses import mask_losses  # noqa: PLC0415

    class NativeMasks(loss_fns.Masks):
        """`Masks` (loss_fns.py:568-707) with the non-sampled tail (bilinear up-sample to the target size + focal + dice)
        done by the fused kernel; matching / filtering of the pairs is the reference's code path."""

        def get_loss(self, outputs, targets, indices, num_boxes):
            src_masks = outputs["pred_masks"]
            if targets["masks"] is None or self.num_sample_points is not None or not src_masks.is_cuda:
                return super().get_loss(outputs, targets, indices, num_boxes)
            target_masks = targets["masks"] if indices[2] is None else targets["masks"][indices[2]]
            keep = targets["is_valid_mask"] if indices[2] is None else targets["is_valid_mask"][indices[2]]
            src_masks = src_masks[(indices[0], indices[1])][keep]
            target_masks = target_masks[keep]
            if target_masks.shape[0] == 0:
                z = src_masks.sum() * 0.0
                return {"loss_mask": z, "loss_dice": z}
            nb = float(num_boxes) if not isinstance(num_boxes, torch.Tensor) else num_boxes
            if isinstance(nb, torch.Tensor):
                # num_boxes is a 0-d device tensor (local normalisation, sam3_loss.py:74-80): keep it on the device
                l = mask_losses(src_masks, target_masks, 1.0, alpha=self.focal_alpha, gamma=self.focal_gamma)
                return {k: v / nb for k, v in l.items()}
            return mask_losses(src_masks, target_masks, nb, alpha=self.focal_alpha, gamma=self.focal_gamma)

    return NativeMasks


def build_objective(native: bool = True, normalization: str = "local"):
    """(matcher, loss_wrapper) with the weights of train_sam3_lora_native.py:743-793.

    `normalization` is Sam3LossWrapper's (sam3_loss.py:74-80): "local" (what the reference CLI passes, `:791`; bit-parity
    with the single-GPU run) divides by this rank's box count; "global" all-reduces the count over the process group
    first - the DDP-correct loss scale when the batch is sharded; "none" leaves the sums unscaled."""
    if normalization not in ("local", "global", "none"):
        raise ValueError(f"normalization={normalization!r}: expected 'local', 'global' or 'none'")
    bridge.import_reference()
    from sam3.train.loss import loss_fns  # noqa: PLC0415
    from sam3.train.loss.sam3_loss import Sam3LossWrapper  # noqa: PLC0415
    from sam3.train import matcher as ref_matcher  # noqa: PLC0415

    ref_focal = getattr(loss_fns, "_sam3b_ref_focal", None) or loss_fns.sigmoid_focal_loss
    loss_fns._sam3b_ref_focal = ref_focal
    if native:
        from . import losses  # noqa: PLC0415
        from .matcher import BinaryHungarianMatcherV2  # noqa: PLC0415

        Masks = _native_masks_class()
        matcher = BinaryHungarianMatcherV2(cost_class=2.0, cost_bbox=5.0, cost_giou=2.0, focal=True)
    else:
        losses = None
        Masks = loss_fns.Masks
        matcher = ref_matcher.BinaryHungarianMatcherV2(cost_class=2.0, cost_bbox=5.0, cost_giou=2.0, focal=True)

    def focal(inputs, targets, num_boxes, *a, **kw):
        if inputs.is_cuda and losses is not None:          # the fused kernel (no Triton)
            return losses.sigmoid_focal_loss(inputs, targets, num_boxes, *a, **kw)
        if not inputs.is_cuda:                             # the reference's eager branch (its Triton kernels need a GPU)
            kw["triton"] = False
        return ref_focal(inputs, targets, num_boxes, *a, **kw)

    loss_fns.sigmoid_focal_loss = focal
    fns = [
        loss_fns.Boxes(weight_dict={"loss_bbox": 5.0, "loss_giou": 2.0}),
        loss_fns.IABCEMdetr(pos_weight=10.0, weight_dict={"loss_ce": 20.0, "presence_loss": 20.0}, pos_focal=False, alpha=0.25,
                            gamma=2, use_presence=True, pad_n_queries=200),
        Masks(weight_dict={"loss_mask": 200.0, "loss_dice": 10.0}, focal_alpha=0.25, focal_gamma=2.0, compute_aux=False),
    ]
    o2m = ref_matcher.BinaryOneToManyMatcher(alpha=0.3, threshold=0.4, topk=4)
    wrapper = Sam3LossWrapper(loss_fns_find=fns, matcher=matcher, o2m_matcher=o2m, o2m_weight=2.0,
                              use_o2m_matcher_on_o2m_aux=False, normalization=normalization, normalize_by_valid_object_num=False)
    return matcher, wrapper


def training_loss(model: nn.Module, batch, matcher, loss_wrapper) -> Tuple[torch.Tensor, dict]:
    """Forward + the trainer's target conversion / matcher pass / loss wrapper (train_sam3_lora_native.py:898-931)."""
    bridge.import_reference()
    from sam3.model.model_misc import SAM3Output  # noqa: PLC0415
    from sam3.train.loss.loss_fns import CORE_LOSS_KEY  # noqa: PLC0415

    with torch.cuda.nvtx.range("sam3b.step.forward"):
        outputs_list = model(batch)
    find_targets = [model.back_convert(t) for t in batch.find_targets]
    torch.cuda.nvtx.range_push("sam3b.step.matcher+loss")
    with SAM3Output.iteration_mode(outputs_list, iter_mode=SAM3Output.IterMode.ALL_STEPS_PER_STAGE) as outputs_iter:
        for stage_outputs, stage_targets in zip(outputs_iter, find_targets):
            for outputs in stage_outputs:
                outputs["indices"] = matcher(outputs, stage_targets)
                for aux in outputs.get("aux_outputs", ()):
                    aux["indices"] = matcher(aux, stage_targets)
    loss_dict = loss_wrapper(outputs_list, find_targets)
    torch.cuda.nvtx.range_pop()
    return loss_dict[CORE_LOSS_KEY], loss_dict


def final_outputs(outputs_list) -> dict:
    """The last step's output dict of the (single) stage: pred_logits / pred_boxes / pred_masks / ..."""
    bridge.import_reference()
    from sam3.model.model_misc import SAM3Output  # noqa: PLC0415

    with SAM3Output.iteration_mode(outputs_list, iter_mode=SAM3Output.IterMode.ALL_STEPS_PER_STAGE) as it:
        stages = [list(s) for s in it]
    return stages[-1][-1]
