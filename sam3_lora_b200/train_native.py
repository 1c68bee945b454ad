"""`train_sam3_lora_native.py` surface (CLI + YAML schema) on top of the native hot path.

Mirrors the reference's trainer shell (train_sam3_lora_native.py:696-1050): one `--config X.yaml` argument; reads `lora.*`,
`training.{learning_rate,weight_decay,data_dir,batch_size,num_epochs}` and `output.output_dir` (the only keys the reference's
native trainer reads); COCO-format `data_dir/{train,valid}/_annotations.coco.json` resized to 1008x1008
and normalised with mean/std 0.5 (:46-232); AdamW over the adapters (:736-740); saves `best_lora_weights.pt` /
`last_lora_weights.pt` with `save_lora_weights` and appends JSON lines to `val_stats.json` (:995-1016).

Objective (`training.objective`, not a reference key):
  * "sam3" (default) - the reference's own step: `Sam3Image.forward` on the swapped model (sam3_bridge.build_native_model:
    native trunk / neck / pixel decoder / every MultiheadAttention), Hungarian matching on the GPU and `Sam3LossWrapper` with the
    fused mask / focal losses (sam3_step.build_objective), exactly the wiring of train_sam3_lora_native.py:743-793, 892-943.
    Needs the reference's `sam3` package importable (SAM3_REFERENCE_ROOT, baseline/_ref, or sys.path) and the base weights:
    `model.checkpoint_path` in the YAML or $SAM3_CHECKPOINT (a `sam3.pt`).  Random base weights are refused unless
    `model.allow_random_init: true` (tests / benchmarks only).
  * "trunk_proxy" - the trunk alone with a 1x1 mask head and BCE + dice against the union mask; for smoke tests of
    non-SAM3 trunk shapes only.
Extra optional keys: `training.gpu_preprocess` (default true), `training.gpu_jpeg_decode` (nvJPEG, default false),
`training.prefetch` (batches prepared ahead by a background thread, default 2; 0 = the reference's synchronous loader),
`training.seed` (default 0), `training.max_steps`, `training.cuda_graphs`.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import time
from pathlib import Path
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
import yaml

from . import dist as D
from .lora_layers import LoRAConfig, apply_lora_to_model, count_parameters, save_lora_weights
from .vit import ViT

RESOLUTION = 1008


# ------------------------------------------------------------------------------------------------
# data: COCO-format segmentation set (train_sam3_lora_native.py:46-232), no pycocotools needed
# ------------------------------------------------------------------------------------------------
def _rle_to_mask(rle: Dict, h: int, w: int) -> np.ndarray:
    counts = rle["counts"]
    if isinstance(counts, str):
        raise ValueError("compressed RLE needs pycocotools; re-export the dataset with polygon or uncompressed RLE")
    flat = np.zeros(h * w, dtype=np.uint8)
    pos, val = 0, 0
    for c in counts:
        if val:
            flat[pos:pos + c] = 1
        pos += c
        val ^= 1
    return flat.reshape(w, h).T  # COCO RLE is column-major


def _ann_to_mask(ann: Dict, h: int, w: int) -> np.ndarray:
    from PIL import Image, ImageDraw

    seg = ann.get("segmentation")
    if isinstance(seg, dict):
        return _rle_to_mask(seg, h, w)
    m = Image.new("L", (w, h), 0)
    if isinstance(seg, list) and seg:
        draw = ImageDraw.Draw(m)
        for poly in seg:
            if len(poly) >= 6:
                draw.polygon([(poly[i], poly[i + 1]) for i in range(0, len(poly), 2)], outline=1, fill=1)
    elif "bbox" in ann:
        x, y, bw, bh = ann["bbox"]
        ImageDraw.Draw(m).rectangle([x, y, x + bw, y + bh], fill=1)
    return np.asarray(m, dtype=np.uint8)


class COCOSegmentDataset(torch.utils.data.Dataset):
    """Same directory contract as the reference: `<data_dir>/<split>/_annotations.coco.json` + images."""

    def __init__(self, data_dir, split: str = "train", mask_size: int = 72, resolution: int = RESOLUTION, device=None,
                 with_instances: bool = False, gpu_jpeg: bool = False):
        # device: a CUDA device -> the image is resized / normalised on the GPU by data.GpuPreprocessor (bit-identical to the
        # PIL + numpy path below; only the raw uint8 pixels cross PCIe).  None -> host path (CPU tests, the reference's way).
        self.gpu_prep = None
        self.gpu_jpeg = gpu_jpeg
        self.with_instances = with_instances      # per-object boxes + [R, R] boolean segments (the SAM3 objective's targets)
        if device is not None and torch.device(device).type == "cuda":
            from .data import GpuPreprocessor  # noqa: PLC0415

            self.gpu_prep = GpuPreprocessor(resolution, device=device)
        self.split_dir = Path(data_dir) / split
        ann_file = self.split_dir / "_annotations.coco.json"
        if not ann_file.exists():
            raise FileNotFoundError(f"COCO annotation file not found: {ann_file}")
        coco = json.loads(ann_file.read_text())
        self.images = {im["id"]: im for im in coco["images"]}
        self.image_ids = sorted(self.images)
        self.img_to_anns: Dict[int, List[Dict]] = {}
        for a in coco["annotations"]:
            self.img_to_anns.setdefault(a["image_id"], []).append(a)
        self.categories = {c["id"]: c["name"] for c in coco["categories"]}
        self.resolution = resolution
        self.mask_size = mask_size
        print(f"Loaded COCO dataset: {split} split")
        print(f"  Images: {len(self.image_ids)}")
        print(f"  Annotations: {len(coco['annotations'])}")
        print(f"  Categories: {self.categories}")

    def __len__(self):
        return len(self.image_ids)

    def _open_image(self, path: Path):
        """-> (pixels for the resize step, original width, original height).  JPEG files go through nvJPEG on the GPU when
        `gpu_jpeg` is on (a few grey levels from libjpeg-turbo's decode: off by default, PIL decode is the reference's)."""
        from PIL import Image

        if self.gpu_prep is not None and self.gpu_jpeg and path.suffix.lower() in (".jpg", ".jpeg"):
            u8 = self.gpu_prep.decode_jpeg(path.read_bytes())
            return u8, u8.shape[1], u8.shape[0]
        img = Image.open(path).convert("RGB")
        return img, img.size[0], img.size[1]

    def _segments_gpu(self, anns: List[Dict], h0: int, w0: int) -> torch.Tensor:
        """Instance masks at the model resolution straight from the annotations on the GPU (data.GpuPreprocessor): polygons
        through the rleFrPoly kernels, RLE through the run-length kernel, box-only annotations as their rectangle."""
        R = self.resolution
        out = torch.zeros(len(anns), R, R, dtype=torch.bool, device=self.gpu_prep.device)
        poly_idx, poly_objs, rle_idx, rle_objs = [], [], [], []
        for i, a in enumerate(anns):
            seg = a.get("segmentation")
            if isinstance(seg, dict):
                hh, ww = seg.get("size", (h0, w0))
                rle_idx.append(i)
                rle_objs.append((seg["counts"], int(hh), int(ww)))
            elif isinstance(seg, list) and seg:
                poly_idx.append(i)
                poly_objs.append((seg, h0, w0))
            else:
                x, y, bw, bh = a["bbox"]
                poly_idx.append(i)
                poly_objs.append(([[x, y, x + bw, y, x + bw, y + bh, x, y + bh]], h0, w0))
        if poly_objs:
            out[torch.tensor(poly_idx, device=out.device)] = self.gpu_prep.polygon_masks(poly_objs)
        if rle_objs:
            out[torch.tensor(rle_idx, device=out.device)] = self.gpu_prep.rle_masks(rle_objs)
        return out

    def __getitem__(self, idx):
        from PIL import Image

        info = self.images[self.image_ids[idx]]
        img, w0, h0 = self._open_image(self.split_dir / info["file_name"])
        if self.gpu_prep is not None:
            x = self.gpu_prep.image(img if isinstance(img, torch.Tensor) else np.asarray(img))
        else:
            img = img.resize((self.resolution, self.resolution), Image.BILINEAR)
            x = torch.from_numpy(np.asarray(img, dtype=np.float32) / 255.0).permute(2, 0, 1)
            x = (x - 0.5) / 0.5
        anns = self.img_to_anns.get(info["id"], [])
        names = [self.categories.get(a.get("category_id"), "object") for a in anns]
        item = {"image": x, "prompt": _query_text(names), "image_id": info["id"], "orig_size": (h0, w0)}
        if self.with_instances:
            # per-object targets of the SAM3 objective: COCO xywh -> normalised xyxy boxes (train_sam3_lora_native.py:129-143)
            # and [R, R] boolean segments (:146-167)
            R = self.resolution
            inst = [a for a in anns if "bbox" in a]
            boxes = [[a["bbox"][0] / w0, a["bbox"][1] / h0, (a["bbox"][0] + a["bbox"][2]) / w0, (a["bbox"][1] + a["bbox"][3]) / h0]
                     for a in inst]
            if not inst:
                seg = torch.zeros(0, R, R, dtype=torch.bool)
            elif self.gpu_prep is not None:
                seg = self._segments_gpu(inst, h0, w0)
            else:       # host path (CPU tests): PIL's scan-line fill, which differs from pycocotools on boundary pixels
                seg = torch.from_numpy(np.stack([_ann_to_mask(a, h0, w0) for a in inst]))
                seg = F.interpolate(seg[:, None].float(), size=(R, R), mode="nearest")[:, 0] > 0.5
            item["boxes"] = torch.tensor(boxes, dtype=torch.float32).reshape(-1, 4)
            item["segments"] = seg
        else:
            # the trunk-proxy objective's target: union of the instance masks, area-averaged to the trunk grid
            union = np.zeros((h0, w0), dtype=np.uint8)
            for a in anns:
                union |= _ann_to_mask(a, h0, w0)
            m = torch.from_numpy(union.astype(np.float32))[None, None]
            item["mask"] = F.interpolate(m, size=(self.mask_size, self.mask_size), mode="area")[0]
        return item


def _query_text(names: List[str]) -> str:
    """The image's text prompt: its (most common) category name, lower-cased; "object" without annotations (:196-210)."""
    if not names:
        return "object"
    from collections import Counter  # noqa: PLC0415

    return Counter(names).most_common(1)[0][0].lower()


def to_datapoint(item: Dict):
    """One dataset item -> the reference's `Datapoint` (Image + Objects + one FindQueryLoaded), train_sam3_lora_native.py:169-232."""
    from . import sam3_bridge  # noqa: PLC0415

    sam3_bridge.import_reference()
    from sam3.train.data.sam3_image_dataset import Datapoint, FindQueryLoaded, Image, InferenceMetadata, Object  # noqa: PLC0415

    R = item["image"].shape[-1]
    objs = []
    for i, (box, seg) in enumerate(zip(item["boxes"], item["segments"])):
        objs.append(Object(bbox=box, area=(box[2] - box[0]) * (box[3] - box[1]), object_id=i, segment=seg))
    q = FindQueryLoaded(query_text=item["prompt"], image_id=0, object_ids_output=[o.object_id for o in objs], is_exhaustive=True,
                        query_processing_order=0,
                        inference_metadata=InferenceMetadata(coco_image_id=item["image_id"], original_image_id=item["image_id"],
                                                             original_category_id=0, original_size=tuple(item["orig_size"]),
                                                             object_id=-1, frame_index=-1))
    return Datapoint(find_queries=[q], images=[Image(data=item["image"], objects=objs, size=(R, R))], raw_images=[None])


def collate_sam3(batch):
    """-> BatchedDatapoint through the reference's own collator (`collate_fn_api`, :826-827)."""
    from .sam3_step import collate as _collate  # noqa: PLC0415

    return _collate([to_datapoint(b) for b in batch])


def collate(batch):
    return {"image": torch.stack([b["image"] for b in batch]), "mask": torch.stack([b["mask"] for b in batch]),
            "prompt": [b["prompt"] for b in batch]}


# ------------------------------------------------------------------------------------------------
class TrunkWithProxyHead(nn.Module):
    """Trunk under the reference's module path + the stand-in mask head (see module docstring)."""

    def __init__(self, max_batch: int, operand_dtype=torch.float16, **vit_kw):
        super().__init__()
        self.backbone = nn.Module()
        self.backbone.vision_backbone = nn.Module()
        self.backbone.vision_backbone.trunk = ViT(max_batch=max_batch, operand_dtype=operand_dtype, **vit_kw)
        d = self.backbone.vision_backbone.trunk.spec.embed_dim
        self.proxy_mask_head = nn.Conv2d(d, 1, kernel_size=1)

    @property
    def trunk(self) -> ViT:
        return self.backbone.vision_backbone.trunk

    def forward(self, images):
        feat = self.trunk(images)[0]
        return self.proxy_mask_head(feat)


def mask_loss(logits, target):
    bce = F.binary_cross_entropy_with_logits(logits, target)
    p = torch.sigmoid(logits)
    inter = (p * target).sum(dim=(1, 2, 3))
    dice = 1 - (2 * inter + 1) / (p.sum(dim=(1, 2, 3)) + target.sum(dim=(1, 2, 3)) + 1)
    return bce + dice.mean()


def _broadcast_from_rank0(model: nn.Module) -> None:
    """Every rank must start from rank 0's parameters and buffers (adapters are initialised from each process's own RNG)."""
    if not (torch.distributed.is_available() and torch.distributed.is_initialized()) or torch.distributed.get_world_size() == 1:
        return
    trunk = getattr(getattr(getattr(model, "backbone", None), "vision_backbone", None), "trunk", None)
    for t in list(model.parameters()) + list(model.buffers()):
        if t.is_complex():
            continue
        torch.distributed.broadcast(t.data, src=0)
    if isinstance(trunk, ViT):
        trunk.refresh_base()


class _FlatGradSync:
    """Averages the gradients of the adapters OUTSIDE the trunk (and of any extra trainable parameter) over ranks with one
    collective; the trunk's own flat buffer is reduced by dist.LoRAGradAllReducer right after its last backward kernel."""

    def __init__(self, params: List[nn.Parameter]):
        self.params = params

    def __call__(self):
        if not self.params or not torch.distributed.is_initialized() or torch.distributed.get_world_size() == 1:
            return
        grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in self.params]
        flat = torch.cat([g.reshape(-1) for g in grads])
        torch.distributed.all_reduce(flat)
        flat.mul_(1.0 / torch.distributed.get_world_size())
        off = 0
        for p, g in zip(self.params, grads):
            n = g.numel()
            p.grad = flat[off:off + n].view_as(g).clone()
            off += n


class SAM3TrainerNative:
    def __init__(self, config_path: str, vit_overrides: Optional[Dict] = None):
        with open(config_path) as f:
            self.config = yaml.safe_load(f)
        self.rank, self.local_rank, self.world = D.init_from_env()
        if not torch.cuda.is_available():
            raise RuntimeError("SAM3TrainerNative needs a CUDA device: the native path has no CPU fallback")
        self.device = torch.device("cuda", self.local_rank)
        torch.cuda.set_device(self.device)
        lc = self.config["lora"]
        tc = self.config["training"]
        mc = self.config.get("model") or {}
        self.batch_size = int(tc["batch_size"])
        self.objective = str(tc.get("objective", "trunk_proxy" if vit_overrides else "sam3"))
        self.max_steps = tc.get("max_steps")
        seed = int(tc.get("seed", 0))
        torch.manual_seed(seed)            # identical construction-time RNG on every rank; rank 0 is broadcast anyway
        np.random.seed(seed)
        dropout = float(lc.get("dropout", 0.0))
        ckpt = mc.get("checkpoint_path") or os.environ.get("SAM3_CHECKPOINT")
        allow_random = bool(mc.get("allow_random_init", False))
        if self.objective == "sam3":
            from . import sam3_bridge, sam3_step  # noqa: PLC0415

            if ckpt is None and not allow_random:
                raise RuntimeError("no base checkpoint: set model.checkpoint_path in the YAML or $SAM3_CHECKPOINT to a sam3.pt "
                                   "(the reference downloads facebook/sam3 from the Hub, train_sam3_lora_native.py:705-711); "
                                   "random base weights need model.allow_random_init: true (tests / benchmarks only)")
            self.model = sam3_bridge.build_native_model("cpu", checkpoint_path=ckpt, seed=seed if ckpt is None else None,
                                                        max_batch=self.batch_size,
                                                        cuda_graphs=bool(tc.get("cuda_graphs", False)))
            # "local" = the reference CLI's choice (bit-parity with its single-GPU run); "global" all-reduces the box count
            # so the loss scale is right when the batch is sharded over ranks (training.loss_normalization)
            self.matcher, self.loss_wrapper = sam3_step.build_objective(native=True,
                                                                        normalization=str(tc.get("loss_normalization", "local")))
        elif self.objective == "trunk_proxy":
            self.model = TrunkWithProxyHead(max_batch=self.batch_size, **(vit_overrides or {}))
            if ckpt is not None:
                sd = torch.load(ckpt, map_location="cpu")
                sd = sd.get("model", sd)
                pre = "backbone.vision_backbone.trunk."
                missing, _ = self.model.trunk.load_state_dict({k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)}, strict=False)
                if [k for k in missing if "freqs_cis" not in k]:
                    raise RuntimeError(f"checkpoint {ckpt} lacks trunk weights: {missing[:5]}")
            elif not allow_random and not vit_overrides:
                raise RuntimeError("no base checkpoint for the trunk (model.checkpoint_path / $SAM3_CHECKPOINT); "
                                   "random init needs model.allow_random_init: true")
        else:
            raise ValueError(f"training.objective must be 'sam3' or 'trunk_proxy', got {self.objective!r}")
        lora_config = LoRAConfig(
            rank=lc["rank"], alpha=lc["alpha"], dropout=dropout, target_modules=lc["target_modules"],
            apply_to_vision_encoder=lc.get("apply_to_vision_encoder", True),
            apply_to_text_encoder=lc.get("apply_to_text_encoder", True),
            apply_to_geometry_encoder=lc.get("apply_to_geometry_encoder", False),
            apply_to_detr_encoder=lc.get("apply_to_detr_encoder", True),
            apply_to_detr_decoder=lc.get("apply_to_detr_decoder", True),
            apply_to_mask_decoder=lc.get("apply_to_mask_decoder", False),
            strict_reference_names=bool(lc.get("strict_reference_names", False)))
        self.model = apply_lora_to_model(self.model, lora_config)
        if self.objective == "trunk_proxy":
            for p in self.model.proxy_mask_head.parameters():
                p.requires_grad = True
        self.model.to(self.device)
        _broadcast_from_rank0(self.model)
        stats = count_parameters(self.model)
        print(f"Trainable params: {stats['trainable_parameters']:,} ({stats['trainable_percentage']:.2f}%)")
        params = [p for p in self.model.parameters() if p.requires_grad]
        self.optimizer = torch.optim.AdamW(params, lr=float(tc["learning_rate"]), weight_decay=float(tc["weight_decay"]), fused=True)
        trunk_ids = {id(p) for p in self.trunk.lora_parameters()}
        self._sync_rest = _FlatGradSync([p for p in params if id(p) not in trunk_ids])
        if self.world > 1:
            self.trunk.grad_hook = D.LoRAGradAllReducer(segments=4)      # slices are reduced while the backward continues
        self.out_dir = Path(self.config["output"]["output_dir"])
        self.out_dir.mkdir(parents=True, exist_ok=True)

    @property
    def trunk(self) -> ViT:
        return self.model.backbone.vision_backbone.trunk

    def _loader(self, split: str, epoch: int, shuffle: bool):
        spec = self.trunk.spec
        gpu_prep = bool(self.config["training"].get("gpu_preprocess", True))      # not a reference key; default on
        sam3 = self.objective == "sam3"
        ds = COCOSegmentDataset(self.config["training"]["data_dir"], split, mask_size=spec.grid, resolution=spec.img_size,
                                device=self.device if gpu_prep else None, with_instances=sam3,
                                gpu_jpeg=bool(self.config["training"].get("gpu_jpeg_decode", False)))
        idx = D.shard_indices(len(ds), self.rank, self.world, epoch=epoch, shuffle=shuffle)
        sub = torch.utils.data.Subset(ds, idx)
        loader = torch.utils.data.DataLoader(sub, batch_size=self.batch_size, shuffle=False, num_workers=0,
                                             collate_fn=collate_sam3 if sam3 else collate,
                                             pin_memory=(not gpu_prep) and not sam3, drop_last=False)
        depth = int(self.config["training"].get("prefetch", 2))
        if depth <= 0:
            return loader
        # the next batch's file reads, annotation parsing, GPU pre-processing and H2D copies run behind the current step
        from .data import Prefetcher  # noqa: PLC0415

        return Prefetcher(loader, depth=depth, device=self.device)

    def _loss(self, batch):
        """(loss, number of images) of one batch under the configured objective."""
        if self.objective == "sam3":
            from . import sam3_step  # noqa: PLC0415

            batch = sam3_step.move_to_device(batch, self.device, non_blocking=True)
            loss, _ = sam3_step.training_loss(self.model, batch, self.matcher, self.loss_wrapper)
            return loss, int(batch.img_batch.shape[0])
        img = batch["image"].to(self.device, non_blocking=True)
        tgt = batch["mask"].to(self.device, non_blocking=True)
        return mask_loss(self.model(img), tgt), int(img.shape[0])

    def train(self):
        epochs = int(self.config["training"]["num_epochs"])
        best = math.inf
        steps = 0
        for epoch in range(epochs):
            self.model.train()
            t0, seen, running = time.time(), 0, 0.0
            for batch in self._loader("train", epoch, shuffle=True):
                loss, n = self._loss(batch)
                self.optimizer.zero_grad(set_to_none=True)
                loss.backward()
                self._sync_rest()
                self.optimizer.step()
                running += loss.item() * n
                seen += n
                steps += 1
                if self.max_steps is not None and steps >= int(self.max_steps):
                    break
            train_loss = running / max(seen, 1)
            val_loss = self.validate(epoch)
            if self.rank == 0:
                dt = time.time() - t0
                print(f"Epoch {epoch + 1}/{epochs}  train_loss {train_loss:.4f}  val_loss {val_loss:.4f}  "
                      f"{seen * self.world / dt:.2f} img/s")
                save_lora_weights(self.model, str(self.out_dir / "last_lora_weights.pt"))
                # without a validation split the reference keeps only the last adapters (train_sam3_lora_native.py:1017-1025)
                if val_loss is not None and not math.isnan(val_loss) and val_loss < best:
                    best = val_loss
                    save_lora_weights(self.model, str(self.out_dir / "best_lora_weights.pt"))
                with open(self.out_dir / "val_stats.json", "a") as f:
                    f.write(json.dumps({"epoch": epoch + 1, "train_loss": train_loss, "val_loss": val_loss}) + "\n")
            if self.max_steps is not None and steps >= int(self.max_steps):
                break
        return best

    @torch.no_grad()
    def validate(self, epoch: int) -> float:
        try:
            loader = self._loader("valid", epoch, shuffle=False)
        except FileNotFoundError:
            return float("nan")
        self.model.eval()          # validation loss only, in eval mode under no_grad (train_sam3_lora_native.py:948-990)
        tot, n = 0.0, 0
        for batch in loader:
            loss, k = self._loss(batch)
            tot += loss.item() * k
            n += k
        t = torch.tensor([tot, n], device=self.device, dtype=torch.float64)
        if self.world > 1:
            torch.distributed.all_reduce(t)
        return (t[0] / t[1].clamp_min(1)).item() if t[1] > 0 else float("nan")


def main(argv=None):
    ap = argparse.ArgumentParser(description="Train SAM3 with LoRA (B200-native trunk)")
    ap.add_argument("--config", type=str, default="configs/full_lora_config.yaml", help="Path to YAML configuration file")
    args = ap.parse_args(argv)
    trainer = SAM3TrainerNative(args.config)
    trainer.train()


if __name__ == "__main__":
    main()
