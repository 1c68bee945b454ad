"""`train_sam3_lora_native.py` surface (CLI + YAML schema) on top of the native trunk.

Mirrors the reference's trainer shell (train_sam3_lora_native.py:696-1050): one `--config X.yaml`
argument; reads `lora.*`, `training.{learning_rate,weight_decay,data_dir,batch_size,num_epochs}` and
`output.output_dir` (the only keys the reference's native trainer reads); COCO-format
`data_dir/{train,valid}/_annotations.coco.json` resized to 1008x1008 and normalised with mean/std 0.5
(:46-232); AdamW over the adapters (:736-740); saves `best_lora_weights.pt` / `last_lora_weights.pt`
with `save_lora_weights` and appends JSON lines to `val_stats.json` (:995-1016).

Scope note: the objective here is a stand-in — a 1x1-conv mask head on the trunk features trained with
BCE + dice against the union of the image's instance masks at 72x72 (the HF-path trainer of the reference also trains with
a plain mask BCE, train_sam3_lora.py:319-355) — because the full `Sam3Image` wiring (text encoder, geometry encoder, DETR
encoder/decoder around the trunk) is the reference's and is not rebuilt in this repo.  The pieces of that wiring that ARE
on the hot path exist as drop-in modules and are exercised end to end by tests/test_chain_gpu.py: `necks.py`,
`maskformer_segmentation.py`, `mha.py`, `matcher.py` (GPU Hungarian matcher), `losses.py` (fused focal / dice / up-sample).
Everything trunk-side — LoRA injection, fused kernels, flat-gradient all-reduce, checkpoint format — is the production path.
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

    def __init__(self, data_dir, split: str = "train", mask_size: int = 72, resolution: int = RESOLUTION, device=None):
        # device: a CUDA device -> the image is resized / normalised on the GPU by data.GpuPreprocessor (bit-identical to the
        # PIL + numpy path below; only the raw uint8 pixels cross PCIe).  None -> host path (CPU tests, the reference's way).
        self.gpu_prep = None
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

    def __getitem__(self, idx):
        from PIL import Image

        info = self.images[self.image_ids[idx]]
        img = Image.open(self.split_dir / info["file_name"]).convert("RGB")
        w0, h0 = img.size
        if self.gpu_prep is not None:
            x = self.gpu_prep.image(np.asarray(img))
        else:
            img = img.resize((self.resolution, self.resolution), Image.BILINEAR)
            x = torch.from_numpy(np.asarray(img, dtype=np.float32) / 255.0).permute(2, 0, 1)
            x = (x - 0.5) / 0.5
        union = np.zeros((h0, w0), dtype=np.uint8)
        names = []
        for a in self.img_to_anns.get(info["id"], []):
            union |= _ann_to_mask(a, h0, w0)
            names.append(self.categories.get(a.get("category_id"), "object"))
        m = torch.from_numpy(union.astype(np.float32))[None, None]
        m = F.interpolate(m, size=(self.mask_size, self.mask_size), mode="area")[0]
        return {"image": x, "mask": m, "prompt": names[0] if names else "object"}


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
        self.batch_size = int(tc["batch_size"])
        dropout = float(lc.get("dropout", 0.0))
        self.model = TrunkWithProxyHead(max_batch=self.batch_size, **(vit_overrides or {}))
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
        for p in self.model.proxy_mask_head.parameters():
            p.requires_grad = True
        self.model.to(self.device)
        stats = count_parameters(self.model)
        print(f"Trainable params: {stats['trainable_parameters']:,} ({stats['trainable_percentage']:.2f}%)")
        params = [p for p in self.model.parameters() if p.requires_grad]
        self.optimizer = torch.optim.AdamW(params, lr=float(tc["learning_rate"]), weight_decay=float(tc["weight_decay"]), fused=True)
        if self.world > 1:
            self.model.trunk.grad_hook = D.LoRAGradAllReducer()
        self.out_dir = Path(self.config["output"]["output_dir"])
        self.out_dir.mkdir(parents=True, exist_ok=True)

    def _loader(self, split: str, epoch: int, shuffle: bool):
        spec = self.model.trunk.spec
        gpu_prep = bool(self.config["training"].get("gpu_preprocess", True))      # not a reference key; default on
        ds = COCOSegmentDataset(self.config["training"]["data_dir"], split, mask_size=spec.grid, resolution=spec.img_size,
                                device=self.device if gpu_prep else None)
        idx = D.shard_indices(len(ds), self.rank, self.world, epoch=epoch, shuffle=shuffle)
        sub = torch.utils.data.Subset(ds, idx)
        return torch.utils.data.DataLoader(sub, batch_size=self.batch_size, shuffle=False, num_workers=0, collate_fn=collate,
                                           pin_memory=not gpu_prep, drop_last=False)

    def _sync_head_grads(self):
        if self.world > 1:
            for p in self.model.proxy_mask_head.parameters():
                torch.distributed.all_reduce(p.grad)
                p.grad.mul_(1.0 / self.world)

    def train(self):
        epochs = int(self.config["training"]["num_epochs"])
        best = math.inf
        for epoch in range(epochs):
            self.model.train()
            t0, seen, running = time.time(), 0, 0.0
            for batch in self._loader("train", epoch, shuffle=True):
                img = batch["image"].to(self.device, non_blocking=True)
                tgt = batch["mask"].to(self.device, non_blocking=True)
                loss = mask_loss(self.model(img), tgt)
                self.optimizer.zero_grad(set_to_none=True)
                loss.backward()
                self._sync_head_grads()
                self.optimizer.step()
                running += loss.item() * img.shape[0]
                seen += img.shape[0]
            train_loss = running / max(seen, 1)
            val_loss = self.validate(epoch)
            if self.rank == 0:
                dt = time.time() - t0
                print(f"Epoch {epoch + 1}/{epochs}  train_loss {train_loss:.4f}  val_loss {val_loss:.4f}  "
                      f"{seen * self.world / dt:.2f} img/s")
                save_lora_weights(self.model, str(self.out_dir / "last_lora_weights.pt"))
                if val_loss < best:
                    best = val_loss
                    save_lora_weights(self.model, str(self.out_dir / "best_lora_weights.pt"))
                with open(self.out_dir / "val_stats.json", "a") as f:
                    f.write(json.dumps({"epoch": epoch + 1, "train_loss": train_loss, "val_loss": val_loss}) + "\n")
        return best

    @torch.no_grad()
    def validate(self, epoch: int) -> float:
        try:
            loader = self._loader("valid", epoch, shuffle=False)
        except FileNotFoundError:
            return float("nan")
        self.model.eval()
        tot, n = 0.0, 0
        for batch in loader:
            img = batch["image"].to(self.device, non_blocking=True)
            tgt = batch["mask"].to(self.device, non_blocking=True)
            tot += mask_loss(self.model(img), tgt).item() * img.shape[0]
            n += img.shape[0]
        t = torch.tensor([tot, n], device=self.device, dtype=torch.float64)
        if self.world > 1:
            torch.distributed.all_reduce(t)
        return (t[0] / t[1].clamp_min(1)).item()


def main(argv=None):
    ap = argparse.ArgumentParser(description="Train SAM3 with LoRA (B200-native trunk)")
    ap.add_argument("--config", type=str, default="configs/full_lora_config.yaml", help="Path to YAML configuration file")
    args = ap.parse_args(argv)
    trainer = SAM3TrainerNative(args.config)
    trainer.train()


if __name__ == "__main__":
    main()
