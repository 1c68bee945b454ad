"""Full-size chained step on one GPU: SAM3 trunk (32 blocks, rank-16 LoRA) -> SimpleFPN neck -> pixel decoder -> instance head ->
mask einsum (200 queries) -> GPU Hungarian matcher -> fused up-sample + focal + dice loss -> backward to the adapters -> AdamW.
Random frozen weights, synthetic targets; prints one JSON line (CUDA events).  Not the headline bench (bench.py measures the
trunk step BASELINE.json names); this shows the built rows composing at SAM3's sizes.

    python tools/bench_chain.py [--batch 8] [--steps 4]
"""
from __future__ import annotations

import argparse
import contextlib
import json
import sys
from pathlib import Path

import torch
import torch.nn as nn

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))


class _NoPos(nn.Module):
    def forward(self, x):
        return x.new_zeros(1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--depth", type=int, default=32)
    ap.add_argument("--objects", type=int, default=4, help="ground-truth objects per image")
    args = ap.parse_args()
    from sam3_lora_b200 import _lib, conv_ops as CO
    from sam3_lora_b200.lora_layers import LoRAConfig, apply_lora_to_model, get_lora_parameters
    from sam3_lora_b200.losses import mask_losses
    from sam3_lora_b200.maskformer_segmentation import PixelDecoder, UniversalSegmentationHead
    from sam3_lora_b200.matcher import BinaryHungarianMatcherV2
    from sam3_lora_b200.necks import Sam3DualViTDetNeck
    from sam3_lora_b200.vit import ViT

    dev = "cuda:0"
    torch.manual_seed(0)
    B, Q, T, d = args.batch, 200, args.objects, 256
    globals_ = tuple(i for i in (7, 15, 23, 31) if i < args.depth) or (args.depth - 1,)
    trunk = ViT(depth=args.depth, global_att_blocks=globals_, max_batch=B, cuda_graphs=True)
    with contextlib.redirect_stdout(sys.stderr):
        apply_lora_to_model(trunk, LoRAConfig(rank=16, alpha=32, dropout=0.0,
                                              target_modules=["q_proj", "k_proj", "v_proj", "out_proj", "fc1", "fc2"]))
    for p in get_lora_parameters(trunk):
        if p.shape[0] == 16:
            nn.init.normal_(p, std=0.02)
    neck = Sam3DualViTDetNeck(trunk, _NoPos(), d_model=d, scale_factors=(4.0, 2.0, 1.0))
    head = UniversalSegmentationHead(d, 2, PixelDecoder(d, 2))
    for m in (neck.convs, head):
        for p in m.parameters():
            p.requires_grad_(False)
    neck, head = neck.to(dev).train(), head.to(dev)
    queries = nn.Parameter(torch.randn(B, Q, d, device=dev) * 0.5)          # stands in for the DETR decoder's output
    params = get_lora_parameters(trunk) + [queries]
    opt = torch.optim.AdamW(params, lr=5e-5, weight_decay=0.01, fused=True)
    matcher = BinaryHungarianMatcherV2(cost_class=2.0, cost_bbox=5.0, cost_giou=2.0, focal=True)

    g = torch.Generator(device=dev).manual_seed(1)
    img = torch.randn(B, 3, 1008, 1008, device=dev, generator=g)
    logits = torch.randn(B, Q, 1, device=dev, generator=g) * 2
    pboxes = torch.cat([torch.rand(B, Q, 2, device=dev, generator=g) * 0.8 + 0.1, torch.rand(B, Q, 2, device=dev, generator=g) * 0.3 + 0.05], -1)
    tboxes = torch.cat([torch.rand(B, T, 2, device=dev, generator=g) * 0.8 + 0.1, torch.rand(B, T, 2, device=dev, generator=g) * 0.3 + 0.05], -1)
    targets = {"boxes_padded": tboxes, "num_boxes": torch.full((B,), T, dtype=torch.long)}
    tmasks = torch.nn.functional.interpolate(torch.rand(B * T, 1, 36, 36, device=dev, generator=g).gt(0.6).float(), size=(1008, 1008))[:, 0].bool()
    num_boxes = float(B * T)

    def step():
        feats = neck(img)[0]
        pix = head.pixel_decoder(feats)
        masks = head.mask_predictor(queries, CO.conv1x1_forward(pix, head.instance_seg_head))      # [B, Q, 288, 288]
        bi, si, _ = matcher({"pred_logits": logits, "pred_boxes": pboxes}, targets)
        losses = mask_losses(masks[(bi, si)], tmasks, num_boxes)
        loss = 200.0 * losses["loss_mask"] + 10.0 * losses["loss_dice"]                             # weights of train_sam3_lora_native.py:766-770
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        return loss

    for _ in range(3):
        last = step()
    torch.cuda.synchronize()
    n0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        last = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    print(json.dumps({"what": "trunk + neck + pixel decoder + mask head + GPU matcher + fused mask loss, fwd+bwd+AdamW", "batch": B,
                      "queries": Q, "objects_per_image": T, "depth": args.depth, "ms_per_step": ms, "images_per_s": B / (ms / 1e3),
                      "loss": float(last.detach()), "eager_launches_per_step": (_lib.launch_count() - n0) // args.steps,
                      "note": "trunk forward/backward replayed from CUDA graphs (not counted in eager launches)",
                      "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}))


if __name__ == "__main__":
    main()
