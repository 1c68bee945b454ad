"""Times row a8 (SimpleFPN neck -> pixel decoder -> instance 1x1 -> mask einsum) forward + backward at SAM3's sizes on one GPU
and prints one JSON line (CUDA events on the current stream; inputs resident in HBM).  Not the headline bench (bench.py): this
is the measurement for the a8 row of DESIGN.md.

    python tools/bench_seg.py [--batch 8] [--queries 200] [--iters 5]
"""
from __future__ import annotations

import argparse
import json
import sys
from pathlib import Path

import torch
import torch.nn as nn

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))


class _Trunk(nn.Module):
    channel_list = [1024]

    def forward(self, x):
        return [x]


class _NoPos(nn.Module):
    def forward(self, x):
        return x.new_zeros(1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--queries", type=int, default=200)
    ap.add_argument("--grid", type=int, default=72)
    ap.add_argument("--iters", type=int, default=5)
    args = ap.parse_args()
    from sam3_lora_b200 import _lib, conv_ops as CO
    from sam3_lora_b200.maskformer_segmentation import PixelDecoder, UniversalSegmentationHead
    from sam3_lora_b200.necks import Sam3DualViTDetNeck

    dev = "cuda:0"
    torch.manual_seed(0)
    B, Q, G, d = args.batch, args.queries, args.grid, 256
    neck = Sam3DualViTDetNeck(_Trunk(), _NoPos(), d_model=d, scale_factors=(4.0, 2.0, 1.0)).to(dev)
    head = UniversalSegmentationHead(d, 2, PixelDecoder(d, 2)).to(dev)
    for m in (neck, head):
        for p in m.parameters():
            p.requires_grad_(False)
    x = torch.randn(B, 1024, G, G, device=dev, requires_grad=True)
    q = torch.randn(B, Q, d, device=dev, requires_grad=True)

    def step():
        feats = neck(x)[0]
        pix = head.pixel_decoder(feats)
        masks = head.mask_predictor(q, CO.conv1x1_forward(pix, head.instance_seg_head))
        return masks

    cot = None
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    fwd_ms, bwd_ms, launches = [], [], 0
    for it in range(args.iters + 2):
        x.grad = q.grad = None
        n0 = _lib.launch_count()
        ev[0].record()
        masks = step()
        ev[1].record()
        if cot is None:
            cot = torch.randn_like(masks) / masks.numel()
        masks.backward(cot)
        ev[2].record()
        torch.cuda.synchronize()
        launches = _lib.launch_count() - n0
        if it >= 2:
            fwd_ms.append(ev[0].elapsed_time(ev[1])); bwd_ms.append(ev[1].elapsed_time(ev[2]))
        del masks
    hw = [(4 * G) ** 2, (2 * G) ** 2, G * G]
    # algorithmic forward FLOPs per image (2*M*N*K): neck deconvs / 1x1 / 3x3, decoder 3x3 x2, instance 1x1, einsum
    fl = 0
    fl += 2 * hw[2] * 1024 * 2048 + 2 * hw[1] * 512 * 1024 + 2 * hw[0] * 256 * d + 2 * hw[0] * 9 * d * d      # scale 4
    fl += 2 * hw[2] * 1024 * 2048 + 2 * hw[1] * 512 * d + 2 * hw[1] * 9 * d * d                                # scale 2
    fl += 2 * hw[2] * 1024 * d + 2 * hw[2] * 9 * d * d                                                         # scale 1
    fl += 2 * hw[1] * 9 * d * d + 2 * hw[0] * 9 * d * d + 2 * hw[0] * d * d + 2 * hw[0] * d * Q
    f, b = sum(fwd_ms) / len(fwd_ms), sum(bwd_ms) / len(bwd_ms)
    print(json.dumps({"what": "a8 neck + pixel decoder + mask einsum, fwd+bwd", "batch": B, "queries": Q, "grid": G,
                      "fwd_ms": f, "bwd_ms": b, "images_per_s": B / ((f + b) / 1e3), "launches_per_step": launches,
                      "fwd_gflop_per_image": fl / 1e9, "fwd_tflops": B * fl / (f / 1e3) / 1e12,
                      "bwd_tflops_dgrad_only": B * fl / (b / 1e3) / 1e12,
                      "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}))


if __name__ == "__main__":
    main()
