"""Times the fused mask loss (row f1) and the GPU matcher (row f2) at SAM3's sizes on one GPU, with their HBM roofline /
CPU-oracle context, and prints one JSON line each.  Not the headline bench (bench.py).

    python tools/bench_loss_matcher.py [--masks 64] [--images 48]
"""
from __future__ import annotations

import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))


def timed(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--masks", type=int, default=64)
    ap.add_argument("--images", type=int, default=48, help="matcher problems per call (8 images x 6 decoder layers)")
    args = ap.parse_args()
    from oracle import loss_oracle as LO, matcher_oracle as MO
    from sam3_lora_b200.losses import mask_losses
    from sam3_lora_b200.matcher import BinaryHungarianMatcherV2

    peaks = {}
    try:
        peaks = json.loads((Path(__file__).resolve().parents[1] / "MEASURED_PEAKS.json").read_text())
    except OSError:
        pass
    hbm = float(peaks.get("hbm_gbs", 6590.0))
    g = torch.Generator().manual_seed(0)
    N, h, H = args.masks, 288, 1008
    src = (torch.randn(N, h, h, generator=g) * 4).cuda().requires_grad_(True)
    tgt = torch.nn.functional.interpolate(torch.rand(N, 1, 36, 36, generator=g).gt(0.6).float(), size=(H, H))[:, 0].bool().cuda()
    gsum = torch.ones(2, device="cuda")

    def fwd():
        return mask_losses(src, tgt, 8.0)

    def fwd_bwd():
        src.grad = None
        o = mask_losses(src, tgt, 8.0)
        (o["loss_mask"] + o["loss_dice"]).backward()

    f_ms = timed(fwd)
    fb_ms = timed(fwd_bwd)
    alg_fwd = N * (H * H * 1 + h * h * 4)                      # uint8 targets + fp32 logits, read once
    alg_bwd = N * (H * H * 1 + 2 * h * h * 4)
    # eager PyTorch on the same GPU (what the reference does without its Triton kernels): interpolate + focal + dice
    def torch_ref():
        s = src.detach().clone().requires_grad_(True)
        o = LO.mask_losses(s, tgt, 8.0)
        (o["loss_mask"] + o["loss_dice"]).backward()
    t_ms = timed(torch_ref, iters=5, warm=2)
    print(json.dumps({"what": "f1 fused up-sample + focal + dice", "masks": N, "src": [h, h], "target": [H, H], "fwd_ms": f_ms,
                      "fwd_bwd_ms": fb_ms, "fwd_GBps_algorithmic": alg_fwd / f_ms / 1e6, "bwd_GBps_algorithmic": alg_bwd / max(fb_ms - f_ms, 1e-6) / 1e6,
                      "hbm_peak_GBps": hbm, "fwd_frac_of_hbm": alg_fwd / f_ms / 1e6 / hbm, "torch_eager_same_math_fwd_bwd_ms": t_ms}))

    # ---- matcher ----
    B, Q, Tmax = args.images, 200, 32
    nb = torch.randint(1, Tmax + 1, (B,), generator=g)
    pb = torch.cat([torch.rand(B, Q, 2, generator=g) * 0.8 + 0.1, torch.rand(B, Q, 2, generator=g) * 0.3 + 0.02], -1).cuda()
    tb = torch.cat([torch.rand(B, Tmax, 2, generator=g) * 0.8 + 0.1, torch.rand(B, Tmax, 2, generator=g) * 0.3 + 0.02], -1).cuda()
    lg = (torch.randn(B, Q, generator=g) * 2).cuda()
    m = BinaryHungarianMatcherV2(focal=True, cost_class=2.0, cost_bbox=5.0, cost_giou=2.0)
    m_ms = timed(lambda: m.match(lg, pb, tb, nb, 1))
    outs = {"pred_logits": lg[:, :, None], "pred_boxes": pb}
    tg = {"boxes_padded": tb, "num_boxes": nb}
    api_ms = timed(lambda: m(outs, tg))
    C = MO.cost_matrix(lg.cpu().numpy(), pb.cpu().numpy(), tb.cpu().numpy(), 2.0, 5.0, 2.0, True)
    t0 = time.perf_counter()
    for _ in range(5):
        Cc = MO.cost_matrix(lg.cpu().numpy(), pb.cpu().numpy(), tb.cpu().numpy(), 2.0, 5.0, 2.0, True)
        MO.match(Cc, nb.tolist(), 1)
    cpu_ms = (time.perf_counter() - t0) / 5 * 1e3
    print(json.dumps({"what": "f2 GPU Hungarian matcher", "problems": B, "queries": Q, "targets_max": Tmax, "kernels_ms": m_ms,
                      "forward_api_ms": api_ms, "cpu_oracle_numpy_scipy_ms": cpu_ms}))


if __name__ == "__main__":
    main()
