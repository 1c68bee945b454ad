"""Reference-side arms of bench.py (measurement infrastructure, not product code).

Everything here runs the UNMODIFIED reference installed under baseline/_ref (tools/install_reference.sh), imported through
sam3_lora_b200.sam3_bridge (which only adds import-time stand-ins for third-party packages this image lacks):

  * reference_trunk(...)        sam3.model.vitdet.ViT exactly as sam3/model_builder.py:69-96 builds it (per-block activation
                                checkpointing in training, vitdet.py:837-838) + the reference's own lora_layers on mlp.fc1/fc2
                                (the only trunk Linears its name matching reaches);
  * time_trunk_cpu(...)         that module, PyTorch CPU fp32, all host threads, one image per step  -> cpu_baseline / --impl reference
  * time_trunk_gpu(...)         that module on the B200: TF32 matmuls (model_builder.py:46-55) + SDPA  -> gpu_eager_baseline
  * time_whole_model(...)       BASELINE.md section 3 variant (ii): Sam3Image forward + the trainer's objective + backward +
                                AdamW, for the reference model (eager) and for the same model with the native modules swapped in.
"""
from __future__ import annotations

import contextlib
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def reference_available() -> bool:
    from sam3_lora_b200 import sam3_bridge

    return sam3_bridge.reference_root() is not None


def reference_trunk(rank: int = 16, device: str = "cpu"):
    """(holder module, trunk, adapter parameter list); adapter B factors are made non-zero so every gradient is exercised."""
    import torch
    import torch.nn as nn

    from sam3_lora_b200 import sam3_bridge

    mb = sam3_bridge.import_reference()
    import lora_layers as ref_lora  # the reference's root-level module

    torch.manual_seed(0)
    with sam3_bridge.cpu_compat():
        trunk = mb._create_vit_backbone()
    holder = nn.Module()
    holder.vision_backbone = nn.Module()
    holder.vision_backbone.trunk = trunk
    cfg = ref_lora.LoRAConfig(rank=rank, alpha=2 * rank, dropout=0.0, target_modules=["q_proj", "k_proj", "v_proj", "out_proj", "fc1", "fc2"])
    with contextlib.redirect_stdout(sys.stderr):
        ref_lora.apply_lora_to_model(holder, cfg)
    params = ref_lora.get_lora_parameters(holder)
    for p in params:
        if p.shape[0] == rank:
            nn.init.normal_(p, std=0.02)
    holder.to(device)
    holder.train()
    return holder, trunk, params


def _all_threads():
    import torch

    # torchrun exports OMP_NUM_THREADS=1; the CPU arm is entitled to every host core
    ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    if torch.get_num_threads() < ncpu:
        torch.set_num_threads(ncpu)
    return torch.get_num_threads()


def time_trunk_cpu(steps: int, warmup: int, budget_s: float, rank: int = 16):
    """One image per step through the reference's 32-block trunk, forward + backward to the adapters + AdamW, on the host.
    Stops early when `budget_s` of wall clock is used up (at least one timed step).  Returns (seconds per step list, cores)."""
    import torch

    cores = _all_threads()
    holder, trunk, params = reference_trunk(rank, "cpu")
    opt = torch.optim.AdamW(params, lr=5e-5, weight_decay=0.01)
    g = torch.Generator().manual_seed(0)
    img = torch.randn(1, 3, 1008, 1008, generator=g)
    gout = torch.randn(1, 1024, 72, 72, generator=g) * 1e-3
    times = []
    t_start = time.perf_counter()
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        feat = trunk(img)[-1]
        loss = (feat * gout).sum()
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
        elapsed = time.perf_counter() - t_start
        if times and elapsed + dt > budget_s:
            break
        if not times and i + 1 >= warmup and elapsed > budget_s:      # warm-up alone ate the budget: time one step anyway
            warmup = i + 1
    return times, cores


def time_trunk_gpu(batch: int, steps: int, warmup: int, rank: int = 16, device="cuda"):
    """The reference trunk on the GPU as the reference runs it: fp32 parameters, TF32 matmuls, F.scaled_dot_product_attention,
    per-block activation checkpointing, its own LoRALinear (3 extra kernels per adapted Linear), torch AdamW.
    Returns (ms per step, description)."""
    import torch

    holder, trunk, params = reference_trunk(rank, device)
    opt = torch.optim.AdamW(params, lr=5e-5, weight_decay=0.01, fused=True)
    img = torch.randn(batch, 3, 1008, 1008, device=device)
    gout = torch.randn(batch, 1024, 72, 72, device=device) * 1e-3

    def step():
        feat = trunk(img)[-1]
        loss = (feat * gout).sum()
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    what = (f"unmodified sam3.model.vitdet.ViT from baseline/_ref on the same GPU, batch {batch}: fp32 + TF32 matmuls "
            f"(allow_tf32={torch.backends.cuda.matmul.allow_tf32}), SDPA, per-block activation checkpointing, reference LoRALinear "
            f"r={rank} on mlp.fc1/fc2 (its name matching reaches no q/k/v/o in the fused-qkv trunk), fused torch AdamW; "
            f"{steps} timed steps after {warmup} warm-up")
    del holder, trunk, params, opt
    torch.cuda.empty_cache()
    return ms, what


def time_trunk_native(rank: int, batch: int, dropout: float, steps: int, warmup: int, device="cuda",
                      targets=("q_proj", "k_proj", "v_proj", "out_proj", "fc1", "fc2")):
    """The native trunk step through the public API (vit.ViT(cuda_graphs=True) + lora_layers + autograd + fused torch AdamW,
    batch resident on the device) for another shipped configuration: rank / per-GPU batch / adapter dropout.  Returns
    (ms per step, trainable parameter count)."""
    import torch

    from sam3_lora_b200.lora_layers import LoRAConfig, apply_lora_to_model, get_lora_parameters
    from sam3_lora_b200.vit import ViT

    torch.manual_seed(0)
    model = ViT(max_batch=batch, cuda_graphs=True)
    with contextlib.redirect_stdout(sys.stderr):
        apply_lora_to_model(model, LoRAConfig(rank=rank, alpha=2 * rank, dropout=dropout, target_modules=list(targets)))
    for p in get_lora_parameters(model):
        if p.shape[0] == rank:
            torch.nn.init.normal_(p, std=0.02)
    model = model.to(device).train()
    img = torch.randn(batch, 3, 1008, 1008, device=device)
    gout = torch.randn(batch, 1024, 72, 72, device=device) * 1e-3
    model(img)                                    # builds the engine, flattens the adapters
    params = model.lora_parameters()
    opt = torch.optim.AdamW(params, lr=5e-5, weight_decay=0.01, fused=True)

    def step():
        f = model(img)[0]
        loss = (f * gout).sum()
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()

    for _ in range(max(3, warmup)):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    n = sum(p.numel() for p in params)
    del model, opt, params
    torch.cuda.empty_cache()
    return ms, n


def time_whole_model(native: bool, batch: int, steps: int, warmup: int, rank: int = 16, device="cuda", stochastic: bool = True):
    """One SAM3 training step of the detector (forward, matcher, Sam3LossWrapper objective, backward, AdamW) at `batch`
    images, synthetic COCO-shaped batch resident on the device.  native=False: the reference model untouched, its SciPy
    matcher, its losses (Triton focal kernels if they run here, else its eager branch).  native=True: sam3_bridge swaps
    + GPU matcher + fused losses.  Adapters: fc1/fc2 of the trunk (the set both sides can express).  Returns (ms, info)."""
    import torch

    from sam3_lora_b200 import sam3_bridge, sam3_step

    torch.manual_seed(0)
    model = sam3_bridge.build_reference_model("cpu", seed=0)
    info = {}
    if not native:
        sam3_bridge.restore_activation_checkpointing()       # the reference arm runs the reference's own recompute wrapper
    if native:
        from sam3_lora_b200.lora_layers import LoRAConfig, apply_lora_to_model

        model = sam3_bridge.build_native_model(device, reference_model=model, max_batch=batch)
        cfg = LoRAConfig(rank=rank, alpha=2 * rank, dropout=0.0, target_modules=["fc1", "fc2"], apply_to_text_encoder=False,
                         apply_to_detr_encoder=False, apply_to_detr_decoder=False, strict_reference_names=True)
        with contextlib.redirect_stdout(sys.stderr):
            model = apply_lora_to_model(model, cfg).to(device)
    else:
        import lora_layers as ref_lora

        cfg = ref_lora.LoRAConfig(rank=rank, alpha=2 * rank, dropout=0.0, target_modules=["fc1", "fc2"], apply_to_text_encoder=False,
                                  apply_to_detr_encoder=False, apply_to_detr_decoder=False)
        with contextlib.redirect_stdout(sys.stderr):
            model = ref_lora.apply_lora_to_model(model, cfg).to(device)
    params = [p for p in model.parameters() if p.requires_grad]
    for p in params:
        if p.shape[0] == rank:
            torch.nn.init.normal_(p, std=0.02)
    model.train()
    if not stochastic:
        sam3_step.disable_stochastic(model)
    matcher, wrapper = sam3_step.build_objective(native=native)
    opt = torch.optim.AdamW(params, lr=5e-5, weight_decay=0.01, fused=True)
    batch_host = sam3_step.collate(sam3_step.synthetic_datapoints(batch, seed=0))
    data = sam3_step.move_to_device(batch_host, device)
    losses = []

    def step():
        loss, _ = sam3_step.training_loss(model, data, matcher, wrapper)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        losses.append(loss.detach())

    try:
        step()
    except Exception as e:  # noqa: BLE001 - the reference's Triton focal loss may not run on this image / GPU
        if native:
            raise
        info["triton_focal"] = f"failed ({type(e).__name__}); eager focal branch timed instead"
        from sam3.train.loss import loss_fns

        ref_focal = loss_fns._sam3b_ref_focal
        loss_fns.sigmoid_focal_loss = lambda i, t, n, *a, **kw: ref_focal(i, t, n, *a, **{**kw, "triton": False})
        step()
    for _ in range(max(0, warmup - 1)):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    info.update({"loss_first": float(losses[0]), "loss_last": float(losses[-1]), "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30,
                 "trainable": sum(p.numel() for p in params)})
    del model, opt, data
    torch.cuda.empty_cache()
    return ms, info


if __name__ == "__main__":
    import argparse
    import json

    ap = argparse.ArgumentParser()
    ap.add_argument("what", choices=["trunk_cpu", "trunk_gpu", "whole_native", "whole_reference"])
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=2)
    a = ap.parse_args()
    if a.what == "trunk_cpu":
        t, c = time_trunk_cpu(a.steps, a.warmup, 600.0)
        print(json.dumps({"what": a.what, "s_per_image": t, "cores": c}))
    elif a.what == "trunk_gpu":
        ms, what = time_trunk_gpu(a.batch, a.steps, a.warmup)
        print(json.dumps({"what": a.what, "ms_per_step": ms, "images_per_sec": a.batch / ms * 1e3, "desc": what}))
    else:
        ms, info = time_whole_model(a.what == "whole_native", a.batch, a.steps, a.warmup)
        print(json.dumps({"what": a.what, "ms_per_step": ms, "images_per_sec": a.batch / ms * 1e3, **info}))
