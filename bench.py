#!/usr/bin/env python
"""bench.py — SAM3 ViT trunk + rank-16 LoRA training throughput (images/sec) on B200.

    python bench.py --gpus 1 --steps 5 --warmup 3                 # this repo (native sm_100a path)
    python bench.py --impl reference --steps 2 --warmup 1         # the unmodified reference trunk (baseline/_ref) on the host cores
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W    # one rank per GPU, weak scaling

A "step" is one pass of the hot path over one batch: trunk forward + backward to the LoRA
adapters (q,k,v,out,fc1,fc2, r=16, alpha=32) + [N>1: one all-reduce of the flat LoRA gradient] +
AdamW, batch 8 x 3x1008x1008 per GPU (BASELINE.json configs[1]; "1024 px" is the source size, the
model computes at 1008).  `value` is timed with the batch resident in HBM;
`e2e` goes through the public API (vit.ViT + lora_layers + autograd) with the batch copied from
pinned host memory and the loss read back every step.  Prints ONE JSON line (rank 0).  Next to the contract's keys the line
carries `roofline` (dominant kernel + the attention / conv kernels timed alone), `cpu_baseline` (the reference's own trunk on
the host, bounded sample), `gpu_eager_baseline` (the reference's own trunk on the same GPU) and `other_workloads` (the whole
detector step, native vs the untouched reference, and the other shipped adapter configs); --no-cpu / --no-gpu-eager /
--no-whole-model / --no-other-configs skip those legs.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "SAM3 ViT-L r=16 LoRA train images/sec @1024px"
# algorithmic GFLOP per image: 2*M*N*K per GEMM, no recompute, frozen base (no wgrad)
GF_ATTN_GEMM_TRAIN = 5522.8 + 130.5     # qkv/proj fwd+dgrad + 3.5x SDPA + r=16 q,k,v,o adapters
GF_TRUNK_TRAIN = 11965.0 + 130.5 + 183.5  # whole trunk train step + adapters on q,k,v,o,fc1,fc2


def read_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"burst": d.get("bf16_tflops"), "sustained": d.get("bf16_tflops_sustained"), "hbm": d.get("hbm_gbs"),
                "source": "measured"}
    return {"burst": 1590.0, "sustained": 1400.0, "hbm": 6650.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons, power = [], None, set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax = float(f[2])
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        busy = [s for s, p in zip(sm, power) if p > 300] or sm
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(power) if power else None}


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference arithmetic (oracle port) on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_unit_seconds(reps: int, warm: int):
    """Times one image through the trunk's repeating 4-block unit (3 window blocks + 1 global block,
    full width, r=16 adapters on q,k,v,out,fc1,fc2), forward + backward to the adapters, PyTorch CPU fp32,
    all host threads.  The 32-block trunk is 8 such units (+ patch embed, < 0.2 % of the FLOPs)."""
    import torch

    from oracle import vit_oracle as O

    # torchrun exports OMP_NUM_THREADS=1; the CPU arm is entitled to every host core
    ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    if torch.get_num_threads() < ncpu:
        torch.set_num_threads(ncpu)
    cfg = O.ViTConfig(depth=4, global_att_blocks=(3,))
    spec = O.LoRASpec(rank=16, alpha=32.0)
    params = O.make_params(cfg, spec, seed=0)
    g = torch.Generator().manual_seed(0)
    img = torch.randn(1, 3, 1008, 1008, generator=g)
    gout = torch.randn(1, 1024, 72, 72, generator=g) * 0.01
    times = []
    for i in range(warm + reps):
        t0 = time.perf_counter()
        O.train_step_reference(img, params, cfg, spec, gout)
        dt = time.perf_counter() - t0
        if i >= warm:
            times.append(dt)
    return times, torch.get_num_threads()


def cpu_line(times, cores, steps, warmup, kind="port"):
    t = statistics.mean(times)
    ips = 1.0 / (8.0 * t)
    return ips, {"value": ips, "unit": "images/sec", "cores": cores, "kind": kind,
                 "sample": f"1 image x 4-block unit (3 window + 1 global, D=1024, r=16 q/k/v/o/fc1/fc2 adapters) fwd+bwd, "
                           f"{len(times)} timed passes of {t:.2f} s; images/sec = 1 / (8 units x pass time)"}


def cpu_reference_arm(steps: int, warmup: int, budget_s: float):
    """The CPU baseline: the reference's OWN trunk (sam3.model.vitdet.ViT from baseline/_ref, all 32 blocks, its per-block
    activation checkpointing, its LoRALinear on mlp.fc1/fc2) on ONE image per step — kind "reference".  Falls back to the
    oracle port's 4-block unit (kind "port") only when the reference is not installed.  Returns (images/sec, cpu_baseline
    dict, mean seconds per step, timed steps)."""
    sys.path.insert(0, str(ROOT / "tools"))
    import bench_arms  # noqa: PLC0415

    if bench_arms.reference_available():
        times, cores = bench_arms.time_trunk_cpu(steps, warmup, budget_s)
        t = statistics.mean(times)
        cb = {"value": 1.0 / t, "unit": "images/sec", "cores": cores, "kind": "reference",
              "sample": f"unmodified sam3.model.vitdet.ViT (baseline/_ref): 32 blocks, 1008x1008, fp32, per-block activation "
                        f"checkpointing as shipped (vitdet.py:837-838), reference LoRALinear r=16 on mlp.fc1/fc2 (its name matching "
                        f"reaches no q/k/v/o in the fused-qkv trunk; the adapters are 1.5 % of the FLOPs), forward + backward + AdamW, "
                        f"ONE image per step ({len(times)} timed steps of {t:.1f} s within a {budget_s:.0f} s budget); a batch of 8 is "
                        f"8 such passes on a CPU, so images/sec = 1 / seconds per image"}
        return 1.0 / t, cb, t, len(times)
    times, cores = cpu_unit_seconds(max(1, min(steps, 6)), max(0, min(warmup, 1)))
    ips, cb = cpu_line(times, cores, steps, warmup)
    return ips, cb, statistics.mean(times), len(times)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ips, cb, t, n = cpu_reference_arm(max(1, args.steps), max(0, min(args.warmup, 1)), args.cpu_budget)
    line = {
        "impl": "reference", "metric": METRIC, "value": ips, "unit": "images/sec", "n_gpus": args.gpus, "steps": n,
        "steps_requested": args.steps, "warmup": max(0, min(args.warmup, 1)), "ms_per_step": 1000.0 * t, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, cpu=True),
        "cpu_baseline": cb,
        "e2e": {"value": ips, "unit": "images/sec", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, cpu=False):
    return {"workload": "SAM3 ViT trunk (32 blocks, D=1024, 16 heads, 24x24 windows + 4 global blocks, 1008x1008 compute "
                        "resolution) forward+backward with rank-16 LoRA (alpha 32, dropout 0) on q,k,v,out,fc1,fc2, AdamW on the "
                        "adapters; full_lora_config.yaml shape at r=16",
            "batch_per_gpu": args.batch, "global_batch": args.batch * (1 if cpu else args.gpus), "image": "3x1008x1008",
            "parallelism": f"dp{args.gpus}", "l2_policy": "per-step working set (>50 GB of activations) exceeds the 126 MB L2",
            "depth": args.depth, "cuda_graph": not getattr(args, "no_graph", False),
            "allreduce": "none (1 GPU)" if cpu or args.gpus == 1 else
                         f"flat LoRA gradient in {args.segments} slices, each all-reduced (NCCL, side stream) while the next block range of the backward runs"}


# ------------------------------------------------------------------------------------------------
# native arm
# ------------------------------------------------------------------------------------------------
def run_native(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from sam3_lora_b200 import _lib as L
    from sam3_lora_b200.lora_layers import LoRAConfig, apply_lora_to_model, count_parameters, get_lora_parameters
    from sam3_lora_b200.vit import ViT

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the native path has no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus or world == 1, f"WORLD_SIZE={world} but --gpus {args.gpus}"
    dt = torch.float16 if args.dtype == "float16" else torch.bfloat16
    B = args.batch

    import contextlib

    torch.manual_seed(0)
    globals_ = tuple(i for i in (7, 15, 23, 31) if i < args.depth) or (args.depth - 1,)
    model = ViT(depth=args.depth, global_att_blocks=globals_, operand_dtype=dt, max_batch=B)
    with contextlib.redirect_stdout(sys.stderr):  # stdout carries exactly one JSON line
        apply_lora_to_model(model, LoRAConfig(rank=args.rank, alpha=2 * args.rank, dropout=0.0,
                                              target_modules=["q_proj", "k_proj", "v_proj", "out_proj", "fc1", "fc2"]))
    for p in get_lora_parameters(model):  # non-zero B so every gradient is exercised
        if p.shape[0] == args.rank:
            torch.nn.init.normal_(p, std=0.02)
    model = model.to(dev)
    model.train()
    counts = count_parameters(model)

    # synthetic COCO-shaped batch: uint8 RGB noise, normalised like the reference loader (mean .5, std .5)
    rng = np.random.default_rng(rank)
    host = torch.from_numpy(((rng.integers(0, 256, size=(B, 3, 1008, 1008), dtype=np.uint8).astype(np.float32) / 255.0) - 0.5)
                            / 0.5).pin_memory()
    gout = torch.randn(B, 1024, 72, 72, device=dev) * 1e-3
    images = host.to(dev, non_blocking=True)

    # -------- device-resident step (value) --------
    feats = model(images)[0]            # builds engine, packs weights
    flat = model.flat_lora()
    gflat = torch.zeros_like(flat)
    m_buf, v_buf = torch.zeros_like(flat), torch.zeros_like(flat)
    eng = model._engine
    step_no = [0]

    out_buf = torch.empty(B, 1024, 72, 72, device=dev)

    # N > 1: the backward runs in `--segments` block ranges; each range's slice of the flat gradient is all-reduced on a side
    # stream while the next range computes (dist.LoRAGradAllReducer), so only the last slice's collective is exposed.
    from sam3_lora_b200.dist import LoRAGradAllReducer

    segs = eng.segments(args.segments if world > 1 else 1)
    reducer = LoRAGradAllReducer(average=False, segments=len(segs))      # the 1/world factor is folded into AdamW's grad scale
    ranges = [eng.grad_range(hi, lo) for hi, lo in segs]

    def part(k):
        if k == 0:
            eng.forward(images, flat, out_buf, save_for_backward=True)
        if len(segs) == 1:
            eng.backward(gout, gflat)
        else:
            eng.backward_segment(gout if k == 0 else None, gflat, segs[k][0], segs[k][1])

    def fwd_bwd():
        for k in range(len(segs)):
            part(k)

    # The ~1000 kernel launches of one forward+backward are captured once in CUDA graphs (one per backward range) and
    # replayed (all pointers are fixed, TMA descriptors are by-value kernel parameters); all-reduce and AdamW stay eager.
    graphs = None
    launches_per_fwd_bwd = None
    if not args.no_graph:
        fwd_bwd()                                   # first-use cudaFuncSetAttribute calls happen outside capture
        torch.cuda.synchronize()
        n_before = L.launch_count()
        graphs = []
        for k in range(len(segs)):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, capture_error_mode="thread_local"):
                part(k)
            graphs.append(g)
        launches_per_fwd_bwd = L.launch_count() - n_before

    def native_step():
        for k in range(len(segs)):
            if graphs is not None:
                graphs[k].replay()
            else:
                part(k)
            if world > 1:
                a, b = ranges[k]
                reducer.reduce_slice(gflat[a:b])
        if world > 1:
            reducer.finish()
        step_no[0] += 1
        L.adamw_step(flat, gflat, m_buf, v_buf, args.lr, 0.9, 0.999, 1e-8, 0.01, step_no[0], 1.0 / world)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        native_step()
    sync_all()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    n0 = L.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if args.profile_range:
        torch.cuda.profiler.start()    # ncu --profile-from-start off: only the timed steps are captured
    e0.record()
    for _ in range(args.steps):
        native_step()
    e1.record()
    sync_all()
    if args.profile_range:
        torch.cuda.profiler.stop()
    launches = (L.launch_count() - n0) // args.steps
    if launches_per_fwd_bwd is not None:
        launches += launches_per_fwd_bwd            # graph replays do not pass through the host-side counter
    ms = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_step = ms.item()
    clocks = sampler.stop() if rank == 0 else None
    value = B * world / (ms_step / 1000.0)

    # -------- end-to-end step through the public API (e2e) --------
    # ViT(cuda_graphs=True) is the public switch for graph replay of the trunk forward / backward (one eager warm-up step,
    # then two replays per step); --e2e-eager measures the same leg with ~1290 eager launches per step.
    model.cuda_graphs = not args.e2e_eager and not args.no_graph
    params = model.lora_parameters()
    if world > 1:
        model.grad_hook = LoRAGradAllReducer(average=True, segments=args.segments)   # sum / world on each slice (DDP semantics)
    opt = torch.optim.AdamW(params, lr=args.lr, weight_decay=0.01, fused=True)
    # double-buffered input: the H2D copy of step i+1 (pinned host memory, side stream) overlaps the compute of
    # step i, as an input pipeline would; every step still copies its own batch inside the timed region.
    dev_bufs = [torch.empty_like(images), torch.empty_like(images)]
    copy_stream = torch.cuda.Stream()
    copy_done = [torch.cuda.Event(), torch.cuda.Event()]
    compute_done = [torch.cuda.Event(), torch.cuda.Event()]
    state = {"i": 0}

    def issue_copy(slot):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(compute_done[slot])     # the trunk has finished reading this buffer
            dev_bufs[slot].copy_(host, non_blocking=True)
            copy_done[slot].record(copy_stream)

    for ev in compute_done:
        ev.record()
    issue_copy(0)

    def e2e_step():
        slot = state["i"] & 1
        state["i"] += 1
        issue_copy(slot ^ 1)                             # prefetch the next step's batch
        torch.cuda.current_stream().wait_event(copy_done[slot])
        f = model(dev_bufs[slot])[0]
        loss = (f * gout).sum()
        opt.zero_grad(set_to_none=True)
        loss.backward()
        compute_done[slot].record()
        opt.step()
        return loss.item()                              # D2H read of the step's loss

    for _ in range(max(3, args.warmup)):
        e2e_step()
    sync_all()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(args.steps):
        e2e_step()
    t1.record()
    sync_all()
    ms2 = torch.tensor([t0.elapsed_time(t1) / args.steps], device=dev)
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    e2e_value = B * world / (ms2.item() / 1000.0)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # -------- roofline of the dominant kernel: the tcgen05 GEMM (fc1 forward shape, GELU epilogue) --------
    peaks = read_peaks()
    M, N, K = B * 5184, 4736, 1024 + 64
    A = (torch.randn(M, K, device=dev) * 0.5).to(dt)
    W = (torch.randn(N, K, device=dev) * 0.02).to(dt)
    H = torch.empty(M, N, device=dev, dtype=dt)
    G = torch.empty(M, N + 64, device=dev, dtype=dt)
    bias = torch.zeros(N, device=dev)
    for _ in range(3):
        L.gemm(A, W, H, epilogue=L.EPI_GELU, bias=bias, C2=G)
    torch.cuda.synchronize()
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 20
    k0.record()
    for _ in range(iters):
        L.gemm(A, W, H, epilogue=L.EPI_GELU, bias=bias, C2=G)
    k1.record()
    torch.cuda.synchronize()
    kms = k0.elapsed_time(k1) / iters
    K_alg = 1024 + args.rank                      # algorithmic K: the 64-wide K-extension carries `rank` live columns
    achieved = 2.0 * M * N * K_alg / kms / 1e9
    traffic = None
    tf = ROOT / "profiles" / "gemm_traffic.json"
    if tf.exists():
        try:
            traffic = json.loads(tf.read_text()).get("dram_bytes_per_launch")
        except (OSError, ValueError):
            traffic = None
    roofline = {"bound": "tensor", "kernel": "gemm2_kernel<EPI_GELU> (CTA pair 256x256) M=%d N=%d K=%d" % (M, N, K),
                "achieved": achieved, "peak": peaks["burst"], "unit": "TFLOP/s", "frac": achieved / peaks["burst"],
                "peak_source": peaks["source"] + " cuBLAS bf16 burst", "traffic": traffic,
                "algorithmic_flops": 2.0 * M * N * K_alg, "launch_ms": kms, "executed_flops_incl_k_padding": 2.0 * M * N * K,
                "algorithmic_bytes": M * K * 2 + N * K * 2 + M * N * 2 * 2,
                "step_trunk_tflops": value / world * GF_TRUNK_TRAIN / 1e3,
                "step_trunk_frac_of_sustained": value / world * GF_TRUNK_TRAIN / 1e3 / peaks["sustained"],
                "attn_gemm_roofline_frac": value / world * GF_ATTN_GEMM_TRAIN / 1e3 / peaks["sustained"]}

    # -------- the other heavy kernels, same method (timed alone, CUDA events): context for the step breakdown --------
    def timed(fn, iters=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / iters

    others = []
    try:
        C1 = torch.empty(M, N, device=dev, dtype=dt)
        ms_plain = timed(lambda: L.gemm(A, W, C1, epilogue=L.EPI_STORE16, bias=bias))
        others.append({"kernel": "gemm2_kernel<EPI_STORE16> M=%d N=%d K=%d" % (M, N, K), "ms": ms_plain, "bound": "tensor",
                       "achieved": 2.0 * M * N * K / ms_plain / 1e9, "peak": peaks["burst"], "unit": "TFLOP/s"})
        del C1
        heads, D = 16, 1024
        for name, Ls, segs in (("window 576", 576, B * 9), ("global 5184", 5184, B)):
            T = Ls * segs
            qkv = (torch.randn(T, 3 * D + 64, device=dev) * 0.5).to(dt)
            O = torch.zeros(T, D + 64, device=dev, dtype=dt)
            lse2 = torch.zeros(heads, T, device=dev)
            dO = (torch.randn(T, D, device=dev) * 0.5).to(dt)
            delta = torch.zeros(heads, T, device=dev)
            dqkv = torch.zeros(T, 3 * D + 64, device=dev, dtype=dt)
            tab = torch.zeros(Ls, 32, 2, device=dev)
            tab[..., 0] = 1.0
            fl = 4.0 * segs * heads * Ls * Ls * 64
            ms_f = timed(lambda: L.attention_fwd(qkv, Ls, D, heads, O, lse2))
            ms_b = timed(lambda: L.attention_bwd(qkv, Ls, D, heads, O, lse2, dO, delta, dqkv, tab, Ls))
            # exp2 floor: one MUFU.EX2 warp instruction per 8 clk per SM sub-partition (tools/micro/mufu_rate.cu)
            sm_hz = (clocks or {}).get("sm_mhz") or 1700.0
            mufu_s = segs * heads * Ls * Ls / (148 * 16.0 * sm_hz * 1e6)
            others.append({"kernel": "attn_fwd_kernel, %s, B=%d" % (name, B), "ms": ms_f, "bound": "mufu-ex2", "achieved": fl / ms_f / 1e9,
                           "unit": "TFLOP/s", "floor_ms": mufu_s * 1e3, "frac_of_floor": mufu_s * 1e3 / ms_f})
            others.append({"kernel": "attn_bwd_dkdv+dq kernels, %s, B=%d" % (name, B), "ms": ms_b, "bound": "mufu-ex2 (2 exp per score)",
                           "achieved": 2.5 * fl / ms_b / 1e9, "unit": "TFLOP/s (5 algorithmic matmuls)", "floor_ms": 2 * mufu_s * 1e3,
                           "frac_of_floor": 2 * mufu_s * 1e3 / ms_b})
            del qkv, O, lse2, dO, delta, dqkv
            # the library kernel on the same shape (torch SDPA, fp16, its flash backend): context only, never on the product path
            try:
                import torch.nn.functional as F   # noqa: PLC0415

                q_, k_, v_ = ((torch.randn(segs, heads, Ls, 64, device=dev) * 0.5).to(dt).requires_grad_(True) for _ in range(3))
                go_ = (torch.randn(segs, heads, Ls, 64, device=dev) * 0.5).to(dt)
                with torch.no_grad():
                    lib_f = timed(lambda: F.scaled_dot_product_attention(q_, k_, v_))
                o_ = F.scaled_dot_product_attention(q_, k_, v_)
                lib_b = timed(lambda: torch.autograd.grad(o_, (q_, k_, v_), go_, retain_graph=True))
                others[-2]["library_sdpa_ms"] = lib_f
                others[-1]["library_sdpa_ms"] = lib_b
                del q_, k_, v_, go_, o_
            except Exception as e:   # noqa: BLE001
                others[-1]["library_sdpa_error"] = f"{type(e).__name__}: {e}"[:200]
        # row a7: the nn.MultiheadAttention core at the DETR sizes (E=256, 8 heads of 32 run as zero-padded 64-wide heads):
        # encoder self-attention over the 5184 image tokens with and without attention dropout, decoder image cross-attention
        # 401 x 5184 with the additive fp32 box-RPB bias
        H7, Ep = 8, 8 * 64
        for name, Lq, Lk, pd, with_bias in (("encoder self-attn 5184x5184, dropout 0.1", 5184, 5184, 0.1, False),
                                            ("encoder self-attn 5184x5184, dropout 0", 5184, 5184, 0.0, False),
                                            ("decoder image cross-attn 401x5184 + box-RPB bias", 401, 5184, 0.0, True)):
            q7 = (torch.randn(B * Lq, Ep, device=dev) * 0.5).to(dt)
            kv7 = (torch.randn(B * Lk, 2 * Ep, device=dev) * 0.5).to(dt)
            O7 = torch.empty(B * Lq, Ep, device=dev, dtype=dt)
            Ls7 = (Lq + 63) // 64 * 64
            lse7 = torch.zeros(H7, B * Ls7, device=dev)
            bias7 = torch.randn(B * H7, Lq, Lk, device=dev) if with_bias else None
            d7 = L.mha_desc(q7, kv7, B, Lq, Lk, H7, 32 ** -0.5, O7, lse7, bias=bias7, drop_p=pd, drop_seed=123)
            keep7 = []

            def fwd7():
                keep7[:] = [L.attention_dropout_bits(d7)]       # large problems with dropout: keep-bits generated per call (timed)
                L.mha_fwd(d7)

            ms_f = timed(fwd7)
            ms_bits = timed(lambda: L.attention_dropout_bits(d7)) if pd > 0 else 0.0
            dO7 = (torch.randn(B * Lq, Ep, device=dev) * 0.5).to(dt)
            delta7 = torch.zeros_like(lse7)
            dq7 = torch.empty_like(q7)
            dkv7 = torch.empty_like(kv7)
            ms_b = timed(lambda: L.mha_bwd(d7, dO7, delta7, dq7, dkv7))
            fl7 = 4.0 * B * H7 * Lq * Lk * 32            # algorithmic: head_dim 32
            others.append({"kernel": "row a7 MHA core (attn_fwd / attn_bwd GEN), %s, B=%d" % (name, B), "ms_fwd": ms_f, "ms_fwd_of_which_mask_bits": ms_bits, "ms_bwd": ms_b,
                           "ms": ms_f + ms_b, "bound": "mufu-ex2 + per-score dropout hash" if pd > 0 else "mufu-ex2",
                           "achieved": 3.5 * fl7 / (ms_f + ms_b) / 1e9, "unit": "TFLOP/s (algorithmic, head_dim 32; executed at 64)"})
            del q7, kv7, O7, lse7, bias7, dO7, delta7, dq7, dkv7
        # row a8's dominant kernel: the im2col-free 3x3 convolution at the 288x288 level (256 -> 256 channels)
        from sam3_lora_b200 import conv_ops as CO  # noqa: PLC0415
        xc = (torch.randn(B, 288, 288, 256, device=dev) * 0.5).to(dt)
        w9 = (torch.randn(256, 9 * 256, device=dev) * 0.02).to(dt)
        ms_c = timed(lambda: CO.conv3x3(xc, w9, None, out_f32=False))
        flc = 2.0 * B * 288 * 288 * 9 * 256 * 256
        others.append({"kernel": "conv3x3_kernel (implicit GEMM, 4-D TMA boxes) B=%d 288x288 256->256" % B, "ms": ms_c, "bound": "tensor",
                       "achieved": flc / ms_c / 1e9, "peak": peaks["burst"], "unit": "TFLOP/s"})
        del xc, w9
    except Exception as e:  # noqa: BLE001 - context only, never fatal for the bench line
        others.append({"error": f"{type(e).__name__}: {e}"[:200]})
    roofline["other_kernels"] = others

    # -------- reference arms beside it (rank 0, N=1 only) --------
    cpu_baseline = None
    gpu_eager = None
    whole = None
    other_cfgs = None
    if world == 1:
        del A, W, H, G
        torch.cuda.empty_cache()
        sys.path.insert(0, str(ROOT / "tools"))
        import bench_arms  # noqa: PLC0415

        have_ref = bench_arms.reference_available()
        if have_ref and not args.no_gpu_eager:
            # the reference's own GPU path on the same device: the real bar for the trunk step
            try:
                ms_ref, what = bench_arms.time_trunk_gpu(B, 3, 2, rank=args.rank, device=dev)
                gpu_eager = {"value": B / ms_ref * 1e3, "unit": "images/sec", "ms_per_step": ms_ref, "kind": "reference", "what": what,
                             "native_over_reference": value / (B / ms_ref * 1e3)}
            except Exception as e:  # noqa: BLE001
                gpu_eager = {"error": f"{type(e).__name__}: {e}"[:300]}
        if have_ref and not args.no_whole_model:
            # BASELINE.md section 3 variant (ii): the whole detector step, native swaps vs the untouched reference, same GPU
            whole = {"workload": f"Sam3Image training step (forward, Hungarian matching, Sam3LossWrapper objective, backward, AdamW), "
                                 f"batch {B}, r={args.rank} adapters on the trunk's fc1/fc2 (the set both sides express), dropout / "
                                 f"DropPath on, synthetic COCO-shaped batch resident on the device"}
            for key, native in (("native", True), ("reference_gpu_eager", False)):
                try:
                    ms_w, info = bench_arms.time_whole_model(native, B, 3, 3, rank=args.rank, device=dev)
                    whole[key] = {"value": B / ms_w * 1e3, "unit": "images/sec", "ms_per_step": ms_w, **info}
                except Exception as e:  # noqa: BLE001
                    whole[key] = {"error": f"{type(e).__name__}: {e}"[:300]}
            if "value" in whole.get("native", {}) and "value" in whole.get("reference_gpu_eager", {}):
                whole["native_over_reference"] = whole["native"]["value"] / whole["reference_gpu_eager"]["value"]
        if not args.no_other_configs:
            # the other shipped shapes of BASELINE.json at the per-GPU level (the trunk is the same 32-block ViT in every
            # config; rank, per-GPU batch and adapter dropout differ), through the public API with graph replay
            other_cfgs = []
            for name, rk, bb, dp in (("full_lora_config.yaml as shipped: r=32, alpha 64, lora.dropout 0.1, batch 8", 32, 8, 0.1),
                                     ("configs[3] shape: r=32 on q,k,v,o,fc1,fc2, batch 4 per GPU (16 over 4 GPUs), dropout 0.1", 32, 4, 0.1),
                                     ("configs[4] shape (crack_detection_config.yaml as BASELINE.json quotes it): r=8, batch 4 per GPU (32 over 8 GPUs), dropout 0.1", 8, 4, 0.1)):
                try:
                    ms_c, n_c = bench_arms.time_trunk_native(rk, bb, dp, 5, 3, device=dev)
                    other_cfgs.append({"config": name, "value": bb / ms_c * 1e3, "unit": "images/sec per GPU", "ms_per_step": ms_c,
                                       "trainable_parameters": n_c, "path": "vit.ViT(cuda_graphs=True) + autograd + fused torch AdamW"})
                except Exception as e:  # noqa: BLE001
                    other_cfgs.append({"config": name, "error": f"{type(e).__name__}: {e}"[:300]})
        if not args.no_cpu:
            _, cpu_baseline, _, _ = cpu_reference_arm(1, 0, 60.0)     # one image through the reference trunk on the host cores

    line = {
        "metric": METRIC, "value": value, "unit": "images/sec", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16" if dt == torch.float16 else "bf16", "data": "synthetic",
        "config": workload_config(args),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "images/sec", "h2d_bytes_per_step": host.numel() * 4 * 1, "d2h_bytes_per_step": 4,
                "ms_per_step": ms2.item()},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "cpu_baseline": cpu_baseline,
        "gpu_eager_baseline": gpu_eager,
        "other_workloads": {"sam3_whole_model_step": whole, "trunk_step_other_configs": other_cfgs}
                           if (whole is not None or other_cfgs is not None) else None,
        "trainable_parameters": counts["trainable_parameters"],
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--batch", type=int, default=8, help="images per GPU")
    ap.add_argument("--depth", type=int, default=32, help="trunk depth (32 = SAM3; smaller only for debugging)")
    ap.add_argument("--rank", type=int, default=16)
    ap.add_argument("--dtype", default="float16", choices=["float16", "bfloat16"],
                    help="tensor-core operand format (fp32 accumulate, fp32 residual stream)")
    ap.add_argument("--lr", type=float, default=5e-5)
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    ap.add_argument("--no-gpu-eager", action="store_true", help="skip the reference's GPU-eager trunk leg (needs baseline/_ref)")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the trunk step at the other shipped ranks / batches")
    ap.add_argument("--no-whole-model", action="store_true", help="skip the whole-detector step legs (needs baseline/_ref)")
    ap.add_argument("--cpu-budget", type=float, default=240.0,
                    help="--impl reference: wall-clock budget in seconds; the arm stops after the step that exceeds it")
    ap.add_argument("--profile-range", action="store_true",
                    help="bracket the timed steps with cudaProfilerStart/Stop (use with `ncu --profile-from-start off`)")
    ap.add_argument("--segments", type=int, default=4,
                    help="N > 1: block ranges of the backward whose gradient slices are all-reduced while the next range runs")
    ap.add_argument("--no-graph", action="store_true", help="launch kernels eagerly instead of replaying a CUDA graph")
    ap.add_argument("--e2e-eager", action="store_true", help="run the end-to-end leg without ViT(cuda_graphs=True)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    main()
