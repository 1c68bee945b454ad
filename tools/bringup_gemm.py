"""GPU bring-up for the tcgen05 GEMM (run under gpurun; one subprocess per case so that a
trapping kernel cannot poison the CUDA context of the other cases).

    python tools/bringup_gemm.py            # runs every case, appends JSON lines to gpurun_out/bringup_gemm.jsonl
    python tools/bringup_gemm.py --case X   # one case in this process
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
OUT = ROOT / "gpurun_out"


def rel_err(got, ref):
    import torch

    got = got.float()
    ref = ref.float()
    if not torch.isfinite(got).all():
        return float("inf")
    return ((got - ref).abs().max() / ref.abs().max().clamp_min(1e-20)).item()


def rope_table(period, dev):
    import torch

    ang = torch.rand(period, 32, device=dev) * 6.28
    return torch.stack([ang.cos(), ang.sin()], dim=-1).contiguous()  # [period,32,2]


def apply_rope_ref(x, table, period, rope_cols):
    """x: [M,N] fp32; rotate adjacent pairs of the first rope_cols columns; head width 64."""
    import torch

    M, N = x.shape
    out = x.clone()
    rows = torch.arange(M, device=x.device) % period
    cs = table[rows]  # [M,32,2]
    xr = x[:, :rope_cols].reshape(M, rope_cols // 64, 32, 2)
    a, b = xr[..., 0], xr[..., 1]
    c, s = cs[:, None, :, 0], cs[:, None, :, 1]
    out[:, :rope_cols] = torch.stack([a * c - b * s, a * s + b * c], dim=-1).reshape(M, rope_cols)
    return out


def run_case(name: str) -> dict:
    import torch

    from sam3_lora_b200 import _lib as L

    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    dev = "cuda"
    res = {"case": name}

    def mk(M, K, dt, scale=1.0):
        return (torch.randn(M, K, device=dev) * scale).to(dt)

    def basic(M, N, K, dt=torch.float16, epi=L.EPI_STORE32, bn=0, **kw):
        A, B = mk(M, K, dt), mk(N, K, dt)
        out_dt = torch.float32 if epi in (L.EPI_STORE32,) else dt
        Cb = torch.zeros(M, N, device=dev, dtype=out_dt)
        L.gemm(A, B, Cb, epilogue=epi, bn=bn, cta_pair=pair, **kw)
        torch.cuda.synchronize()
        ref = A.float() @ B.float().T
        return rel_err(Cb, ref)

    pair = 0
    if name.startswith("p_") or name.startswith("perfp_"):   # same cases on the CTA-pair (cta_group::2) kernel
        pair = 2
        name = ("k_" + name[2:]) if name.startswith("p_") else ("perf_" + name[6:])
    if name == "k_tile1":
        res["err"] = basic(256 if pair else 128, 256, 64)
    elif name == "k_tile1_bn64":
        res["err"] = basic(128, 64, 64, bn=64)
    elif name == "k_multi":
        res["err"] = basic(384, 512, 256)
    elif name == "k_ragged":
        res["err"] = basic(200, 328, 80)  # M, N not tile multiples; K = 64 + 16
    elif name == "k_big_f16":
        res["err"] = basic(5184, 3072, 1088)
    elif name == "k_big_bf16":
        res["err"] = basic(5184, 3072, 1088, dt=torch.bfloat16)
    elif name == "k_store16":
        res["err"] = basic(640, 1024, 320, epi=L.EPI_STORE16)
    elif name == "k_skinny":
        res["err"] = basic(5184, 64, 1024, epi=L.EPI_STORE16, bn=64)
    elif name == "k_skinny_full":   # the trunk's size: 324 row tiles over 148 persistent CTAs
        res["err"] = basic(41472, 64, 1024, epi=L.EPI_STORE16, bn=64)
    elif name == "k_skinny_ragged":   # M not a multiple of 128 nor 8; N = 48 (q|k|v of one rank-16 adapter)
        res["err"] = basic(20736 + 1003, 48, 1024 + 16, epi=L.EPI_STORE16, bn=64)
    elif name == "k_persist":  # more tiles than CTAs: exercises phase wrap-around on every barrier
        res["err"] = basic(128 * 40, 256 * 12, 64 * 9, max_ctas=7)
    elif name == "epi_residual":
        M, N, K = 640, 1024, 256
        A, B = mk(M, K, torch.float16), mk(N, K, torch.float16)
        bias = torch.randn(N, device=dev)
        resid = torch.randn(160, N, device=dev)
        Cb = torch.zeros(M, N, device=dev)
        L.gemm(A, B, Cb, epilogue=L.EPI_RESIDUAL_F32, bias=bias, residual=resid, res_row_mod=160)
        torch.cuda.synchronize()
        ref = A.float() @ B.float().T + bias + resid.repeat(4, 1)
        res["err"] = rel_err(Cb, ref)
    elif name == "epi_gelu":
        M, N, K = 384, 608, 128
        A, B = mk(M, K, torch.float16, 0.3), mk(N, K, torch.float16, 0.3)
        bias = torch.randn(N, device=dev)
        H = torch.zeros(M, N, device=dev, dtype=torch.float16)
        G = torch.zeros(M, N + 64, device=dev, dtype=torch.float16)
        L.gemm(A, B, H, epilogue=L.EPI_GELU, bias=bias, C2=G)
        torch.cuda.synchronize()
        h = A.float() @ B.float().T + bias
        res["err_h"] = rel_err(H, h)
        res["err"] = rel_err(G[:, :N], torch.nn.functional.gelu(h))
        res["pad_untouched"] = bool((G[:, N:] == 0).all().item())
    elif name == "epi_dgelu":
        M, N, K = 384, 608, 128
        A, B = mk(M, K, torch.float16, 0.3), mk(N, K, torch.float16, 0.3)
        h = torch.randn(M, N, device=dev).to(torch.float16)
        D = torch.zeros(M, N, device=dev, dtype=torch.float16)
        L.gemm(A, B, D, epilogue=L.EPI_DGELU, aux=h)
        torch.cuda.synchronize()
        hf = h.float().requires_grad_(True)
        torch.nn.functional.gelu(hf).sum().backward()
        ref = (A.float() @ B.float().T) * hf.grad
        res["err"] = rel_err(D, ref)
    elif name == "epi_rope":
        M, N, K = 576 * 2, 3 * 256, 192
        A, B = mk(M, K, torch.float16, 0.3), mk(N, K, torch.float16, 0.3)
        bias = torch.randn(N, device=dev)
        tab = rope_table(576, dev)
        Cb = torch.zeros(M, N, device=dev, dtype=torch.float16)
        L.gemm(A, B, Cb, epilogue=L.EPI_QKV_ROPE, bias=bias, rope=tab, rope_period=576, rope_cols=512)
        torch.cuda.synchronize()
        ref = apply_rope_ref(A.float() @ B.float().T + bias, tab, 576, 512)
        res["err"] = rel_err(Cb, ref)
    elif name.startswith("mn_"):
        # C[M][N] += A^T-stored . B^T-stored ; A stored [K][M], B stored [K][N]; split-K atomics
        _, lbo, sbo = name.split("_")
        M, N, K = 256, 64, 64 * 6
        At = mk(K, M, torch.float16)
        Bt = mk(K, N, torch.float16)
        Cb = torch.zeros(M, 48, device=dev)
        L.gemm(At, Bt[:, :48], Cb, epilogue=L.EPI_ATOMIC_F32, a_mn=True, b_mn=True, N=48, splitk=3,
               dbg_lbo=int(lbo), dbg_sbo=int(sbo))
        torch.cuda.synchronize()
        ref = At.float().T @ Bt[:, :48].float()
        res["err"] = rel_err(Cb, ref)
        Ct = torch.zeros(48, M, device=dev)
        L.gemm(At, Bt[:, :48], Ct, epilogue=L.EPI_ATOMIC_F32, a_mn=True, b_mn=True, N=48, splitk=2, c_trans=True,
               dbg_lbo=int(lbo), dbg_sbo=int(sbo))
        torch.cuda.synchronize()
        res["err_trans"] = rel_err(Ct, ref.T)
    elif name.startswith("perf_"):
        shapes = {
            "perf_qkv": (41472, 3072, 1088, L.EPI_STORE16),
            "perf_fc1": (41472, 4736, 1088, L.EPI_STORE16),
            "perf_fc2": (41472, 1024, 4800, L.EPI_STORE16),
            "perf_proj": (41472, 1024, 1088, L.EPI_STORE16),
            "perf_skinny": (41472, 64, 1024, L.EPI_STORE16),
            "perf_skinny_fc2": (41472, 64, 4736, L.EPI_STORE16),   # A = 393 MB: larger than L2, the HBM-bound case
            "perf_skinny_b4": (20736, 64, 4736, L.EPI_STORE16),
            "perf_gelu": (41472, 4736, 1088, L.EPI_GELU),
            "perf_dgelu": (41472, 4736, 1088, L.EPI_DGELU),
        }
        M, N, K, epi = shapes[name]
        A, B = mk(M, K, torch.float16), mk(N, K, torch.float16)
        Cb = torch.zeros(M, N, device=dev, dtype=torch.float16)
        if epi in (L.EPI_GELU, L.EPI_DGELU):   # epilogue cost only: time the GELU / GELU' variants of the same shape
            G2 = torch.zeros(M, N, device=dev, dtype=torch.float16)
            Hh = (torch.randn(M, N, device=dev)).to(torch.float16)
            kw2 = dict(C2=G2) if epi == L.EPI_GELU else dict(aux=Hh)
            for _ in range(3):
                L.gemm(A, B, Cb, epilogue=epi, cta_pair=pair, **kw2)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                L.gemm(A, B, Cb, epilogue=epi, cta_pair=pair, **kw2)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            res["ms"] = ms
            res["tflops"] = 2.0 * M * N * K / ms / 1e9
            return res
        for _ in range(3):
            L.gemm(A, B, Cb, epilogue=epi, cta_pair=pair)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 10
        e0.record()
        for _ in range(iters):
            L.gemm(A, B, Cb, epilogue=epi, cta_pair=pair)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        res["ms"] = ms
        res["tflops"] = 2.0 * M * N * K / ms / 1e9
        res["gbs"] = (M * K * 2 + N * K * 2 + M * N * 2) / ms / 1e6
        ref = A[:256].float() @ B.float().T
        res["err"] = rel_err(Cb[:256], ref)
        # cuBLAS for comparison (library GEMM; context only)
        for _ in range(3):
            torch.matmul(A, B.T)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(iters):
            torch.matmul(A, B.T)
        e1.record()
        torch.cuda.synchronize()
        res["cublas_tflops"] = 2.0 * M * N * K / (e0.elapsed_time(e1) / iters) / 1e9
    else:
        raise SystemExit(f"unknown case {name}")
    return res


CASES = [
    "k_tile1", "k_tile1_bn64", "k_multi", "k_ragged", "k_store16", "k_skinny", "k_skinny_full", "k_skinny_ragged", "k_persist", "k_big_f16", "k_big_bf16",
    "epi_residual", "epi_gelu", "epi_dgelu", "epi_rope",
    "mn_8192_1024", "mn_1024_8192", "mn_16_1024", "mn_8192_128",
    "perf_qkv", "perf_fc1", "perf_fc2", "perf_proj", "perf_skinny", "perf_skinny_fc2", "perf_skinny_b4", "perf_gelu", "perf_dgelu",
    "p_tile1", "p_multi", "p_ragged", "p_store16", "p_persist", "p_big_f16", "p_big_bf16",
    "perfp_qkv", "perfp_fc1", "perfp_fc2", "perfp_proj", "perfp_gelu", "perfp_dgelu",
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case")
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    if args.case:
        try:
            r = run_case(args.case)
        except Exception as e:  # noqa: BLE001
            r = {"case": args.case, "error": f"{type(e).__name__}: {e}"[:500]}
        print("RESULT " + json.dumps(r), flush=True)
        return
    OUT.mkdir(exist_ok=True)
    log = OUT / "bringup_gemm.jsonl"
    cases = [c for c in CASES if args.only in c]
    with open(log, "a") as f:
        for c in cases:
            t0 = time.time()
            try:
                p = subprocess.run([sys.executable, __file__, "--case", c], capture_output=True, text=True, timeout=240)
                lines = [ln for ln in p.stdout.splitlines() if ln.startswith("RESULT ")]
                if lines:
                    r = json.loads(lines[-1][7:])
                else:
                    r = {"case": c, "error": "no result", "rc": p.returncode, "stdout": p.stdout[-600:], "stderr": p.stderr[-1200:]}
            except subprocess.TimeoutExpired:
                r = {"case": c, "error": "timeout"}
            r["wall_s"] = round(time.time() - t0, 1)
            f.write(json.dumps(r) + "\n")
            f.flush()
            print(json.dumps(r), flush=True)


if __name__ == "__main__":
    main()
