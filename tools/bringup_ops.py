"""GPU bring-up for attention fwd/bwd and the elementwise kernels (run under gpurun).

One subprocess per case; JSON lines appended to gpurun_out/bringup_ops.jsonl.
"""
from __future__ import annotations

import argparse
import json
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
OUT = ROOT / "gpurun_out"


def rel_err(got, ref):
    import torch

    got, ref = got.float(), ref.float()
    if not torch.isfinite(got).all():
        return float("inf")
    return ((got - ref).abs().max() / ref.abs().max().clamp_min(1e-20)).item()


def rope_table(period, dev):
    import torch

    ang = torch.rand(period, 32, device=dev) * 6.28
    return torch.stack([ang.cos(), ang.sin()], dim=-1).contiguous()


def rope_apply(x, table, period):
    """x: [T, H, 64] fp32 -> rotated (adjacent pairs)."""
    import torch

    T = x.shape[0]
    cs = table[torch.arange(T, device=x.device) % period]  # [T,32,2]
    xr = x.reshape(T, x.shape[1], 32, 2)
    a, b = xr[..., 0], xr[..., 1]
    c, s = cs[:, None, :, 0], cs[:, None, :, 1]
    return torch.stack([a * c - b * s, a * s + b * c], dim=-1).reshape(x.shape)


def attn_ref(q, k, v, L):
    """q,k,v: [T,H,64] fp32 (already rotated); segments of L rows.  Returns O [T,H,64], lse2 [T,H]."""
    import torch

    T, H, _ = q.shape
    S = T // L
    qs, ks, vs = (t.reshape(S, L, H, 64).permute(0, 2, 1, 3) for t in (q, k, v))
    sc = (qs @ ks.transpose(-1, -2)) * 0.125
    lse = torch.logsumexp(sc, dim=-1)  # [S,H,L]
    O = torch.softmax(sc, dim=-1) @ vs
    return O.permute(0, 2, 1, 3).reshape(T, H, 64), (lse * 1.4426950408889634).permute(0, 2, 1).reshape(T, H)


def run_case(name: str) -> dict:
    import torch

    from sam3_lora_b200 import _lib as L

    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    dev = "cuda"
    res = {"case": name}

    if name.startswith("attn_"):
        # attn_<fwd|bwd|perf>_<L>_<segs>_<heads>[_bf16]
        parts = name.split("_")
        mode, Ls, segs, heads = parts[1], int(parts[2]), int(parts[3]), int(parts[4])
        dt = torch.bfloat16 if parts[-1] == "bf16" else torch.float16
        D = heads * 64
        T = Ls * segs
        period = Ls
        tab = rope_table(period, dev)
        raw = torch.randn(T, 3, heads, 64, device=dev) * 1.0
        raw16 = raw.to(dt)  # the qkv projection output as the GEMM epilogue would see it (before rope) is fp32;
        # here we start from 16-bit-rounded raw values so both paths see identical inputs.
        rawf = raw16.float().requires_grad_(True)
        q_rot = rope_apply(rawf[:, 0], tab, period)
        k_rot = rope_apply(rawf[:, 1], tab, period)
        v = rawf[:, 2]
        # kernel input: rotated q,k rounded to 16-bit
        qkv = torch.empty(T, 3 * D + 64, device=dev, dtype=dt)
        qkv[:, :D] = q_rot.detach().reshape(T, D).to(dt)
        qkv[:, D:2 * D] = k_rot.detach().reshape(T, D).to(dt)
        qkv[:, 2 * D:3 * D] = v.detach().reshape(T, D).to(dt)
        O = torch.zeros(T, D + 64, device=dev, dtype=dt)
        lse2 = torch.zeros(heads, T, device=dev)   # head-major
        L.attention_fwd(qkv, Ls, D, heads, O, lse2)
        torch.cuda.synchronize()
        qf = qkv[:, :D].float().reshape(T, heads, 64)
        kf = qkv[:, D:2 * D].float().reshape(T, heads, 64)
        vf = qkv[:, 2 * D:3 * D].float().reshape(T, heads, 64)
        O_ref, lse_ref = attn_ref(qf, kf, vf, Ls)
        res["err_O"] = rel_err(O[:, :D].reshape(T, heads, 64), O_ref)
        res["err_lse"] = rel_err(lse2, lse_ref.T)
        res["pad_untouched"] = bool((O[:, D:] == 0).all().item())
        if mode in ("bwd", "perf"):
            dO = (torch.randn(T, D, device=dev) * 0.5).to(dt)
            delta = torch.zeros(heads, T, device=dev)
            dqkv = torch.zeros(T, 3 * D + 64, device=dev, dtype=dt)
            L.attention_bwd(qkv, Ls, D, heads, O, lse2, dO, delta, dqkv, tab, period)
            torch.cuda.synchronize()
            # reference: autograd from the raw (un-rotated) q,k through rope + attention, fp32.
            # The kernel sees 16-bit-rounded rotated q,k; do the same in the reference via a straight-through round.
            def ste(x):
                return x + (x.detach().to(dt).float() - x.detach())
            O_r, _ = attn_ref(ste(q_rot), ste(k_rot), ste(v), Ls)
            (O_r.reshape(T, D) * dO.float()).sum().backward()
            g = rawf.grad  # [T,3,H,64]
            res["err_dq"] = rel_err(dqkv[:, :D].reshape(T, heads, 64), g[:, 0])
            res["err_dk"] = rel_err(dqkv[:, D:2 * D].reshape(T, heads, 64), g[:, 1])
            res["err_dv"] = rel_err(dqkv[:, 2 * D:3 * D].reshape(T, heads, 64), g[:, 2])
            delta_ref = (dO.float().reshape(T, heads, 64) * O[:, :D].float().reshape(T, heads, 64)).sum(-1)
            res["err_delta"] = rel_err(delta, delta_ref.T)
        if mode == "perf":
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for _ in range(3):
                L.attention_fwd(qkv, Ls, D, heads, O, lse2)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(10):
                L.attention_fwd(qkv, Ls, D, heads, O, lse2)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            flops = 4.0 * segs * heads * Ls * Ls * 64
            res["fwd_ms"] = ms
            res["fwd_tflops"] = flops / ms / 1e9
            for _ in range(3):
                L.attention_bwd(qkv, Ls, D, heads, O, lse2, dO, delta, dqkv, tab, period)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(10):
                L.attention_bwd(qkv, Ls, D, heads, O, lse2, dO, delta, dqkv, tab, period)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            res["bwd_ms"] = ms
            res["bwd_tflops_algo"] = 2.5 * flops / ms / 1e9  # 5 matmuls algorithmic (7 executed)
            # library SDPA for context
            qs = qf.reshape(segs, Ls, heads, 64).permute(0, 2, 1, 3).to(dt).contiguous()
            ks_ = kf.reshape(segs, Ls, heads, 64).permute(0, 2, 1, 3).to(dt).contiguous()
            vs_ = vf.reshape(segs, Ls, heads, 64).permute(0, 2, 1, 3).to(dt).contiguous()
            for _ in range(3):
                torch.nn.functional.scaled_dot_product_attention(qs, ks_, vs_)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(10):
                torch.nn.functional.scaled_dot_product_attention(qs, ks_, vs_)
            e1.record()
            torch.cuda.synchronize()
            res["torch_sdpa_fwd_tflops"] = flops / (e0.elapsed_time(e1) / 10) / 1e9
    elif name.startswith("ln_"):
        D = int(name.split("_")[1])
        rows = 1000
        x = torch.randn(rows, D, device=dev) * 2 + 0.5
        gamma, beta = torch.randn(D, device=dev), torch.randn(D, device=dev)
        y = torch.zeros(rows, D + 64, device=dev, dtype=torch.float16)
        mean, rstd = torch.zeros(rows, device=dev), torch.zeros(rows, device=dev)
        L.layernorm_fwd(x, gamma, beta, 1e-5, y, mean, rstd)
        torch.cuda.synchronize()
        xr = x.clone().requires_grad_(True)
        yr = torch.nn.functional.layer_norm(xr, (D,), gamma, beta, 1e-5)
        res["err_fwd"] = rel_err(y[:, :D], yr)
        res["err_mean"] = rel_err(mean, x.mean(-1))
        dy = (torch.randn(rows, D, device=dev)).to(torch.float16)
        dres = torch.randn(rows, D, device=dev)
        dx = torch.zeros(rows, D, device=dev)
        dx16 = torch.zeros(rows, D + 64, device=dev, dtype=torch.float16)
        L.layernorm_bwd(dy, x, mean, rstd, gamma, dres, dx, dx16)
        torch.cuda.synchronize()
        yr.backward(dy.float())
        ref = xr.grad + dres
        res["err_bwd"] = rel_err(dx, ref)
        res["err_bwd16"] = rel_err(dx16[:, :D], ref)
        c16 = torch.zeros(rows, D + 64, device=dev, dtype=torch.bfloat16)
        L.cast_rows_16(x, c16)
        torch.cuda.synchronize()
        res["err_cast"] = rel_err(c16[:, :D], x.to(torch.bfloat16))
    elif name == "patch":
        B, P, ws, Gd = 2, 14, 8, 24
        img = torch.randn(B, 3, P * Gd, P * Gd, device=dev)
        Kp = 640
        out = torch.zeros(B * Gd * Gd, Kp, device=dev, dtype=torch.float16)
        L.patch_gather(img, P, ws, out, Kp)
        torch.cuda.synchronize()
        un = torch.nn.functional.unfold(img, kernel_size=P, stride=P)  # [B, 588, G*G] row-major positions
        un = un.transpose(1, 2).reshape(B, Gd // ws, ws, Gd // ws, ws, 588).permute(0, 1, 3, 2, 4, 5).reshape(B * Gd * Gd, 588)
        res["err"] = rel_err(out[:, :588], un.to(torch.float16))
        res["pad_zero"] = bool((out[:, 588:] == 0).all().item())
        # layout transposes
        D = 128
        x = torch.randn(B, Gd * Gd, D, device=dev)  # window-major tokens
        nchw = torch.zeros(B, D, Gd, Gd, device=dev)
        L.tokens_to_nchw(x, B, Gd, ws, D, nchw)
        torch.cuda.synchronize()
        ref = x.reshape(B, Gd // ws, Gd // ws, ws, ws, D).permute(0, 1, 3, 2, 4, 5).reshape(B, Gd, Gd, D).permute(0, 3, 1, 2)
        res["err_to_nchw"] = rel_err(nchw, ref)
        back = torch.zeros(B * Gd * Gd, D, device=dev)
        back16 = torch.zeros(B * Gd * Gd, D + 64, device=dev, dtype=torch.float16)
        L.nchw_to_tokens(nchw, B, Gd, ws, D, back, back16)
        torch.cuda.synchronize()
        res["err_roundtrip"] = rel_err(back, x.reshape(-1, D))
        res["err_roundtrip16"] = rel_err(back16[:, :D], x.reshape(-1, D))
    elif name == "lora_pack":
        inf, r, rpad = 256, 16, 64
        outs = [(0, 128), (128, 128), (256, 128)]
        A = [torch.randn(inf, r, device=dev) for _ in outs]
        Bm = [torch.randn(r, ln, device=dev) for _, ln in outs]
        site = L.make_lora_site(inf, 384, r, rpad, [(o, ln, A[i], Bm[i]) for i, (o, ln) in enumerate(outs)])
        down_T = torch.zeros(rpad, inf, device=dev, dtype=torch.float16)
        w_ext = torch.full((384, inf + rpad), 7.0, device=dev, dtype=torch.float16)
        up = torch.zeros(rpad, 384, device=dev, dtype=torch.float16)
        wt_ext = torch.full((inf, 384 + rpad), 7.0, device=dev, dtype=torch.float16)
        L.lora_pack(site, down_T, w_ext, up, wt_ext, torch.float16)
        torch.cuda.synchronize()
        ok = True
        for i, (o, ln) in enumerate(outs):
            ok &= torch.equal(down_T[i * r:(i + 1) * r], A[i].T.to(torch.float16))
            ok &= torch.equal(w_ext[o:o + ln, inf + i * r: inf + (i + 1) * r], Bm[i].T.to(torch.float16))
            ok &= torch.equal(up[i * r:(i + 1) * r, o:o + ln], Bm[i].to(torch.float16))
            ok &= torch.equal(wt_ext[:, 384 + i * r: 384 + (i + 1) * r], A[i].to(torch.float16))
        ok &= bool((down_T[48:] == 0).all()) and bool((w_ext[:, :inf] == 7).all()) and bool((wt_ext[:, :384] == 7).all())
        ok &= bool((up[0:16, 128:] == 0).all()) and bool((w_ext[128:, inf:inf + 16] == 0).all())
        res["ok"] = bool(ok)
        dA_pack = torch.randn(inf, rpad, device=dev)
        dB_pack = torch.randn(rpad, 384, device=dev)
        dA = [torch.zeros(inf, r, device=dev) for _ in outs]
        dB = [torch.zeros(r, ln, device=dev) for _, ln in outs]
        L.lora_unpack_grads(site, dA_pack, dB_pack, dA, dB)
        torch.cuda.synchronize()
        ok2 = True
        for i, (o, ln) in enumerate(outs):
            ok2 &= torch.equal(dA[i], dA_pack[:, i * r:(i + 1) * r])
            ok2 &= torch.equal(dB[i], dB_pack[i * r:(i + 1) * r, o:o + ln])
        res["ok_unpack"] = bool(ok2)
    elif name == "adamw":
        n = 100003
        p0 = torch.randn(n, device=dev)
        p = p0.clone()
        m, v = torch.zeros(n, device=dev), torch.zeros(n, device=dev)
        pt = torch.nn.Parameter(p0.clone())
        opt = torch.optim.AdamW([pt], lr=5e-3, weight_decay=0.01)
        for step in range(1, 4):
            g = torch.randn(n, device=dev)
            L.adamw_step(p, g, m, v, 5e-3, 0.9, 0.999, 1e-8, 0.01, step)
            pt.grad = g.clone()
            opt.step()
        torch.cuda.synchronize()
        res["err"] = rel_err(p, pt.detach())
    else:
        raise SystemExit(f"unknown case {name}")
    return res


CASES = [
    "ln_128", "ln_1024", "patch", "lora_pack", "adamw",
    "attn_fwd_576_2_2", "attn_fwd_64_3_2", "attn_fwd_192_1_1", "attn_fwd_1152_1_2",
    "attn_bwd_576_2_2", "attn_bwd_64_3_2", "attn_bwd_576_2_2_bf16",
    "attn_bwd_5184_1_16",
    "attn_perf_576_72_16", "attn_perf_5184_8_16",
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
            r = {"case": args.case, "error": f"{type(e).__name__}: {e}"[:600]}
        print("RESULT " + json.dumps(r), flush=True)
        return
    OUT.mkdir(exist_ok=True)
    log = OUT / "bringup_ops.jsonl"
    cases = [c for c in CASES if args.only in c]
    with open(log, "a") as f:
        for c in cases:
            t0 = time.time()
            try:
                p = subprocess.run([sys.executable, __file__, "--case", c], capture_output=True, text=True, timeout=300)
                lines = [ln for ln in p.stdout.splitlines() if ln.startswith("RESULT ")]
                if lines:
                    r = json.loads(lines[-1][7:])
                    if p.returncode != 0 or "error" in r:
                        r["stderr"] = p.stderr[-800:]
                        r["stdout"] = p.stdout[-800:]
                else:
                    r = {"case": c, "error": "no result", "rc": p.returncode, "stdout": p.stdout[-1500:], "stderr": p.stderr[-1500:]}
            except subprocess.TimeoutExpired:
                r = {"case": c, "error": "timeout"}
            r["wall_s"] = round(time.time() - t0, 1)
            f.write(json.dumps(r) + "\n")
            f.flush()
            print(json.dumps(r), flush=True)


if __name__ == "__main__":
    main()
