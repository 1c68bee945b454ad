"""Stand-alone fused LoRA-Linear op (row a1 of the hot-path table): the CUDA implementation behind
`lora_layers.LoRALinear.forward` for adapted Linears that live outside the ViT trunk engine
(reference: lora_layers.py:49-55, 87-91).

forward   y  = [x | s*(drop(x) A)] . [W | B^T]^T + b          one tcgen05 GEMM (K-extension) + one skinny GEMM
backward  dx = [dy | s*(dy B^T)] . [W^T | A]^T                 same, on the transposed frozen weight
          dA = drop(x)^T . (s*dy B^T)     dB = (s*drop(x) A)^T . dy     split-K MN-major GEMMs (fp32 atomics)

Only PyTorch plumbing happens here (buffers, autograd.Function); there is no CPU path.
"""
from __future__ import annotations

import weakref
from typing import Optional, Tuple

import torch

from . import _lib as L

_OPERAND_DTYPE = torch.float16  # 10-bit mantissa like the TF32 path of the reference (model_builder.py:46-55)


def set_operand_dtype(dt) -> None:
    global _OPERAND_DTYPE
    if dt not in (torch.float16, torch.bfloat16):
        raise L.Sam3bError("operand dtype must be torch.float16 or torch.bfloat16")
    _OPERAND_DTYPE = dt


def _require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise L.Sam3bError(f"{what} is on {t.device}: the LoRA hot path has no CPU fallback (needs a CUDA device "
                           "and sam3_lora_b200/libsam3b.so)")


def _rpad(r: int) -> int:
    return (r + 63) // 64 * 64


class _FrozenPack:
    """16-bit copies of a frozen weight with room for the adapter K-extension: W_ext [out, in+R],
    W^T_ext [in, out+R].  Cached per weight tensor (re-packed if the tensor is modified in place)."""
    _cache: dict = {}

    @classmethod
    def get(cls, W: torch.Tensor, R: int, dt) -> Tuple[torch.Tensor, torch.Tensor]:
        ident = id(W)
        key = (W._version, W.data_ptr(), tuple(W.shape), R, dt)
        hit = cls._cache.get(ident)
        if hit is not None and hit[0] == key:
            return hit[1], hit[2]
        out_f, in_f = W.shape
        w_ext = torch.zeros(out_f, in_f + R, device=W.device, dtype=dt)
        wt_ext = torch.zeros(in_f, out_f + R, device=W.device, dtype=dt)
        w_ext[:, :in_f] = W.detach()
        wt_ext[:, :out_f] = W.detach().t()
        if hit is None:
            weakref.finalize(W, cls._cache.pop, ident, None)   # drop the packed copies with the weight
        cls._cache[ident] = (key, w_ext, wt_ext)
        return w_ext, wt_ext


def _pack_adapter(A, B, in_f, out_f, r, R, w_ext, wt_ext, dt):
    down_T = torch.empty(R, in_f, device=A.device, dtype=dt)
    up_pack = torch.empty(R, out_f, device=A.device, dtype=dt)
    site = L.make_lora_site(in_f, out_f, r, R, [(0, out_f, A, B)])
    L.lora_pack(site, down_T, w_ext, up_pack, wt_ext, dt)
    return site, down_T, up_pack


class _LoRALinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, W, bias, A, B, scaling: float, dropout_p: float):
        _require_cuda(x, "LoRALinear input")
        _require_cuda(W, "LoRALinear weight")
        dt = _OPERAND_DTYPE
        has_lora = A is not None
        out_f, in_f = W.shape
        r = A.shape[1] if has_lora else 0
        if in_f % 8 or out_f % 8:
            raise L.Sam3bError(f"LoRALinear {in_f}->{out_f}: feature sizes must be multiples of 8 for the TMA path")
        R = _rpad(r) if has_lora else 0
        lead = x.shape[:-1]
        x2 = x.reshape(-1, in_f)
        M = x2.shape[0]
        w_ext, wt_ext = _FrozenPack.get(W, R, dt)
        site = down_T = up_pack = None
        if has_lora:
            Af, Bf = A.detach().float().contiguous(), B.detach().float().contiguous()
            site, down_T, up_pack = _pack_adapter(Af, Bf, in_f, out_f, r, R, w_ext, wt_ext, dt)
        x16 = torch.empty(M, in_f + R, device=x.device, dtype=dt)
        xf = x2.float().contiguous()
        L.cast_rows_16(xf, x16)
        if not has_lora:
            mask, xd16 = None, x16
        elif dropout_p > 0.0:
            # adapter branch sees inverted-dropout(x) (lora_layers.py:54); the base branch sees x
            mask = (torch.rand_like(xf) >= dropout_p).to(dt) * (1.0 / (1.0 - dropout_p))
            xd16 = torch.empty(M, in_f + R, device=x.device, dtype=dt)
            xd16[:, :in_f] = x16[:, :in_f] * mask
            L.gemm(xd16[:, :in_f], down_T, xd16[:, in_f:], epilogue=L.EPI_STORE16, alpha=scaling, bn=64)
            x16[:, in_f:] = xd16[:, in_f:]
        else:
            mask, xd16 = None, x16
            L.gemm(x16[:, :in_f], down_T, x16[:, in_f:], epilogue=L.EPI_STORE16, alpha=scaling, bn=64)
        y = torch.empty(M, out_f, device=x.device, dtype=torch.float32)
        L.gemm(x16, w_ext, y, epilogue=L.EPI_STORE32, bias=None if bias is None else bias.detach().float().contiguous())
        ctx.save_for_backward(xd16, mask if mask is not None else torch.empty(0, device=x.device), wt_ext,
                              up_pack if up_pack is not None else torch.empty(0, device=x.device))
        ctx.meta = (in_f, out_f, r, R, scaling, dropout_p, lead, site, x.dtype, A.dtype if has_lora else torch.float32)
        return y.reshape(*lead, out_f).to(x.dtype)

    @staticmethod
    def backward(ctx, gy):
        xd16, mask, wt_ext, up_pack = ctx.saved_tensors
        in_f, out_f, r, R, scaling, dropout_p, lead, site, xdtype, pdtype = ctx.meta
        dt = xd16.dtype
        g2 = gy.reshape(-1, out_f).float().contiguous()
        M = g2.shape[0]
        sc = L.grad_scale(g2)                      # the chain below runs on s*gy; outputs are multiplied by 1/s (on the device)
        s_in, s_out = sc[0:1], sc[1:2]
        dy16 = torch.empty(M, out_f + R, device=gy.device, dtype=dt)
        L.cast_rows_16(g2, dy16, s_in)
        if R > 0:
            L.gemm(dy16[:, :out_f], up_pack, dy16[:, out_f:], epilogue=L.EPI_STORE16, alpha=scaling, bn=64)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty(M, in_f, device=gy.device, dtype=torch.float32)
            if dropout_p > 0.0:
                L.gemm(dy16[:, :out_f], wt_ext[:, :out_f], dx, epilogue=L.EPI_STORE32, row_scale=s_out, rows_per_scale=L.ALL_ROWS)
                dxl = torch.empty(M, in_f, device=gy.device, dtype=torch.float32)
                L.gemm(dy16[:, out_f:], wt_ext[:, out_f:], dxl, epilogue=L.EPI_STORE32, row_scale=s_out, rows_per_scale=L.ALL_ROWS)
                dx += dxl * mask.float()
            else:
                L.gemm(dy16, wt_ext, dx, epilogue=L.EPI_STORE32, row_scale=s_out, rows_per_scale=L.ALL_ROWS)
            dx = dx.reshape(*lead, in_f).to(xdtype)
        if R == 0:
            return dx, None, None, None, None, None, None
        dA_pack = torch.zeros(in_f, R, device=gy.device, dtype=torch.float32)
        dB_pack = torch.zeros(R, out_f, device=gy.device, dtype=torch.float32)
        kb = max(1, (M + 63) // 64)
        sk = max(1, min(kb // 4, 148 // max(1, (max(in_f, out_f) + 127) // 128)))
        L.gemm(dy16[:, :out_f], xd16[:, in_f:], dB_pack, epilogue=L.EPI_ATOMIC_F32, a_mn=True, b_mn=True, M=out_f, N=R, K=M,
               splitk=sk, c_trans=True)
        L.gemm(xd16[:, :in_f], dy16[:, out_f:], dA_pack, epilogue=L.EPI_ATOMIC_F32, a_mn=True, b_mn=True, M=in_f, N=R, K=M,
               splitk=sk)
        dA = (dA_pack[:, :r] * s_out).to(pdtype)
        dB = (dB_pack[:r] * s_out).to(pdtype)
        return dx, None, None, dA, dB, None, None


def lora_linear(x, W, bias, A, B, scaling: float, dropout_p: float = 0.0):
    """y = x W^T + b + scaling * (dropout(x) A) B, fused (see module docstring).  A = B = None: frozen Linear only."""
    return _LoRALinearFn.apply(x, W, bias, A, B, float(scaling), float(dropout_p))


def lora_branch(x, A, B, scaling: float, owner=None):
    """(x @ A @ B) * scaling alone (LoRALayer.forward): the same fused kernel with a zero frozen weight.
    (LoRALinear never takes this route; it exists so a bare LoRALayer behaves like the reference's.)
    The zero weight — and with it the packed operand buffers whose adapter columns each forward rewrites — belongs to
    `owner` (the calling LoRALayer), never to another layer of the same shape: a second layer's forward must not touch
    what the first layer's backward still has to read."""
    holder = owner if owner is not None else lora_branch
    cache = holder.__dict__.setdefault("_sam3b_zero_w", {})
    key = (A.shape[0], B.shape[1], str(x.device))
    if key not in cache:
        cache[key] = torch.zeros(B.shape[1], A.shape[0], device=x.device, dtype=torch.float32)
    return lora_linear(x, cache[key], None, A, B, scaling, 0.0)
