"""Drop-in for the `nn.MultiheadAttention` sites of the SAM3 detector (row a7 of the hot-path table).

Reference call sites (all through `MultiheadAttentionWrapper`, sam3/model/model_misc.py:31-34, i.e.
`need_weights=False`):
    DETR encoder   self_attn / cross_attn_image   sam3/model/encoder.py:139-201   (E=256, 8 heads, dropout 0.1)
    DETR decoder   self_attn / ca_text / cross_attn with additive box-RPB attn_mask   sam3/model/decoder.py:80-187
    seg head       cross_attend_prompt            sam3/model/maskformer_segmentation.py:281-289

Same constructor subset, parameter names (`in_proj_weight [3E,E]`, `in_proj_bias`, `out_proj.{weight,bias}`)
and call signature as torch's module, so checkpoints and call sites are unchanged.  LoRA: the reference's
q_proj / k_proj / v_proj / out_proj target names address the three row-slices of `in_proj_weight` and `out_proj`
as virtual children — `out_proj` is a real adapter here, not the inert one of the reference.

Arithmetic: the three projections and the output projection are fused LoRA GEMMs (ops.lora_linear);
the attention core is the tcgen05 flash kernel of csrc/attn_fwd.cu / attn_bwd.cu in its GEN instantiation
(additive float attn_mask, boolean key_padding_mask, dropout on the probabilities).  head_dim 32 runs as
zero-padded 64-wide heads: the frozen weights are expanded once so the GEMMs emit / consume the padded layout.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch
import torch.nn as nn

from . import _lib as L
from .lora_layers import LoRALinear, LoRAVirtual
from .ops import lora_linear, _OPERAND_DTYPE  # noqa: F401


def _pad_index(E: int, H: int, device) -> torch.Tensor:
    """positions of the E real features inside the padded [H*64] layout (head h at 64*h .. 64*h + E/H)."""
    hd = E // H
    return (torch.arange(H, device=device).view(H, 1) * 64 + torch.arange(hd, device=device).view(1, hd)).reshape(-1)


class _AttnCoreFn(torch.autograd.Function):
    """softmax(q k^T * scale + attn_mask + key_padding_mask) -> dropout -> @ v on padded 64-wide heads."""

    @staticmethod
    def forward(ctx, q, k, v, nseg: int, Lq: int, Lk: int, heads: int, scale: float, bias, kpm, drop_p: float, seed: int):
        from .ops import _OPERAND_DTYPE as dt  # noqa: PLC0415

        Ep = heads * 64
        dev = q.device
        q16 = torch.empty(nseg * Lq, Ep, device=dev, dtype=dt)
        kv16 = torch.empty(nseg * Lk, 2 * Ep, device=dev, dtype=dt)
        L.cast_rows_16(q.float().contiguous(), q16)
        L.cast_rows_16(k.float().contiguous(), kv16[:, :Ep])
        L.cast_rows_16(v.float().contiguous(), kv16[:, Ep:])
        O16 = torch.empty(nseg * Lq, Ep, device=dev, dtype=dt)
        Ls = (Lq + 63) // 64 * 64
        lse2 = torch.zeros(heads, nseg * Ls, device=dev, dtype=torch.float32)
        d = L.mha_desc(q16, kv16, nseg, Lq, Lk, heads, scale, O16, lse2, bias=bias, kpm=kpm, drop_p=drop_p, drop_seed=seed)
        ctx.drop_bits = L.attention_dropout_bits(d)       # large problems: the mask as keep-bits, hashed once (None otherwise)
        L.mha_fwd(d)
        ctx.save_for_backward(q16, kv16, O16, lse2, bias if bias is not None else torch.empty(0, device=dev),
                              kpm if kpm is not None else torch.empty(0, device=dev, dtype=torch.uint8))
        ctx.meta = (nseg, Lq, Lk, heads, scale, bias is not None, kpm is not None, drop_p, seed, q.dtype)
        return O16.float()

    @staticmethod
    def backward(ctx, gO):
        q16, kv16, O16, lse2, bias, kpm = ctx.saved_tensors
        nseg, Lq, Lk, heads, scale, has_bias, has_kpm, drop_p, seed, qdt = ctx.meta
        Ep = heads * 64
        dev = gO.device
        dO16 = torch.empty(nseg * Lq, Ep, device=dev, dtype=q16.dtype)
        g32 = gO.float().contiguous()
        # backward on s*gO (linear in gO), outputs times 1/s; see _lib.grad_scale.  dV[k] = sum_q P[q,k] dO[q] and dK are sums
        # over all Lq queries and are stored in 16 bits: with max|s*gO| = T the sums are bounded by Lq*T, so T = 32768/Lq keeps
        # them below fp16's 65504 whatever the attention pattern (5184 image tokens attending 33 prompt tokens overflowed at 256)
        sc = L.grad_scale(g32, target=min(L.GRAD_SCALE_TARGET, max(1.0, 32768.0 / Lq)))
        L.cast_rows_16(g32, dO16, sc[0:1])
        delta = torch.zeros_like(lse2)
        dq16 = torch.empty(nseg * Lq, Ep, device=dev, dtype=q16.dtype)
        dkv16 = torch.empty(nseg * Lk, 2 * Ep, device=dev, dtype=q16.dtype)
        d = L.mha_desc(q16, kv16, nseg, Lq, Lk, heads, scale, O16, lse2, bias=bias if has_bias else None,
                       kpm=kpm if has_kpm else None, drop_p=drop_p, drop_seed=seed)
        if ctx.drop_bits is not None:
            d.drop_bits, d.drop_bitsT = L.ptr(ctx.drop_bits[0]), L.ptr(ctx.drop_bits[1])
        L.mha_bwd(d, dO16, delta, dq16, dkv16)
        inv = sc[1:2]
        return ((dq16.float() * inv).to(qdt), (dkv16[:, :Ep].float() * inv).to(qdt), (dkv16[:, Ep:].float() * inv).to(qdt),
                None, None, None, None, None, None, None, None, None)


class _FusedMHAFn(torch.autograd.Function):
    """The whole adapter-free module in 16-bit between the kernels: input cast(s) -> projection GEMMs that write the padded-
    head q | k | v layout the attention kernel reads (q and k share one GEMM when they share their input: the encoder's
    q = k = x + pos; k and v likewise in cross attention) -> flash attention -> output projection (fp32 out).  Backward:
    out-proj dgrad -> attention backward -> projection dgrads (one GEMM for a shared input).  No fp32 round trip of q, k, v,
    O or their gradients; frozen weights need no weight gradients."""

    @staticmethod
    def forward(ctx, q_in, k_in, v_in, same_qk: bool, same_kv: bool, packs, nseg: int, Lq: int, Lk: int, heads: int, scale: float,
                bias, kpm, drop_p: float, seed: int):
        from .ops import _OPERAND_DTYPE as dt  # noqa: PLC0415

        Ep = heads * 64
        dev = q_in.device
        E = q_in.shape[1]
        Mq, Mk = nseg * Lq, nseg * Lk
        W = packs[dt]

        def cast(x):
            y = torch.empty(x.shape[0], E, device=dev, dtype=dt)
            L.cast_rows_16(x.float().contiguous(), y)
            return y

        xq = cast(q_in)
        xk = xq if same_qk else cast(k_in)
        xv = xk if same_kv else (xq if (v_in is q_in) else cast(v_in))
        self_attn = same_qk and Mq == Mk
        if self_attn:
            buf = torch.empty(Mq, 3 * Ep, device=dev, dtype=dt)            # q | k | v
            if same_kv:                                                     # q = k = v = x (the text tower's resblocks)
                L.gemm(xq, W["Wqkv"], buf, epilogue=L.EPI_STORE16, bias=W["bqkv"])
            else:
                L.gemm(xq, W["Wqk"], buf[:, :2 * Ep], epilogue=L.EPI_STORE16, bias=W["bqk"])
                L.gemm(xv, W["Wv"], buf[:, 2 * Ep:], epilogue=L.EPI_STORE16, bias=W["bv"])
            q16, kv16, qc, kc, vc = buf, buf, 0, Ep, 2 * Ep
        else:
            q16 = torch.empty(Mq, Ep, device=dev, dtype=dt)
            kv16 = torch.empty(Mk, 2 * Ep, device=dev, dtype=dt)
            L.gemm(xq, W["Wq"], q16, epilogue=L.EPI_STORE16, bias=W["bq"])
            if same_kv:
                L.gemm(xk, W["Wkv"], kv16, epilogue=L.EPI_STORE16, bias=W["bkv"])
            else:
                L.gemm(xk, W["Wk"], kv16[:, :Ep], epilogue=L.EPI_STORE16, bias=W["bk"])
                L.gemm(xv, W["Wv"], kv16[:, Ep:], epilogue=L.EPI_STORE16, bias=W["bv"])
            qc, kc, vc = 0, 0, Ep
        O16 = torch.empty(Mq, Ep, device=dev, dtype=dt)
        Ls = (Lq + 63) // 64 * 64
        lse2 = torch.zeros(heads, nseg * Ls, device=dev, dtype=torch.float32)
        d = L.mha_desc(q16, kv16, nseg, Lq, Lk, heads, scale, O16, lse2, q_col0=qc, k_col0=kc, v_col0=vc, bias=bias, kpm=kpm,
                       drop_p=drop_p, drop_seed=seed)
        ctx.drop_bits = L.attention_dropout_bits(d)
        L.mha_fwd(d)
        out = torch.empty(Mq, E, device=dev, dtype=torch.float32)
        L.gemm(O16, W["Wo"], out, epilogue=L.EPI_STORE32, bias=W["bo"])
        ctx.save_for_backward(q16, kv16, O16, lse2, bias if bias is not None else torch.empty(0, device=dev),
                              kpm if kpm is not None else torch.empty(0, device=dev, dtype=torch.uint8))
        ctx.meta = (nseg, Lq, Lk, heads, scale, bias is not None, kpm is not None, drop_p, seed, q_in.dtype, same_qk, same_kv,
                    self_attn, (qc, kc, vc), E, v_in is q_in)
        ctx.W = W
        return out.to(q_in.dtype)

    @staticmethod
    def backward(ctx, g):
        q16, kv16, O16, lse2, bias, kpm = ctx.saved_tensors
        (nseg, Lq, Lk, heads, scale, has_bias, has_kpm, drop_p, seed, xdt, same_qk, same_kv, self_attn, (qc, kc, vc), E,
         v_is_q) = ctx.meta
        W = ctx.W
        Ep = heads * 64
        dev = g.device
        dt = q16.dtype
        Mq, Mk = nseg * Lq, nseg * Lk
        g32 = g.reshape(Mq, E).float().contiguous()
        sc = L.grad_scale(g32, target=min(L.GRAD_SCALE_TARGET, max(1.0, 32768.0 / Lq)))      # see _AttnCoreFn.backward
        dy16 = torch.empty(Mq, E, device=dev, dtype=dt)
        L.cast_rows_16(g32, dy16, sc[0:1])
        dO16 = torch.empty(Mq, Ep, device=dev, dtype=dt)
        L.gemm(dy16, W["WoT"], dO16, epilogue=L.EPI_STORE16)
        delta = torch.zeros_like(lse2)
        d = L.mha_desc(q16, kv16, nseg, Lq, Lk, heads, scale, O16, lse2, q_col0=qc, k_col0=kc, v_col0=vc,
                       bias=bias if has_bias else None, kpm=kpm if has_kpm else None, drop_p=drop_p, drop_seed=seed)
        if ctx.drop_bits is not None:
            d.drop_bits, d.drop_bitsT = L.ptr(ctx.drop_bits[0]), L.ptr(ctx.drop_bits[1])
        inv = sc[1:2]
        need_q, need_k, need_v = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[2]

        def dgrad(dy, Wt, rows):
            dx = torch.empty(rows, E, device=dev, dtype=torch.float32)
            L.gemm(dy, Wt, dx, epilogue=L.EPI_STORE32, row_scale=inv, rows_per_scale=L.ALL_ROWS)
            return dx.to(xdt)

        if self_attn:
            dbuf = torch.empty(Mq, 3 * Ep, device=dev, dtype=dt)          # dq | dk | dv
            L.mha_bwd(d, dO16, delta, dbuf, dbuf, dk_col0=Ep, dv_col0=2 * Ep)
            if same_kv:
                return ((dgrad(dbuf, W["WqkvT"], Mq) if need_q else None), None, None) + (None,) * 12
            gq = dgrad(dbuf[:, :2 * Ep], W["WqkT"], Mq) if (need_q or need_k) else None      # [dq | dk] . [Wq ; Wk]
            gv = dgrad(dbuf[:, 2 * Ep:], W["WvT"], Mk) if (need_v or (v_is_q and need_q)) else None
            if v_is_q and gv is not None:
                gq = gv if gq is None else gq + gv
                gv = None
            return (gq, None, gv) + (None,) * 12
        dq16 = torch.empty(Mq, Ep, device=dev, dtype=dt)
        dkv16 = torch.empty(Mk, 2 * Ep, device=dev, dtype=dt)
        L.mha_bwd(d, dO16, delta, dq16, dkv16)
        gq = dgrad(dq16, W["WqT"], Mq) if need_q else None
        gk = gv = None
        if same_kv:
            if need_k or need_v:
                gk = dgrad(dkv16, W["WkvT"], Mk)                          # [dk | dv] . [Wk ; Wv]
        else:
            gk = dgrad(dkv16[:, :Ep], W["WkT"], Mk) if need_k else None
            gv = dgrad(dkv16[:, Ep:], W["WvT"], Mk) if need_v else None
        if same_qk and gk is not None:                                    # cross-shaped call with q is k (Mq == Mk handled above)
            gq = gk if gq is None else gq + gk
            gk = None
        return (gq, gk, gv) + (None,) * 12


class MultiheadAttention(nn.Module):
    lora_out_proj_ok = True   # apply_lora_to_model may wrap out_proj in a LoRALinear (see lora_layers.py)

    def __init__(self, embed_dim: int, num_heads: int, dropout: float = 0.0, bias: bool = True, batch_first: bool = False):
        super().__init__()
        if embed_dim % num_heads:
            raise ValueError("embed_dim must be divisible by num_heads")
        hd = embed_dim // num_heads
        if hd not in (32, 64):
            raise L.Sam3bError(f"head_dim {hd} not supported by the fused kernel (32 or 64)")
        self.embed_dim, self.num_heads, self.head_dim = embed_dim, num_heads, hd
        self.dropout = dropout
        self.batch_first = batch_first
        self.in_proj_weight = nn.Parameter(torch.empty(3 * embed_dim, embed_dim))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * embed_dim)) if bias else None
        self.out_proj = nn.Linear(embed_dim, embed_dim, bias=bias)   # NonDynamicallyQuantizableLinear in torch: same params
        nn.init.xavier_uniform_(self.in_proj_weight)
        if bias:
            nn.init.zeros_(self.out_proj.bias)
        self.dropout_seed_override: Optional[int] = None
        self._packed = None
        self._packed16: Dict = {}
        self.fused = True          # adapter-free calls take the 16-bit fused path (_FusedMHAFn); False = the composed path
        self._seed_state: Optional[int] = None

    def _next_seed(self) -> int:
        """Per-call dropout seed without touching the device: a host-side Weyl sequence started from torch's CPU generator
        (so torch.manual_seed reproduces it) — the kernels hash (seed, row, col) themselves (csrc/rng.cuh)."""
        if self._seed_state is None:
            self._seed_state = int(torch.randint(0, 2 ** 31 - 1, (1,), device="cpu").item())
        self._seed_state = (self._seed_state + 0x9E3779B1) & 0xFFFFFFFF
        return self._seed_state & 0x7FFFFFFF

    # ---- LoRA -------------------------------------------------------------------------------
    def lora_virtual_targets(self) -> Dict[str, Tuple[int, int]]:
        e = self.embed_dim
        return {"q_proj": (e, e), "k_proj": (e, e), "v_proj": (e, e)}   # out_proj is a real Linear: wrapped, not virtual

    def _adapter(self, name: str):
        m = getattr(self, name, None)
        return m.lora if isinstance(m, (LoRAVirtual, LoRALinear)) else None

    @property
    def _out_linear(self) -> nn.Linear:
        return self.out_proj.original_layer if isinstance(self.out_proj, LoRALinear) else self.out_proj

    # ---- frozen weights in the padded-head layout (built once per device / weight version) ------
    def _pack(self):
        W, b = self.in_proj_weight, self.in_proj_bias
        ow = self._out_linear.weight
        key = (W._version, W.data_ptr(), ow._version, ow.data_ptr(), str(W.device))
        if self._packed is not None and self._packed[0] == key:
            return self._packed[1]
        E, H = self.embed_dim, self.num_heads
        Ep = H * 64
        idx = _pad_index(E, H, W.device)
        out = {"idx": idx}
        for i, n in enumerate(("q", "k", "v")):
            Wp = torch.zeros(Ep, E, device=W.device, dtype=torch.float32)
            Wp[idx] = W.detach()[i * E:(i + 1) * E].float()
            bp = torch.zeros(Ep, device=W.device, dtype=torch.float32)
            if b is not None:
                bp[idx] = b.detach()[i * E:(i + 1) * E].float()
            out["W" + n], out["b" + n] = Wp, bp
        Wo = torch.zeros(E, Ep, device=W.device, dtype=torch.float32)
        Wo[:, idx] = ow.detach().float()
        out["Wo"] = Wo
        out["bo"] = None if self._out_linear.bias is None else self._out_linear.bias.detach().float().contiguous()
        self._packed = (key, out)
        self._packed16 = {}
        return out

    def _pack16(self, dt):
        """16-bit operand copies of the frozen projections for the fused path: forward [N, K] and transposed (dgrad) forms,
        single and concatenated (q|k, k|v share one GEMM when they share their input).  Built once per weight version."""
        pk = self._pack()
        hit = self._packed16.get(dt)
        if hit is not None:
            return {dt: hit}
        c = lambda t: t.to(dt).contiguous()                                                   # noqa: E731
        Wq, Wk, Wv, Wo = pk["Wq"], pk["Wk"], pk["Wv"], pk["Wo"]
        w = {"Wq": c(Wq), "Wk": c(Wk), "Wv": c(Wv), "Wqk": c(torch.cat([Wq, Wk], 0)), "Wkv": c(torch.cat([Wk, Wv], 0)), "Wo": c(Wo),
             "WqT": c(Wq.t()), "WkT": c(Wk.t()), "WvT": c(Wv.t()), "WqkT": c(torch.cat([Wq, Wk], 0).t()),
             "WkvT": c(torch.cat([Wk, Wv], 0).t()), "WoT": c(Wo.t()),
             "Wqkv": c(torch.cat([Wq, Wk, Wv], 0)), "WqkvT": c(torch.cat([Wq, Wk, Wv], 0).t()),
             "bqkv": torch.cat([pk["bq"], pk["bk"], pk["bv"]]).contiguous(),
             "bq": pk["bq"], "bk": pk["bk"], "bv": pk["bv"], "bqk": torch.cat([pk["bq"], pk["bk"]]).contiguous(),
             "bkv": torch.cat([pk["bk"], pk["bv"]]).contiguous(), "bo": pk["bo"]}
        self._packed16[dt] = w
        return {dt: w}

    @staticmethod
    def _masks(attn_mask, key_padding_mask, B, H, Lq, Lk, device):
        """torch's attn_mask (bool = masked, or additive float; [Lq, Lk] or [B*H, Lq, Lk]) and key_padding_mask -> the kernel's
        additive fp32 bias [B*H, Lq, Lk] and uint8 key mask [B, Lk]."""
        bias = None
        if attn_mask is not None:
            if attn_mask.dtype == torch.bool:
                bias = torch.zeros(attn_mask.shape, device=device, dtype=torch.float32).masked_fill_(attn_mask, float("-inf"))
            else:
                bias = attn_mask.float()
            if bias.dim() == 2:
                bias = bias.expand(B * H, Lq, Lk)
            bias = bias.contiguous()
        kpm = None
        if key_padding_mask is not None:
            kpm = key_padding_mask.to(torch.uint8).contiguous() if key_padding_mask.dtype == torch.bool \
                else (key_padding_mask != 0).to(torch.uint8).contiguous()
        return bias, kpm

    def forward(self, query, key, value, key_padding_mask=None, need_weights: bool = False, attn_mask=None, **_unused):
        if need_weights:
            raise L.Sam3bError("need_weights=True is not supported by the fused attention (the reference always passes False)")
        if not query.is_cuda:
            raise L.Sam3bError("MultiheadAttention: inputs are on the CPU; the fused path has no CPU fallback")
        E, H, hd = self.embed_dim, self.num_heads, self.head_dim
        Ep = H * 64
        if self.batch_first:
            B, Lq, _ = query.shape
            Lk = key.shape[1]
            q_in, k_in, v_in = query, key, value
        else:
            Lq, B, _ = query.shape
            Lk = key.shape[0]
            q_in, k_in, v_in = query.transpose(0, 1), key.transpose(0, 1), value.transpose(0, 1)
        same_qk, same_kv, v_is_q = key is query, value is key, value is query
        q_in = q_in.contiguous().reshape(-1, E)
        k_in = q_in if same_qk else k_in.contiguous().reshape(-1, E)
        v_in = k_in if same_kv else (q_in if v_is_q else v_in.contiguous().reshape(-1, E))
        pk = self._pack()
        idx = pk["idx"]
        if self.fused and all(self._adapter(n) is None for n in ("q_proj", "k_proj", "v_proj", "out_proj")):
            from .ops import _OPERAND_DTYPE  # noqa: PLC0415

            bias, kpm = self._masks(attn_mask, key_padding_mask, B, H, Lq, Lk, query.device)
            p_drop = self.dropout if self.training else 0.0
            seed = self.dropout_seed_override
            if seed is None:
                seed = self._next_seed() if p_drop > 0 else 0
            out = _FusedMHAFn.apply(q_in, k_in, v_in, same_qk, same_kv, self._pack16(_OPERAND_DTYPE), B, Lq, Lk, H,
                                    1.0 / math.sqrt(hd), bias, kpm, p_drop, seed)
            out = out.reshape(B, Lq, E)
            if not self.batch_first:
                out = out.transpose(0, 1)
            return out.to(query.dtype), None

        def proj(x, n):
            lo = self._adapter(n + "_proj")
            if lo is None:
                return lora_linear(x, pk["W" + n], pk["b" + n], None, None, 1.0)
            Bp = lo.lora_B.new_zeros(lo.rank, Ep).index_copy(1, idx, lo.lora_B)   # differentiable scatter into the padded layout
            p = lo.dropout_p if self.training else 0.0
            return lora_linear(x, pk["W" + n], pk["b" + n], lo.lora_A, Bp, lo.scaling, dropout_p=p)

        q, k, v = proj(q_in, "q"), proj(k_in, "k"), proj(v_in, "v")
        bias, kpm = self._masks(attn_mask, key_padding_mask, B, H, Lq, Lk, query.device)
        p_drop = self.dropout if self.training else 0.0
        seed = self.dropout_seed_override
        if seed is None:
            seed = self._next_seed() if p_drop > 0 else 0
        o = _AttnCoreFn.apply(q, k, v, B, Lq, Lk, H, 1.0 / math.sqrt(hd), bias, kpm, p_drop, seed)
        lo = self._adapter("out_proj")
        ob = self._out_linear.bias
        if lo is None:
            out = lora_linear(o, pk["Wo"], ob, None, None, 1.0)
        else:
            Ap = lo.lora_A.new_zeros(Ep, lo.rank).index_copy(0, idx, lo.lora_A)
            p = lo.dropout_p if self.training else 0.0
            out = lora_linear(o, pk["Wo"], ob, Ap, lo.lora_B, lo.scaling, dropout_p=p)
        out = out.reshape(B, Lq, E)
        if not self.batch_first:
            out = out.transpose(0, 1)
        return out.to(query.dtype), None


def replace_torch_mha(model: nn.Module, skip: Tuple[str, ...] = ()) -> int:
    """Swap every nn.MultiheadAttention (incl. the reference's MultiheadAttentionWrapper subclass) for the fused
    module, keeping parameters and requires_grad flags.  Module paths containing any of the `skip` substrings, and
    head sizes the kernel does not cover, are left alone.  Returns the number of modules replaced."""
    n = 0
    for name, m in list(model.named_modules()):
        if any(s in name for s in skip):
            continue
        if isinstance(m, nn.MultiheadAttention) and m._qkv_same_embed_dim and m.embed_dim // m.num_heads in (32, 64):
            new = MultiheadAttention(m.embed_dim, m.num_heads, dropout=m.dropout, bias=m.in_proj_bias is not None,
                                     batch_first=m.batch_first)
            new.load_state_dict(m.state_dict())
            new.to(m.in_proj_weight.device)
            new.train(m.training)
            for (_, p_new), (_, p_old) in zip(new.named_parameters(), m.named_parameters()):
                p_new.requires_grad = p_old.requires_grad
            *path, leaf = name.split(".")
            parent = model
            for p in path:
                parent = getattr(parent, p)
            setattr(parent, leaf, new)
            n += 1
    return n
