"""CPU oracle for the nn.MultiheadAttention sites (row a7).  TEST INFRASTRUCTURE ONLY.

Restates torch.nn.MultiheadAttention's forward (the arithmetic behind sam3/model/model_misc.py:31-34
`MultiheadAttentionWrapper`, need_weights=False) with the LoRA virtual targets of this repo:
    q = query Wq^T + bq (+ s * query A_q B_q)   k, v likewise from key / value
    P = softmax(q k^T / sqrt(hd) + attn_mask + (-inf where key_padding_mask))      per head
    out = (dropout(P) v) Wo^T + bo (+ s * o A_o B_o)
It is pinned against torch's own module in tests/test_mha_oracle.py (torch is the un-vendored dependency that owns
this arithmetic; the reference only calls it).  Dropout uses the CUDA path's stateless hash (csrc/attn_fwd.cu
attn_drop_keep) so that p > 0 can be compared element-for-element.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch

from .vit_oracle import _lowbias32, lora_delta

Tensor = torch.Tensor


def attn_drop_scale_mask(seed: int, B: int, H: int, Lq: int, Lk: int, p: float) -> Tensor:
    """[B, H, Lq, Lk] inverted-dropout scale mask identical to the kernels' (0 or 1/(1-p))."""
    thr = min(int(p * 4294967296.0), 0xFFFFFFFF)
    bh = np.arange(B * H, dtype=np.uint32)
    with np.errstate(over="ignore"):
        base = _lowbias32(np.uint32(seed) ^ (bh * np.uint32(0x9E3779B1) + np.uint32(0x85EBCA6B)))
        qk = (np.arange(Lq, dtype=np.uint32).reshape(-1, 1) * np.uint32(Lk) + np.arange(Lk, dtype=np.uint32).reshape(1, -1))
        h = _lowbias32(base.reshape(-1, 1, 1) ^ qk.reshape(1, Lq, Lk))
    keep = (h >= np.uint32(thr)).astype(np.float32) / (1.0 - p)
    return torch.from_numpy(keep).reshape(B, H, Lq, Lk)


def mha_forward(query: Tensor, key: Tensor, value: Tensor, params: Dict[str, Tensor], num_heads: int,
                attn_mask: Optional[Tensor] = None, key_padding_mask: Optional[Tensor] = None, scaling: float = 1.0,
                dropout=None) -> Tensor:
    """Batch-first [B, L, E] tensors.  params: in_proj_weight, in_proj_bias, out_proj.weight, out_proj.bias and optional
    '{q,k,v,out}_proj.lora.lora_{A,B}'.  dropout = (p, seed) or None."""
    B, Lq, E = query.shape
    Lk = key.shape[1]
    H, hd = num_heads, E // num_heads
    W, b = params["in_proj_weight"], params["in_proj_bias"]

    def proj(x, i, name):
        y = x @ W[i * E:(i + 1) * E].T + b[i * E:(i + 1) * E]
        ka = f"{name}.lora.lora_A"
        if ka in params:
            y = y + lora_delta(x, params[ka], params[f"{name}.lora.lora_B"], scaling)
        return y

    q = proj(query, 0, "q_proj").reshape(B, Lq, H, hd).permute(0, 2, 1, 3)
    k = proj(key, 1, "k_proj").reshape(B, Lk, H, hd).permute(0, 2, 1, 3)
    v = proj(value, 2, "v_proj").reshape(B, Lk, H, hd).permute(0, 2, 1, 3)
    s = (q @ k.transpose(-1, -2)) / math.sqrt(hd)
    if attn_mask is not None:
        s = s + attn_mask.reshape(B, H, Lq, Lk) if attn_mask.dim() == 3 else s + attn_mask
    if key_padding_mask is not None:
        s = s.masked_fill(key_padding_mask.bool().view(B, 1, 1, Lk), float("-inf"))
    P = torch.softmax(s, dim=-1)
    if dropout is not None and dropout[0] > 0:
        P = P * attn_drop_scale_mask(dropout[1], B, H, Lq, Lk, dropout[0]).to(P.dtype)
    o = (P @ v).permute(0, 2, 1, 3).reshape(B, Lq, E)
    y = o @ params["out_proj.weight"].T + params["out_proj.bias"]
    if "out_proj.lora.lora_A" in params:
        y = y + lora_delta(o, params["out_proj.lora.lora_A"], params["out_proj.lora.lora_B"], scaling)
    return y
