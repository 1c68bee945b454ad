"""CPU oracle for the SAM3 ViT + LoRA hot path.  TEST INFRASTRUCTURE ONLY.

A plain-PyTorch (CPU, fp32 or fp64) restatement of the reference arithmetic, written as pure
functions over a flat parameter dict that uses the reference's own state-dict key names.  It is
the checker for the CUDA path: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` leg may import it; nothing under sam3_lora_b200/ does.

Pinned against the reference: tests/golden/make_golden.py imports the real reference modules
(/root/reference: sam3/model/vitdet.py `ViT`, lora_layers.py `LoRALinear`) in the build container
and commits their outputs as tests/golden/*.npz; tests/test_oracle.py checks this file against those
vectors.  (The reference ships no golden vectors of its own — SURVEY.md §4, §8c.)

Each function cites the reference lines it restates.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


@dataclass
class ViTConfig:
    """Hyper-parameters of sam3/model/vitdet.py `ViT` (defaults = sam3/model_builder.py:69-96)."""
    img_size: int = 1008
    patch_size: int = 14
    embed_dim: int = 1024
    depth: int = 32
    num_heads: int = 16
    mlp_hidden: int = 4736            # int(1024 * 4.625), vitdet.py:587
    window_size: int = 24
    global_att_blocks: Tuple[int, ...] = (7, 15, 23, 31)
    pretrain_img_size: int = 336      # pos_embed grid = pretrain/patch (+1 cls slot), vitdet.py:737-743
    ln_eps: float = 1e-5              # vitdet.py:719
    rope_theta: float = 10000.0

    @property
    def grid(self) -> int:
        return self.img_size // self.patch_size

    @property
    def head_dim(self) -> int:
        return self.embed_dim // self.num_heads


@dataclass
class LoRASpec:
    """Which virtual projections carry adapters (north-star aliasing)."""
    rank: int = 16
    alpha: float = 32.0
    targets: Tuple[str, ...] = ("q_proj", "k_proj", "v_proj", "out_proj", "fc1", "fc2")

    @property
    def scaling(self) -> float:
        return self.alpha / self.rank   # lora_layers.py:37


# --------------------------------------------------------------------------------------------
# RoPE (vitdet.py:32-90)
# --------------------------------------------------------------------------------------------
def axial_rope_angles(dim: int, end_x: int, end_y: int, theta: float, scale_pos: float) -> Tensor:
    """Angles [end_x*end_y, dim/2]: first dim/4 columns = x-axis, last dim/4 = y-axis (compute_axial_cis,
    vitdet.py:41-57; init_t_xy :32-38).  float64 so that cos/sin are accurate before any rounding."""
    freqs = 1.0 / (theta ** (torch.arange(0, dim, 4)[: dim // 4].double() / dim))
    t = torch.arange(end_x * end_y, dtype=torch.float64)
    t_x = (t % end_x) * scale_pos
    t_y = torch.div(t, end_x, rounding_mode="floor") * scale_pos
    return torch.cat([torch.outer(t_x, freqs), torch.outer(t_y, freqs)], dim=-1)


def apply_rope(x: Tensor, ang: Tensor) -> Tensor:
    """x [..., L, hd]: complex multiply of adjacent pairs (x[2i], x[2i+1]) by exp(i*ang[:, i])
    (apply_rotary_enc, vitdet.py:68-90)."""
    xr = x.reshape(*x.shape[:-1], -1, 2)
    a, b = xr[..., 0], xr[..., 1]
    c, s = ang.cos().to(device=x.device, dtype=x.dtype), ang.sin().to(device=x.device, dtype=x.dtype)
    return torch.stack([a * c - b * s, a * s + b * c], dim=-1).reshape(x.shape)


def rope_angles_for_block(cfg: ViTConfig, is_global: bool) -> Tensor:
    """Window blocks: rope over the 24x24 window, scale 1.  Global blocks: 72x72 with
    scale_pos = window/grid (rope_interp, vitdet.py:438-447; Block wiring :762-769)."""
    if is_global:
        return axial_rope_angles(cfg.head_dim, cfg.grid, cfg.grid, cfg.rope_theta, cfg.window_size / cfg.grid)
    return axial_rope_angles(cfg.head_dim, cfg.window_size, cfg.window_size, cfg.rope_theta, 1.0)


# --------------------------------------------------------------------------------------------
# LoRA (lora_layers.py:49-55, 87-91)
# --------------------------------------------------------------------------------------------
def lora_delta(x: Tensor, A: Tensor, B: Tensor, scaling: float, mask: Optional[Tensor] = None) -> Tensor:
    """(dropout(x) @ A @ B) * alpha/r with A:[in,r], B:[r,out]; `mask` = inverted-dropout scale mask on x
    (adapter branch only, lora_layers.py:54) or None."""
    if mask is not None:
        x = x * mask.to(x.dtype)
    return (x @ A @ B) * scaling


# adapter dropout: the CUDA path draws its mask from a stateless hash (sam3_lora_b200/csrc/rng.cuh); restated here
# so that parity runs with p > 0 see the identical mask.  Rows are the engine's window-major token indices.
def _lowbias32(x):
    import numpy as np

    x = x.astype(np.uint32)
    x ^= x >> np.uint32(16); x *= np.uint32(0x7feb352d)
    x ^= x >> np.uint32(15); x *= np.uint32(0x846ca68b)
    x ^= x >> np.uint32(16)
    return x


def site_seed(seed: int, block_idx: int, site: int) -> int:
    return (seed + 0x9E3779B9 * (block_idx * 4 + site + 1)) & 0xFFFFFFFF


def dropout_scale_mask(rows: Tensor, cols: int, p: float, seed: int) -> Tensor:
    """[..., cols] float mask: 1/(1-p) where kept, 0 where dropped; `rows` = engine row index of each token."""
    import numpy as np

    r = rows.reshape(-1, 1).numpy().astype(np.uint32)
    c = np.arange(cols, dtype=np.uint32).reshape(1, -1)
    with np.errstate(over="ignore"):
        h = _lowbias32(np.uint32(seed) ^ _lowbias32(r * np.uint32(cols) + c))
    thr = min(int(p * 4294967296.0), 0xFFFFFFFF)
    keep = h >= np.uint32(thr)
    return torch.from_numpy(keep.astype(np.float32) / (1.0 - p)).reshape(*rows.shape, cols)


def engine_row_index(cfg: "ViTConfig", batch: int) -> Tensor:
    """[B, G, G] window-major row index of every token (the order the CUDA engine stores tokens in)."""
    G, ws = cfg.grid, cfg.window_size
    nwx = G // ws
    pi = torch.arange(G).view(G, 1).expand(G, G)
    pj = torch.arange(G).view(1, G).expand(G, G)
    tok = ((pi // ws) * nwx + pj // ws) * (ws * ws) + (pi % ws) * ws + (pj % ws)
    return torch.arange(batch).view(-1, 1, 1) * (G * G) + tok.unsqueeze(0)


def _lora(params: Dict[str, Tensor], prefix: str, name: str) -> Optional[Tuple[Tensor, Tensor]]:
    ka, kb = f"{prefix}.{name}.lora.lora_A", f"{prefix}.{name}.lora.lora_B"
    if ka in params:
        return params[ka], params[kb]
    return None


# --------------------------------------------------------------------------------------------
# Attention / Block / ViT (vitdet.py:466-515, 597-613, 813-859)
# --------------------------------------------------------------------------------------------
def attention(x: Tensor, params: Dict[str, Tensor], prefix: str, cfg: ViTConfig, ang: Tensor,
              scaling: float, drop=None) -> Tensor:
    """x [Bw, L, D] -> [Bw, L, D].  qkv Linear, per-head split (reshape(B,L,3,H,hd), vitdet.py:480-482),
    RoPE on q,k (:485), softmax(q k^T / sqrt(hd)) v (:502), proj (:513).  Adapters on the virtual
    q_proj/k_proj/v_proj (row slices of qkv) and out_proj (= proj)."""
    Bw, L, D = x.shape
    H, hd = cfg.num_heads, cfg.head_dim
    W, b = params[f"{prefix}.qkv.weight"], params[f"{prefix}.qkv.bias"]
    qkv = x @ W.T + b
    # drop = (rows [Bw, L], p, seed, block index): one mask per site, shared by the q/k/v adapters (they read the
    # same masked activation in the fused kernel)
    m_in = dropout_scale_mask(drop[0], D, drop[1], site_seed(drop[2], drop[3], 0)) if drop else None
    m_o = dropout_scale_mask(drop[0], D, drop[1], site_seed(drop[2], drop[3], 1)) if drop else None
    for i, name in enumerate(("q_proj", "k_proj", "v_proj")):
        ab = _lora(params, prefix, name)
        if ab is not None:
            qkv = torch.cat([qkv[..., : i * D], qkv[..., i * D:(i + 1) * D] + lora_delta(x, ab[0], ab[1], scaling, m_in),
                             qkv[..., (i + 1) * D:]], dim=-1)
    qkv = qkv.reshape(Bw, L, 3, H, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    q, k = apply_rope(q, ang), apply_rope(k, ang)
    att = torch.softmax((q @ k.transpose(-1, -2)) * (hd ** -0.5), dim=-1)
    o = (att @ v).permute(0, 2, 1, 3).reshape(Bw, L, D)
    y = o @ params[f"{prefix}.proj.weight"].T + params[f"{prefix}.proj.bias"]
    ab = _lora(params, prefix, "out_proj")
    if ab is not None:
        y = y + lora_delta(o, ab[0], ab[1], scaling, m_o)
    return y


def window_partition(x: Tensor, ws: int) -> Tensor:
    """[B,G,G,C] -> [B*nw, ws, ws, C] (vitdet.py:93-115; G is a multiple of ws here, no padding)."""
    B, G, _, C = x.shape
    x = x.view(B, G // ws, ws, G // ws, ws, C)
    return x.permute(0, 1, 3, 2, 4, 5).reshape(-1, ws, ws, C)


def window_unpartition(w: Tensor, ws: int, G: int) -> Tensor:
    """inverse of window_partition (vitdet.py:118-139)."""
    B = w.shape[0] // ((G // ws) ** 2)
    x = w.reshape(B, G // ws, G // ws, ws, ws, -1)
    return x.permute(0, 1, 3, 2, 4, 5).reshape(B, G, G, -1)


def block(x: Tensor, params: Dict[str, Tensor], i: int, cfg: ViTConfig, scaling: float,
          prefix: str = "", drop: Optional[Tensor] = None, lora_dropout=None) -> Tensor:
    """Pre-norm residual block (vitdet.py:597-613).  `drop` [2, B]: per-sample DropPath scales (0 or 1/keep)
    of the attention and MLP branches (timm DropPath, scale_by_keep); None = identity (eval / parity mode)."""
    p = f"{prefix}blocks.{i}"
    D = cfg.embed_dim
    G = cfg.grid
    is_global = i in cfg.global_att_blocks
    ang = rope_angles_for_block(cfg, is_global)
    shortcut = x
    rows = engine_row_index(cfg, x.shape[0]) if lora_dropout else None     # lora_dropout = (p, seed)
    h = F.layer_norm(x, (D,), params[f"{p}.norm1.weight"], params[f"{p}.norm1.bias"], cfg.ln_eps)
    if is_global:
        dr = (rows.reshape(x.shape[0], G * G), lora_dropout[0], lora_dropout[1], i) if lora_dropout else None
        a = attention(h.reshape(h.shape[0], G * G, D), params, f"{p}.attn", cfg, ang, scaling, dr).reshape(h.shape)
    else:
        ws = cfg.window_size
        w = window_partition(h, ws)
        dr = None
        if lora_dropout:
            rw = window_partition(rows.unsqueeze(-1), ws).reshape(-1, ws * ws)
            dr = (rw, lora_dropout[0], lora_dropout[1], i)
        a = attention(w.reshape(w.shape[0], ws * ws, D), params, f"{p}.attn", cfg, ang, scaling, dr)
        a = window_unpartition(a.reshape(-1, ws, ws, D), ws, G)
    if drop is not None:
        a = a * drop[0].to(a.dtype).view(-1, 1, 1, 1)
    x = shortcut + a
    h = F.layer_norm(x, (D,), params[f"{p}.norm2.weight"], params[f"{p}.norm2.bias"], cfg.ln_eps)
    h1 = h @ params[f"{p}.mlp.fc1.weight"].T + params[f"{p}.mlp.fc1.bias"]
    ab = _lora(params, f"{p}.mlp", "fc1")
    if ab is not None:
        m1 = dropout_scale_mask(rows, D, lora_dropout[0], site_seed(lora_dropout[1], i, 2)) if lora_dropout else None
        h1 = h1 + lora_delta(h, ab[0], ab[1], scaling, m1)
    g = F.gelu(h1)  # exact erf GELU (nn.GELU default, timm Mlp)
    h2 = g @ params[f"{p}.mlp.fc2.weight"].T + params[f"{p}.mlp.fc2.bias"]
    ab = _lora(params, f"{p}.mlp", "fc2")
    if ab is not None:
        m2 = dropout_scale_mask(rows, cfg.mlp_hidden, lora_dropout[0], site_seed(lora_dropout[1], i, 3)) if lora_dropout else None
        h2 = h2 + lora_delta(g, ab[0], ab[1], scaling, m2)
    if drop is not None:
        h2 = h2 * drop[1].to(h2.dtype).view(-1, 1, 1, 1)
    return x + h2


def patch_embed(img: Tensor, params: Dict[str, Tensor], cfg: ViTConfig, prefix: str = "") -> Tensor:
    """Conv k=s=patch, no bias -> NHWC (vitdet.py:332-336), + tiled abs pos (get_abs_pos tiling,
    vitdet.py:199-222: drop the cls slot, tile the pretrain grid, crop), + ln_pre (:833)."""
    x = F.conv2d(img, params[f"{prefix}patch_embed.proj.weight"], None, stride=cfg.patch_size).permute(0, 2, 3, 1)
    G = cfg.grid
    pos = params[f"{prefix}pos_embed"][:, 1:]
    size = int(math.isqrt(pos.shape[1]))
    pos = pos.reshape(1, size, size, -1)
    if size != G:
        reps = G // size + 1
        pos = pos.permute(0, 3, 1, 2).tile(1, 1, reps, reps)[:, :, :G, :G].permute(0, 2, 3, 1)
    x = x + pos
    return F.layer_norm(x, (cfg.embed_dim,), params[f"{prefix}ln_pre.weight"], params[f"{prefix}ln_pre.bias"], cfg.ln_eps)


def vit_forward(img: Tensor, params: Dict[str, Tensor], cfg: ViTConfig, scaling: float = 1.0,
                prefix: str = "", return_blocks: bool = False, drop_scales: Optional[Tensor] = None, lora_dropout=None):
    """ViT.forward (vitdet.py:813-859): returns the NCHW feature map [B, D, G, G] after the last block
    (ln_post is Identity in SAM3, model_builder.py:92)."""
    x = patch_embed(img, params, cfg, prefix)
    outs = [x]
    for i in range(cfg.depth):
        x = block(x, params, i, cfg, scaling, prefix, None if drop_scales is None else drop_scales[i], lora_dropout)
        if return_blocks:
            outs.append(x)
    y = x.permute(0, 3, 1, 2)
    return (y, outs) if return_blocks else y


# --------------------------------------------------------------------------------------------
# synthetic parameters (no checkpoint is available offline)
# --------------------------------------------------------------------------------------------
def lora_param_names(cfg: ViTConfig, spec: LoRASpec, prefix: str = "") -> List[Tuple[str, int, int]]:
    """[(module_path, in_features, out_features)] in state-dict order."""
    D, Dm = cfg.embed_dim, cfg.mlp_hidden
    out = []
    for i in range(cfg.depth):
        p = f"{prefix}blocks.{i}"
        for t in ("q_proj", "k_proj", "v_proj", "out_proj"):
            if t in spec.targets:
                out.append((f"{p}.attn.{t}", D, D))
        if "fc1" in spec.targets:
            out.append((f"{p}.mlp.fc1", D, Dm))
        if "fc2" in spec.targets:
            out.append((f"{p}.mlp.fc2", Dm, D))
    return out


def make_params(cfg: ViTConfig, spec: Optional[LoRASpec], seed: int = 0, dtype=torch.float32,
                lora_b_std: float = 0.02, prefix: str = "") -> Dict[str, Tensor]:
    """Seeded synthetic weights: Linear/pos ~ N(0, 0.02) as trunc_normal_(std=0.02) (vitdet.py:806,
    _init_weights :796-804) but with non-zero biases and non-trivial LN affine so every term is
    exercised; LoRA A ~ kaiming-uniform(a=sqrt(5)) (lora_layers.py:46), B ~ N(0, lora_b_std) instead
    of zeros so the adapter path is not a no-op."""
    g = torch.Generator().manual_seed(seed)
    D, Dm, P = cfg.embed_dim, cfg.mlp_hidden, cfg.patch_size

    def n(*shape, std=0.02):
        return (torch.randn(*shape, generator=g, dtype=torch.float64) * std).to(dtype)

    params: Dict[str, Tensor] = {}
    params[f"{prefix}patch_embed.proj.weight"] = n(D, 3, P, P)
    n_pos = (cfg.pretrain_img_size // P) ** 2 + 1
    params[f"{prefix}pos_embed"] = n(1, n_pos, D)
    params[f"{prefix}ln_pre.weight"] = 1.0 + n(D, std=0.1)
    params[f"{prefix}ln_pre.bias"] = n(D, std=0.1)
    for i in range(cfg.depth):
        p = f"{prefix}blocks.{i}"
        for ln in ("norm1", "norm2"):
            params[f"{p}.{ln}.weight"] = 1.0 + n(D, std=0.1)
            params[f"{p}.{ln}.bias"] = n(D, std=0.1)
        params[f"{p}.attn.qkv.weight"] = n(3 * D, D)
        params[f"{p}.attn.qkv.bias"] = n(3 * D)
        params[f"{p}.attn.proj.weight"] = n(D, D)
        params[f"{p}.attn.proj.bias"] = n(D)
        params[f"{p}.mlp.fc1.weight"] = n(Dm, D)
        params[f"{p}.mlp.fc1.bias"] = n(Dm)
        params[f"{p}.mlp.fc2.weight"] = n(D, Dm)
        params[f"{p}.mlp.fc2.bias"] = n(D)
    if spec is not None:
        for path, fin, fout in lora_param_names(cfg, spec, prefix):
            bound = 1.0 / math.sqrt(spec.rank)  # kaiming_uniform_(a=sqrt(5)) on [in, r]: fan_in = r
            A = (torch.rand(fin, spec.rank, generator=g, dtype=torch.float64) * 2 - 1) * bound
            params[f"{path}.lora.lora_A"] = A.to(dtype)
            params[f"{path}.lora.lora_B"] = n(spec.rank, fout, std=lora_b_std)
    return params


def lora_keys(params: Dict[str, Tensor]) -> List[str]:
    return [k for k in params if k.endswith(".lora.lora_A") or k.endswith(".lora.lora_B")]


def train_step_reference(img: Tensor, params: Dict[str, Tensor], cfg: ViTConfig, spec: LoRASpec,
                         gout: Tensor, drop_scales: Optional[Tensor] = None, lora_dropout=None) -> Tuple[Tensor, Dict[str, Tensor]]:
    """Forward + backward of the trunk with only the adapters trainable (apply_lora_to_model freezes
    everything else, lora_layers.py:171-172).  Loss = sum(out * gout).  Returns (out, {lora key: grad})."""
    keys = lora_keys(params)
    leaf = {k: params[k].detach().clone().requires_grad_(True) for k in keys}
    p = dict(params)
    p.update(leaf)
    out = vit_forward(img, p, cfg, spec.scaling, drop_scales=drop_scales, lora_dropout=lora_dropout)
    (out * gout).sum().backward()
    return out.detach(), {k: leaf[k].grad for k in keys}
