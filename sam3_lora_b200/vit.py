"""Host-side mirror of the reference's ViT image-encoder trunk (sam3/model/vitdet.py), backed by
the native engine.

Module tree and parameter names are the reference's, so its checkpoints load unchanged
(`backbone.vision_backbone.trunk.*`):

    ViT.patch_embed.proj.weight [D,3,P,P] | pos_embed [1,1+side^2,D] | ln_pre.{weight,bias}
    ViT.blocks.{i}.norm1 | .attn.qkv | .attn.proj | .attn.freqs_cis (complex64 buffer) | .norm2 | .mlp.fc1 | .mlp.fc2

`ViT.forward(images)` returns `[feat]` with feat [B, D, G, G] like the reference (vitdet.py:813-859).
The arithmetic of the whole trunk — patch embed, abs-pos, ln_pre, 32 blocks (LayerNorm, qkv GEMM
with fused LoRA + RoPE, flash attention, proj, MLP with GELU) and its backward — runs in
libsam3b.so; this file only keeps parameters, flattens the adapters into one buffer (so the DDP
exchange is a single all-reduce) and wires autograd.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from . import _lib as L
from .engine import VitEngine, VitSpec, base_tensor_names
from .lora.lora_layer import LinearWithLoRA, LoRALayer as PkgLoRALayer
from .lora_layers import LoRALayer, LoRALinear, LoRAVirtual


class PatchEmbed(nn.Module):
    """Parameter holder for the k=s=patch conv (vitdet.py:299-336); bias-free in SAM3 (model_builder.py:94)."""

    def __init__(self, patch: int, in_chans: int, embed_dim: int):
        super().__init__()
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch, stride=patch, bias=False)


class Mlp(nn.Module):
    """timm Mlp attribute names: fc1 -> GELU -> fc2 (vitdet.py:585-590)."""

    def __init__(self, dim: int, hidden: int):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.act = nn.GELU()
        self.fc2 = nn.Linear(hidden, dim)


def _axial_cis(dim: int, end_x: int, end_y: int, theta: float, scale: float) -> torch.Tensor:
    """complex64 [end_x*end_y, dim/2] table, same construction as compute_axial_cis (vitdet.py:41-57);
    kept only so that the state dict has the reference's `attn.freqs_cis` buffer."""
    freqs = 1.0 / (theta ** (torch.arange(0, dim, 4)[: dim // 4].float() / dim))
    t = torch.arange(end_x * end_y, dtype=torch.float32)
    tx = (t % end_x) * scale
    ty = torch.div(t, end_x, rounding_mode="floor") * scale
    fx, fy = torch.outer(tx, freqs), torch.outer(ty, freqs)
    return torch.cat([torch.polar(torch.ones_like(fx), fx), torch.polar(torch.ones_like(fy), fy)], dim=-1)


class Attention(nn.Module):
    def __init__(self, dim: int, num_heads: int, input_size: Tuple[int, int], rope_pt: int, theta: float):
        super().__init__()
        self.num_heads = num_heads
        self.head_dim = dim // num_heads
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.proj = nn.Linear(dim, dim)
        scale = rope_pt / input_size[0]
        self.register_buffer("freqs_cis", _axial_cis(self.head_dim, input_size[0], input_size[1], theta, scale))

    def lora_virtual_targets(self) -> Dict[str, Tuple[int, int]]:
        d = self.qkv.in_features
        return {"q_proj": (d, d), "k_proj": (d, d), "v_proj": (d, d), "out_proj": (d, d)}


class Block(nn.Module):
    def __init__(self, dim, num_heads, mlp_hidden, window_size, grid, theta, eps):
        super().__init__()
        self.window_size = window_size  # 0 = global attention
        size = (window_size, window_size) if window_size > 0 else (grid, grid)
        self.norm1 = nn.LayerNorm(dim, eps=eps)
        self.attn = Attention(dim, num_heads, size, rope_pt=size[0], theta=theta)  # global blocks: table replaced by ViT
        self.norm2 = nn.LayerNorm(dim, eps=eps)
        self.mlp = Mlp(dim, mlp_hidden)


class _GraphState:
    """Static buffers + captured graphs of one (engine binding, batch, adapter buffer) combination."""

    def __init__(self, key, B, spec, device, has_drop):
        self.key = key
        self.img = torch.empty(B, spec.in_chans, spec.img_size, spec.img_size, device=device, dtype=torch.float32)
        self.out = torch.empty(B, spec.embed_dim, spec.grid, spec.grid, device=device, dtype=torch.float32)
        self.gout = torch.empty_like(self.out)
        self.drop = torch.ones(spec.depth, 2, B, device=device, dtype=torch.float32) if has_drop else None
        self.seed_dev = torch.zeros(1, device=device, dtype=torch.int32)   # per-step adapter-dropout seed (read by the kernels)
        self.g_fwd = self.g_bwd = None
        self.warm = False        # one eager step has run (first-use attribute calls, allocator warm-up)


class _TrunkFn(torch.autograd.Function):
    """autograd node for the whole trunk: forward/backward are one C-ABI call each (or one CUDA-graph replay each)."""

    @staticmethod
    def forward(ctx, images, vit: "ViT", need_grad: bool, *lora_params):
        with torch.cuda.nvtx.range("sam3b.trunk.forward"):           # NVTX ranges: visible in nsys / ncu --nvtx timelines
            return _TrunkFn._forward(ctx, images, vit, need_grad, *lora_params)

    @staticmethod
    def _forward(ctx, images, vit: "ViT", need_grad: bool, *lora_params):
        eng = vit._engine_for(images)
        B = images.shape[0]
        eng.bind(images.device, B, training=need_grad or vit._keep_training_workspace)
        if not getattr(eng, "_loaded", False):
            eng.load_base(vit._base_tensors())
        flat = vit._sync_flat()
        spec = vit.spec
        drop = vit._drop_scales_for(B, images.device)
        p_drop = vit._lora_dropout_p if (vit.training and need_grad) else 0.0
        ctx.vit = vit
        ctx.n = len(lora_params)
        ctx.graph = None
        # the engine keeps ONE set of saved activations: a backward must belong to the latest saving forward
        if need_grad:
            vit._fwd_generation += 1
        ctx.generation = vit._fwd_generation if need_grad else -1
        if vit.cuda_graphs and need_grad and flat is not None:
            key = (str(images.device), B, eng.work_buf.data_ptr(), eng.weight_buf.data_ptr(), flat.data_ptr(), drop is not None, p_drop)
            st = vit._graph_state
            if st is None or st.key != key:
                st = vit._graph_state = _GraphState(key, B, spec, images.device, drop is not None)
            st.img.copy_(images.detach())
            if drop is not None:
                st.drop.copy_(drop)
            eng.set_drop_path(st.drop)
            if p_drop > 0.0:
                # adapter dropout under graph replay: the per-step seed lives in device memory (the kernels add it to the
                # captured base seed), redrawn here on the device - no host round trip, reproducible under torch.manual_seed
                if vit.lora_dropout_seed_override is not None:
                    st.seed_dev.fill_(int(vit.lora_dropout_seed_override) & 0x7FFFFFFF)
                else:
                    st.seed_dev.random_(0, 2 ** 31 - 1)
                eng.set_lora_dropout(p_drop, 0, st.seed_dev)
            else:
                eng.set_lora_dropout(0.0, 0)
            if not st.warm:
                eng.forward(st.img, flat, st.out, save_for_backward=True)
            elif st.g_fwd is None:
                st.g_fwd = torch.cuda.CUDAGraph()
                with torch.cuda.graph(st.g_fwd, capture_error_mode="thread_local"):   # other threads (NCCL watchdog) may touch CUDA
                    eng.forward(st.img, flat, st.out, save_for_backward=True)
                st.g_fwd.replay()
            else:
                st.g_fwd.replay()
            ctx.graph = st
            return st.out.clone()
        out = torch.empty(B, spec.embed_dim, spec.grid, spec.grid, device=images.device, dtype=torch.float32)
        img = images.detach().float().contiguous()
        ctx.drop = drop   # keeps the tensor alive until backward
        eng.set_drop_path(ctx.drop)
        seed = vit.lora_dropout_seed_override
        if seed is None:
            seed = int(torch.randint(0, 2 ** 31 - 1, (1,)).item())
        eng.set_lora_dropout(p_drop, seed)
        eng.forward(img, flat, out, save_for_backward=need_grad)
        return out

    @staticmethod
    def backward(ctx, gout):
        with torch.cuda.nvtx.range("sam3b.trunk.backward"):
            return _TrunkFn._backward(ctx, gout)

    @staticmethod
    def _backward(ctx, gout):
        vit: "ViT" = ctx.vit
        eng = vit._engine
        if ctx.generation != vit._fwd_generation:
            raise L.Sam3bError("ViT backward: another trunk forward ran after the one this backward belongs to; the engine keeps a "
                               "single set of saved activations (one in-flight forward per ViT)")
        gflat = vit._flat_grad_buffer()
        st = ctx.graph
        hook = vit.grad_hook
        nseg = int(getattr(hook, "segments", 1)) if hook is not None and hasattr(hook, "reduce_slice") else 1
        segs = eng.segments(nseg)
        g_in = gout.float().contiguous() if st is None else st.gout
        if st is not None:
            st.gout.copy_(gout)

        def run_segment(k):
            hi, lo = segs[k]
            if len(segs) == 1:
                eng.backward(g_in, gflat)
            else:
                eng.backward_segment(g_in if k == 0 else None, gflat, hi, lo)

        for k in range(len(segs)):
            if st is None or not st.warm:
                run_segment(k)
            else:
                key = (gflat.data_ptr(), len(segs))
                if st.g_bwd is None or st.g_bwd[0] != key:
                    st.g_bwd = (key, {})
                graphs = st.g_bwd[1]
                if k not in graphs:
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, capture_error_mode="thread_local"):
                        run_segment(k)
                    graphs[k] = g
                graphs[k].replay()
            if len(segs) > 1:
                a, b = eng.grad_range(*segs[k])
                hook.reduce_slice(gflat[a:b])
        if st is not None:
            st.warm = True
        if len(segs) > 1:
            hook.finish()
            return (None, None, None, *vit._grads_from_flat(gflat))
        vit._after_backward(gflat)
        return (None, None, None, *vit._grads_from_flat(gflat))


class ViT(nn.Module):
    """Drop-in for sam3.model.vitdet.ViT as configured by sam3/model_builder.py:69-96."""

    # keyword arguments of sam3.model.vitdet.ViT.__init__ (vitdet.py:623-657) whose only supported value is the one
    # sam3/model_builder.py:69-96 passes: accepted so `_create_vit_backbone` can call this class unchanged, validated so
    # a different architecture raises instead of silently computing something else.
    _FIXED_REFERENCE_KWARGS = {
        "norm_layer": ("LayerNorm",), "act_layer": (nn.GELU,), "qkv_bias": (True,), "use_abs_pos": (True,),
        "tile_abs_pos": (True,), "rel_pos_blocks": ((), [], False), "rel_pos_zero_init": (True, False), "use_rope": (True,),
        "use_interp_rope": (True,), "rope_pt_size": (None,), "pretrain_use_cls_token": (True,), "retain_cls_token": (False,),
        "dropout": (0.0,), "return_interm_layers": (False,), "init_values": (None,), "ln_pre": (True,), "ln_post": (False,),
        "bias_patch_embed": (False,), "compile_mode": (None,), "use_act_checkpoint": (True, False),
    }

    def __init__(self, img_size=1008, patch_size=14, in_chans=3, embed_dim=1024, depth=32, num_heads=16,
                 mlp_ratio=4.625, window_size=24, global_att_blocks=(7, 15, 23, 31), pretrain_img_size=336,
                 ln_eps=1e-5, rope_theta=10000.0, drop_path_rate=0.1, operand_dtype=torch.float16, max_batch=8,
                 cuda_graphs: bool = False, **reference_kwargs):
        for k, v in reference_kwargs.items():
            allowed = self._FIXED_REFERENCE_KWARGS.get(k)
            if allowed is None:
                raise TypeError(f"ViT.__init__() got an unexpected keyword argument {k!r}")
            if not any(v is a or v == a for a in allowed):
                raise L.Sam3bError(f"ViT({k}={v!r}) is not supported by the native trunk (SAM3 builds it with {allowed[0]!r}; "
                                   "use_act_checkpoint is accepted and ignored: nothing is recomputed)")
        super().__init__()
        # cuda_graphs=True: after one eager warm-up step the ~1280 launches of the trunk forward and of its backward are
        # each replayed from a CUDA graph (training mode; inputs, DropPath scales and the adapter-dropout seed live in static
        # device buffers that are rewritten before every replay).
        self.cuda_graphs = bool(cuda_graphs)
        self._graph_state = None
        # stochastic depth decay rule of the reference: linspace(0, rate, depth) (vitdet.py:746), 0.1 in SAM3
        # (model_builder.py:80); active in train() mode only.
        self.drop_path_rates = [drop_path_rate * i / (depth - 1) for i in range(depth)] if depth > 1 else [float(drop_path_rate)]
        self.drop_scales_override: Optional[torch.Tensor] = None  # tests: inject [depth, 2, B] branch scales
        self.lora_dropout_seed_override: Optional[int] = None     # tests: fix the adapter-dropout mask seed
        self._lora_dropout_p = 0.0
        self.spec = VitSpec(img_size=img_size, patch_size=patch_size, in_chans=in_chans, embed_dim=embed_dim, depth=depth,
                            num_heads=num_heads, mlp_hidden=int(embed_dim * mlp_ratio), window_size=window_size,
                            global_blocks=tuple(global_att_blocks), pretrain_img_size=pretrain_img_size, ln_eps=ln_eps,
                            rope_theta=rope_theta)
        self.full_attn_ids = list(global_att_blocks)
        self.channel_list = [embed_dim]
        self.operand_dtype = operand_dtype
        self.max_batch = max_batch
        g = self.spec.grid
        self.patch_embed = PatchEmbed(patch_size, in_chans, embed_dim)
        self.pos_embed = nn.Parameter(torch.zeros(1, self.spec.pos_side ** 2 + 1, embed_dim))
        self.ln_pre = nn.LayerNorm(embed_dim, eps=ln_eps)
        self.blocks = nn.ModuleList([
            Block(embed_dim, num_heads, self.spec.mlp_hidden, 0 if i in global_att_blocks else window_size, g, rope_theta,
                  ln_eps) for i in range(depth)])
        # global blocks: rope interpolated from the window size (rope_interp, vitdet.py:438-447)
        for i in global_att_blocks:
            self.blocks[i].attn.freqs_cis = _axial_cis(embed_dim // num_heads, g, g, rope_theta, window_size / g)
        nn.init.trunc_normal_(self.pos_embed, std=0.02)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, std=0.02)
                nn.init.zeros_(m.bias)
        self._engine: Optional[VitEngine] = None
        self._engine_key = None
        self._flat: Optional[torch.Tensor] = None
        self._flat_grad: Optional[torch.Tensor] = None
        self._flat_index: List[Tuple[int, int, torch.Size]] = []
        self._lora_params: List[nn.Parameter] = []
        self._flat_transposed: List[bool] = []
        self._keep_training_workspace = False
        self._fwd_generation = 0
        self.grad_hook = None  # callable(flat_grad) run right after the backward kernels (DDP all-reduce)
        self.register_load_state_dict_post_hook(lambda module, incompatible: module.refresh_base())

    def refresh_base(self):
        """Frozen weights changed (checkpoint load): re-pack them into the engine on the next forward."""
        if self._engine is not None:
            self._engine._loaded = False

    # ---- adapters -----------------------------------------------------------------------------
    def on_lora_changed(self):
        """Called by apply_lora_to_model / load_lora_weights: the engine layout must be rebuilt."""
        self._engine_key = None
        self._flat = None

    def _adapter_layers(self) -> List[Tuple[int, str, nn.Module]]:
        """(block, target, adapter) of every adapter in the trunk.  Adapters of the package-layout API
        (lora.LinearWithLoRA on fc1 / fc2: factors stored [r, in] / [out, r]) are accepted as well; `_is_transposed`
        tells the two layouts apart."""
        out = []
        for i, blk in enumerate(self.blocks):
            for t in ("q_proj", "k_proj", "v_proj", "out_proj"):
                m = getattr(blk.attn, t, None)
                if isinstance(m, LoRAVirtual):
                    out.append((i, t, m.lora))
            for t in ("fc1", "fc2"):
                m = getattr(blk.mlp, t)
                if isinstance(m, (LoRALinear, LinearWithLoRA)):
                    out.append((i, t, m.lora))
            for t in ("qkv", "proj"):
                if isinstance(getattr(blk.attn, t), (LoRALinear, LinearWithLoRA)):
                    raise L.Sam3bError(f"blocks.{i}.attn.{t} is wrapped as a whole; the trunk adapts the fused projections through "
                                       "the q_proj / k_proj / v_proj / out_proj targets of lora_layers.apply_lora_to_model")
        return out

    @staticmethod
    def _is_transposed(lora) -> bool:
        return isinstance(lora, PkgLoRALayer)

    def _linear(self, m) -> nn.Linear:
        if isinstance(m, LoRALinear):
            return m.original_layer
        return m.linear if isinstance(m, LinearWithLoRA) else m

    def _base_tensors(self) -> Dict[str, torch.Tensor]:
        t = {"patch_embed.proj.weight": self.patch_embed.proj.weight, "pos_embed": self.pos_embed,
             "ln_pre.weight": self.ln_pre.weight, "ln_pre.bias": self.ln_pre.bias}
        for i, b in enumerate(self.blocks):
            p = f"blocks.{i}."
            for n, m in (("norm1", b.norm1), ("attn.qkv", b.attn.qkv), ("attn.proj", b.attn.proj), ("norm2", b.norm2),
                         ("mlp.fc1", self._linear(b.mlp.fc1)), ("mlp.fc2", self._linear(b.mlp.fc2))):
                t[p + n + ".weight"] = m.weight
                t[p + n + ".bias"] = m.bias
        return t

    def _engine_for(self, images: torch.Tensor) -> VitEngine:
        if not images.is_cuda:
            raise L.Sam3bError("ViT.forward: input is on the CPU; the trunk has no CPU fallback (needs a CUDA device)")
        layers = self._adapter_layers()
        targets = sorted({t for _, t, _ in layers})
        ranks = {l.rank for _, _, l in layers}
        scal = {float(l.scaling) for _, _, l in layers}
        if len(ranks) > 1 or len(scal) > 1:
            raise L.Sam3bError("all trunk adapters must share one rank and alpha")
        per_block = {}
        for i, t, _ in layers:
            per_block.setdefault(i, set()).add(t)
        if layers and (len(per_block) != self.spec.depth or any(v != set(targets) for v in per_block.values())):
            raise L.Sam3bError("the trunk engine needs the same adapter targets in every block")
        pset = {float(l.dropout_p) for _, _, l in layers}
        if len(pset) > 1:
            raise L.Sam3bError("all trunk adapters must share one dropout probability")
        self._lora_dropout_p = pset.pop() if pset else 0.0
        rank = ranks.pop() if ranks else 0
        scaling = scal.pop() if scal else 1.0
        key = (tuple(targets), rank, scaling, self.operand_dtype, str(images.device))
        if self._engine is None or self._engine_key != key:
            self._engine = VitEngine(self.spec, lora_rank=rank, lora_scaling=scaling, lora_targets=targets,
                                     dtype=self.operand_dtype, max_batch=self.max_batch)
            self._engine_key = key
            self._flat = None
        return self._engine

    def lora_parameters(self) -> List[nn.Parameter]:
        self._build_index()
        return list(self._lora_params)

    def _build_index(self):
        """Order adapters as the engine's flat layout: per block q,k,v,o,fc1,fc2; A [in, r] then B [r, out]."""
        layers = {(i, t): l for i, t, l in self._adapter_layers()}
        index, params, transposed = [], [], []
        eng = self._engine
        if eng is None:
            order = sorted(layers, key=lambda k: (k[0], ("q_proj", "k_proj", "v_proj", "out_proj", "fc1", "fc2").index(k[1])))
            off = 0
            for k in order:
                l = layers[k]
                for p in (l.lora_A, l.lora_B):
                    index.append((off, p.numel(), p.shape))
                    params.append(p)
                    transposed.append(self._is_transposed(l))
                    off += p.numel()
        else:
            for e in eng.entries:
                l = layers[(e.block, e.target)]
                index.append((e.a_off, l.lora_A.numel(), l.lora_A.shape))
                index.append((e.b_off, l.lora_B.numel(), l.lora_B.shape))
                params += [l.lora_A, l.lora_B]
                transposed += [self._is_transposed(l)] * 2
        self._flat_index, self._lora_params, self._flat_transposed = index, params, transposed

    def _sync_flat(self) -> Optional[torch.Tensor]:
        """Make every adapter Parameter a view of one flat fp32 CUDA buffer (values preserved).  Package-layout factors
        ([r, in] / [out, r]) cannot alias the engine's [in, r] / [r, out] layout: they are copied in (transposed) before
        every forward, and their gradients are transposed back (a few MB per step)."""
        eng = self._engine
        if eng.lora_numel == 0:
            return None
        self._build_index()
        dev = self.pos_embed.device
        ok = self._flat is not None and self._flat.device == dev
        if ok:
            base = self._flat.data_ptr()
            ok = all(tr or (p.data_ptr() == base + 4 * a and p.dtype == torch.float32) for p, (a, _, _), tr in
                     zip(self._lora_params, self._flat_index, self._flat_transposed))
        if not ok:
            flat = torch.zeros(eng.lora_numel, device=dev, dtype=torch.float32)
            for p, (a, n, shape), tr in zip(self._lora_params, self._flat_index, self._flat_transposed):
                if tr:
                    continue
                flat[a:a + n] = p.detach().reshape(-1).to(device=dev, dtype=torch.float32)
                p.data = flat[a:a + n].view(shape)
            self._flat = flat
            self._flat_grad = None
        for p, (a, n, shape), tr in zip(self._lora_params, self._flat_index, self._flat_transposed):
            if tr:
                self._flat[a:a + n] = p.detach().t().reshape(-1).to(device=dev, dtype=torch.float32)
        return self._flat

    def _grads_from_flat(self, gflat: torch.Tensor) -> List[torch.Tensor]:
        """Per-parameter gradients for autograd.  They are views of ONE fresh copy of the flat buffer (a single 40 MB
        device copy), never of the persistent buffer itself: autograd's AccumulateGrad may keep what it is handed as p.grad,
        and the next backward overwrites the persistent buffer (gradient accumulation over micro-batches or
        zero_grad(set_to_none=False) would otherwise read aliased memory)."""
        g = gflat.clone()
        return [g[a:a + n].view(shape[::-1]).t().contiguous() if tr else g[a:a + n].view(shape)
                for (a, n, shape), tr in zip(self._flat_index, self._flat_transposed)]

    def flat_lora(self) -> Optional[torch.Tensor]:
        return self._flat

    def _flat_grad_buffer(self) -> torch.Tensor:
        if self._flat_grad is None or self._flat_grad.numel() != self._flat.numel():
            self._flat_grad = torch.zeros_like(self._flat)
        return self._flat_grad

    def _after_backward(self, gflat: torch.Tensor):
        if self.grad_hook is not None:
            self.grad_hook(gflat)

    def _drop_scales_for(self, batch: int, device) -> Optional[torch.Tensor]:
        """DropPath (timm semantics, scale_by_keep): one Bernoulli(keep) draw per sample and per residual branch;
        returns [depth, 2, batch] with entries 0 or 1/keep, or None when inactive."""
        if self.drop_scales_override is not None:
            return self.drop_scales_override.to(device=device, dtype=torch.float32).contiguous()
        if not self.training or max(self.drop_path_rates) <= 0.0:
            return None
        keep = 1.0 - torch.tensor(self.drop_path_rates, device=device, dtype=torch.float32).view(-1, 1, 1)
        mask = (torch.rand(self.spec.depth, 2, batch, device=device) < keep).float()
        return (mask / keep).contiguous()

    # ---- forward ------------------------------------------------------------------------------
    def forward(self, x: torch.Tensor) -> List[torch.Tensor]:
        self._engine_for(x)
        self._build_index()
        need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self._lora_params)
        return [_TrunkFn.apply(x, self, need_grad, *self._lora_params)]


def build_sam3_vit(**overrides) -> ViT:
    """The one trunk SAM3 ships (sam3/model_builder.py:69-96)."""
    return ViT(**overrides)
