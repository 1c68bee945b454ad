"""Drop-in for the reference's root-level `lora_layers` module (the one
`train_sam3_lora_native.py:37` imports): same classes, functions, signatures, printed summary and
on-disk state-dict layout, with the math routed to the sm_100a kernels.

Reference API being mirrored (file:line in /root/reference):
    LoRALayer            lora_layers.py:13-55    lora_A [in, r] kaiming-uniform(a=sqrt 5), lora_B [r, out] zeros,
                                                 forward = dropout(x) @ A @ B * (alpha / r)
    LoRALinear           lora_layers.py:58-91    children `.original_layer`, `.lora`
    LoRAConfig           lora_layers.py:94-155   + to_dict()
    apply_lora_to_model  lora_layers.py:158-228  freeze everything, wrap matching nn.Linear by basename,
                                                 component gates by substring of the dotted module name
    get_lora_parameters  :231   count_parameters :248   save_lora_weights :265   load_lora_weights :283

Differences, all additive:
  * Fused projections.  The native SAM3 attention has a fused `qkv` Linear, so the reference's target
    names q_proj/k_proj/v_proj/out_proj match nothing there.  For modules that declare
    `lora_virtual_targets()` (our `vit.Attention`) the same names address the row-slices of `qkv`
    and `proj` as *virtual* children: `...attn.q_proj.lora.lora_A` etc. — the reference's key pattern
    `{module_path}.lora.lora_{A,B}`.  `strict_reference_names=True` reproduces the reference's
    behaviour exactly (no virtual targets; `out_proj` never wrapped).
  * `segmentation_head` is accepted as an alias of the `mask_decoder` gate (the native model's name).
  * LoRALinear.forward on CUDA runs one fused tcgen05 GEMM (adapter up-projection as K-extension);
    there is no CPU fallback for the hot path: CPU tensors raise.
"""
from __future__ import annotations

import math
from typing import Dict, Iterable, List, Optional

import torch
import torch.nn as nn


class LoRALayer(nn.Module):
    """Low-rank adapter holding lora_A [in_features, rank] and lora_B [rank, out_features]."""

    def __init__(self, in_features: int, out_features: int, rank: int = 8, alpha: int = 16, dropout: float = 0.0):
        super().__init__()
        self.rank = rank
        self.alpha = alpha
        self.scaling = alpha / rank
        self.lora_A = nn.Parameter(torch.empty(in_features, rank))
        self.lora_B = nn.Parameter(torch.empty(rank, out_features))
        self.dropout = nn.Dropout(p=dropout) if dropout > 0 else nn.Identity()
        nn.init.kaiming_uniform_(self.lora_A, a=math.sqrt(5))
        nn.init.zeros_(self.lora_B)

    @property
    def dropout_p(self) -> float:
        return self.dropout.p if isinstance(self.dropout, nn.Dropout) else 0.0

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """Adapter branch alone: dropout(x) @ A @ B * scaling (two skinny tcgen05 GEMMs)."""
        from .ops import lora_branch  # noqa: PLC0415

        return lora_branch(self.dropout(x), self.lora_A, self.lora_B, self.scaling, owner=self)


class LoRALinear(nn.Module):
    """Frozen nn.Linear + adapter; `y = W x + b + s * (drop(x) A) B` as ONE fused GEMM on CUDA."""

    def __init__(self, original_layer: nn.Linear, rank: int = 8, alpha: int = 16, dropout: float = 0.0):
        super().__init__()
        self.original_layer = original_layer
        for p in self.original_layer.parameters():
            p.requires_grad = False
        self.lora = LoRALayer(original_layer.in_features, original_layer.out_features, rank=rank, alpha=alpha,
                              dropout=dropout)

    # passthroughs so code that inspects the wrapped layer keeps working
    @property
    def in_features(self) -> int:
        return self.original_layer.in_features

    @property
    def out_features(self) -> int:
        return self.original_layer.out_features

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        from .ops import lora_linear  # noqa: PLC0415

        p = self.lora.dropout_p if self.training else 0.0
        return lora_linear(x, self.original_layer.weight, self.original_layer.bias, self.lora.lora_A, self.lora.lora_B,
                           self.lora.scaling, dropout_p=p)


class LoRAVirtual(nn.Module):
    """A virtual projection of a fused Linear (e.g. the q rows of `qkv`).  It owns only the adapter;
    the owning module (vit.Attention) applies it inside its fused kernel schedule."""

    def __init__(self, in_features: int, out_features: int, rank: int, alpha: int, dropout: float):
        super().__init__()
        self.lora = LoRALayer(in_features, out_features, rank=rank, alpha=alpha, dropout=dropout)

    def forward(self, *a, **k):  # pragma: no cover - never called directly
        raise RuntimeError("LoRAVirtual is applied by its parent module's fused kernel, not called directly")


class LoRAConfig:
    """Which modules get adapters (same constructor and to_dict() as the reference)."""

    def __init__(self, rank: int = 8, alpha: int = 16, dropout: float = 0.0, target_modules: Optional[List[str]] = None,
                 apply_to_vision_encoder: bool = True, apply_to_text_encoder: bool = True,
                 apply_to_geometry_encoder: bool = False, apply_to_detr_encoder: bool = True,
                 apply_to_detr_decoder: bool = True, apply_to_mask_decoder: bool = False,
                 strict_reference_names: bool = False):
        self.rank = rank
        self.alpha = alpha
        self.dropout = dropout
        if target_modules is None:
            target_modules = ["q_proj", "k_proj", "v_proj", "out_proj"]
        self.target_modules = set(target_modules)
        self.apply_to_vision_encoder = apply_to_vision_encoder
        self.apply_to_text_encoder = apply_to_text_encoder
        self.apply_to_geometry_encoder = apply_to_geometry_encoder
        self.apply_to_detr_encoder = apply_to_detr_encoder
        self.apply_to_detr_decoder = apply_to_detr_decoder
        self.apply_to_mask_decoder = apply_to_mask_decoder
        self.strict_reference_names = strict_reference_names

    def to_dict(self) -> Dict:
        return {
            "rank": self.rank,
            "alpha": self.alpha,
            "dropout": self.dropout,
            "target_modules": list(self.target_modules),
            "apply_to_vision_encoder": self.apply_to_vision_encoder,
            "apply_to_text_encoder": self.apply_to_text_encoder,
            "apply_to_geometry_encoder": self.apply_to_geometry_encoder,
            "apply_to_detr_encoder": self.apply_to_detr_encoder,
            "apply_to_detr_decoder": self.apply_to_detr_decoder,
            "apply_to_mask_decoder": self.apply_to_mask_decoder,
        }


_COMPONENT_GATES = (
    (("vision_encoder", "vision_backbone"), "apply_to_vision_encoder"),
    (("text_encoder", "language_backbone"), "apply_to_text_encoder"),
    (("geometry_encoder",), "apply_to_geometry_encoder"),
    (("detr_encoder", "transformer.encoder"), "apply_to_detr_encoder"),
    (("detr_decoder", "transformer.decoder"), "apply_to_detr_decoder"),
    (("mask_decoder",), "apply_to_mask_decoder"),
)


def _component_allows(name: str, config: LoRAConfig) -> bool:
    gates = _COMPONENT_GATES
    if not config.strict_reference_names:
        gates = gates[:-1] + ((("mask_decoder", "segmentation_head"), "apply_to_mask_decoder"),)
    for needles, flag in gates:
        if any(n in name for n in needles) and not getattr(config, flag):
            return False
    return True


def _set_child(model: nn.Module, dotted: str, new: nn.Module) -> None:
    *path, leaf = dotted.split(".")
    parent = model
    for p in path:
        parent = getattr(parent, p)
    setattr(parent, leaf, new)


def apply_lora_to_model(model: nn.Module, config: LoRAConfig) -> nn.Module:
    """Freeze every parameter, then attach adapters to the modules the config selects."""
    for p in model.parameters():
        p.requires_grad = False

    applied: List[str] = []
    # (1) the reference's rule: nn.Linear whose basename is a target (out_proj is always skipped there
    #     because nn.MultiheadAttention reads out_proj.weight directly, lora_layers.py:194-196)
    for name, module in list(model.named_modules()):
        if not isinstance(module, nn.Linear) or not _component_allows(name, config):
            continue
        base = name.rsplit(".", 1)[-1]
        if base not in config.target_modules:
            continue
        if base == "out_proj":
            # torch's nn.MultiheadAttention reads out_proj.weight directly, so the reference never wraps it; our fused
            # MultiheadAttention (mha.py) declares that it understands a wrapped out_proj.
            parent_name = name.rsplit(".", 1)[0] if "." in name else ""
            parent = model.get_submodule(parent_name) if parent_name else model
            if config.strict_reference_names or not getattr(parent, "lora_out_proj_ok", False):
                continue
        _set_child(model, name, LoRALinear(module, rank=config.rank, alpha=config.alpha, dropout=config.dropout))
        applied.append(name)
    # (2) virtual targets on fused projections (skipped in strict mode)
    if not config.strict_reference_names:
        for name, module in list(model.named_modules()):
            fn = getattr(module, "lora_virtual_targets", None)
            if fn is None or not _component_allows(name, config):
                continue
            for vname, (fin, fout) in fn().items():
                if vname not in config.target_modules or hasattr(module, vname):
                    continue
                module.add_module(vname, LoRAVirtual(fin, fout, config.rank, config.alpha, config.dropout))
                applied.append(f"{name}.{vname}" if name else vname)
    for module in model.modules():
        hook = getattr(module, "on_lora_changed", None)
        if hook is not None:
            hook()

    print(f"Applied LoRA to {len(applied)} modules:")
    for n in applied[:10]:
        print(f"  - {n}")
    if len(applied) > 10:
        print(f"  ... and {len(applied) - 10} more")
    return model


def get_lora_parameters(model: nn.Module) -> List[nn.Parameter]:
    out: List[nn.Parameter] = []
    for m in model.modules():
        if isinstance(m, LoRALayer):
            out.extend([m.lora_A, m.lora_B])
    return out


def count_parameters(model: nn.Module) -> Dict[str, int]:
    total = sum(p.numel() for p in model.parameters())
    trainable = sum(p.numel() for p in model.parameters() if p.requires_grad)
    return {
        "total_parameters": total,
        "trainable_parameters": trainable,
        "trainable_percentage": 100 * trainable / total if total > 0 else 0,
    }


def lora_state_dict(model: nn.Module) -> Dict[str, torch.Tensor]:
    """{"<module path>.lora_A" / ".lora_B": tensor} — exactly the keys save_lora_weights writes."""
    sd = {}
    for name, m in model.named_modules():
        if isinstance(m, LoRALayer):
            sd[f"{name}.lora_A"] = m.lora_A
            sd[f"{name}.lora_B"] = m.lora_B
    return sd


def save_lora_weights(model: nn.Module, save_path: str):
    """torch.save of a flat dict, adapters only (fp32, A [in, r], B [r, out])."""
    sd = {k: v.detach().clone() for k, v in lora_state_dict(model).items()}
    torch.save(sd, save_path)
    print(f"Saved LoRA weights to {save_path}")


def load_lora_weights(model: nn.Module, load_path: str):
    sd = torch.load(load_path, map_location="cpu")
    model.load_state_dict(sd, strict=False)
    for module in model.modules():
        hook = getattr(module, "on_lora_changed", None)
        if hook is not None:
            hook()
    print(f"Loaded LoRA weights from {load_path}")
