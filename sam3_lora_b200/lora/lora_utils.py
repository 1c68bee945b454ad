"""Drop-in for `sam3_lora.lora.lora_utils` (sam3_lora/lora/lora_utils.py:14-289): the injection API the Hydra trainer uses
(`sam3_lora/train/native_trainer.py:21-26`).

Behaviour kept from the reference:
  * `inject_lora_into_model` wraps every `nn.Linear` whose DOTTED NAME CONTAINS one of the target strings
    (`_should_inject_lora`, :59-92: plain substring test first, so `"self_attn"` selects every Linear below a self-attention
    module) and does NOT freeze anything else — the trainer freezes the base itself;
  * state dict keys `{module path}.lora.lora_A [r, in]`, `{module path}.lora.lora_B [out, r]`;
  * `merge_lora_weights` turns every wrapper back into a plain `nn.Linear` with `W + B A * alpha / r`.
"""
from __future__ import annotations

import re
from typing import Dict, List, Optional, Set

import torch
import torch.nn as nn

from .lora_layer import LinearWithLoRA

_ALL_TARGETS = ("q_proj", "k_proj", "v_proj", "out_proj", "linear1", "linear2", "in_proj", "cross_attn", "self_attn")
_DEFAULT_TARGETS = ("q_proj", "k_proj", "v_proj", "out_proj", "linear1", "linear2")
# second-chance patterns of the reference (:76-90): a pattern only counts when a target string occurs in the PATTERN text
_NAME_PATTERNS = (r".*\.self_attn\.", r".*\.cross_attn\.", r".*\.cross_attn_image\.", r".*\.ca_text\.", r".*\.linear[12]$",
                  r".*\.(q|k|v|out)_proj$")


class LoRAConfig:
    def __init__(self, rank: int = 4, alpha: float = 1.0, dropout: float = 0.0, target_modules: Optional[List[str]] = None):
        self.rank, self.alpha, self.dropout = rank, alpha, dropout
        targets = set(_DEFAULT_TARGETS if target_modules is None else target_modules)
        self.target_modules: Set[str] = set(_ALL_TARGETS) if "all" in targets else targets


def _should_inject_lora(name: str, target_modules: Set[str]) -> bool:
    if any(t in name for t in target_modules):
        return True
    return any(re.match(pat, name) and any(t in pat for t in target_modules) for pat in _NAME_PATTERNS)


def inject_lora_into_model(model: nn.Module, config: LoRAConfig, verbose: bool = True) -> nn.Module:
    hits = [(n, m) for n, m in model.named_modules() if isinstance(m, nn.Linear) and _should_inject_lora(n, config.target_modules)]
    # a Linear that already sits inside a wrapper ("<path>.linear") is the wrapper's own frozen layer, not a new site
    wrapped_children = {f"{n}.linear" for n, m in model.named_modules() if isinstance(m, LinearWithLoRA)}
    n_params = 0
    count = 0
    for name, lin in hits:
        if name in wrapped_children:
            continue
        parent_name, _, leaf = name.rpartition(".")
        parent = model.get_submodule(parent_name) if parent_name else model
        wrapper = LinearWithLoRA(lin, rank=config.rank, alpha=config.alpha, dropout=config.dropout)
        setattr(parent, leaf, wrapper)
        added = sum(p.numel() for p in wrapper.lora.parameters())
        n_params += added
        count += 1
        if verbose:
            print(f"Injected LoRA into {name}: {lin.in_features}x{lin.out_features} -> {added:,} trainable params")
    for module in model.modules():          # fused engines re-read their adapter layout
        hook = getattr(module, "on_lora_changed", None)
        if hook is not None:
            hook()
    if verbose:
        total = sum(p.numel() for p in model.parameters())
        trainable = sum(p.numel() for p in model.parameters() if p.requires_grad)
        print(f"\nTotal LoRA injections: {count}")
        print(f"Total LoRA parameters: {n_params:,}")
        print(f"Total model parameters: {total:,}")
        print(f"Trainable parameters: {trainable:,}")
        print(f"Trainable ratio: {100 * trainable / total:.2f}%")
    return model


def _wrappers(model: nn.Module):
    return [(n, m) for n, m in model.named_modules() if isinstance(m, LinearWithLoRA)]


def get_lora_parameters(model: nn.Module) -> List[nn.Parameter]:
    return [p for _, m in _wrappers(model) for p in m.lora.parameters()]


def get_lora_state_dict(model: nn.Module) -> Dict[str, torch.Tensor]:
    sd: Dict[str, torch.Tensor] = {}
    for name, m in _wrappers(model):
        sd[f"{name}.lora.lora_A"] = m.lora.lora_A.data
        sd[f"{name}.lora.lora_B"] = m.lora.lora_B.data
    return sd


def load_lora_state_dict(model: nn.Module, state_dict: Dict[str, torch.Tensor]):
    """Missing keys are skipped silently, as in the reference (:211-227); values are copied onto the parameter's device."""
    for name, m in _wrappers(model):
        for leaf, p in (("lora_A", m.lora.lora_A), ("lora_B", m.lora.lora_B)):
            t = state_dict.get(f"{name}.lora.{leaf}")
            if t is None:
                continue
            if tuple(t.shape) != tuple(p.shape):
                raise ValueError(f"{name}.lora.{leaf}: checkpoint shape {tuple(t.shape)} != parameter shape {tuple(p.shape)}")
            p.data = t.to(device=p.device, dtype=p.dtype).clone()


def merge_lora_weights(model: nn.Module) -> nn.Module:
    for name, m in _wrappers(model):
        parent_name, _, leaf = name.rpartition(".")
        parent = model.get_submodule(parent_name) if parent_name else model
        setattr(parent, leaf, m.merge_weights())
    return model


def print_trainable_parameters(model: nn.Module):
    trainable = sum(p.numel() for p in model.parameters() if p.requires_grad)
    total = sum(p.numel() for p in model.parameters())
    print(f"trainable params: {trainable:,} || all params: {total:,} || trainable%: {100 * trainable / total:.2f}")
