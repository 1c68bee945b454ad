"""Package-layout adapter classes: drop-in for `sam3_lora.lora.lora_layer` (sam3_lora/lora/lora_layer.py:16-178).

Unlike the root-level `lora_layers` module, this API stores the factors the PEFT way —
`lora_A [rank, in_features]` (kaiming-uniform, a = sqrt 5), `lora_B [out_features, rank]` (zeros) — applies
`x -> dropout(x) A^T B^T * (alpha / rank)` next to the frozen Linear, exposes `.linear` / `.lora` / `.weight` / `.bias`
on the wrapper and can fold the update back into a plain `nn.Linear` (`merge_weights`).  The arithmetic is the same fused
tcgen05 GEMM as `lora_layers.LoRALinear` (ops.lora_linear: adapter up-projection as a K-extension of the frozen
weight's K loop); the transposed views of the two small factors are all that differs.  CUDA only: CPU tensors raise.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn


class LoRALayer(nn.Module):
    """Adapter branch alone (lora_layer.py:16-88)."""

    def __init__(self, in_features: int, out_features: int, rank: int = 4, alpha: float = 1.0, dropout: float = 0.0):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.rank, self.alpha = rank, alpha
        self.scaling = alpha / rank
        self.lora_A = nn.Parameter(torch.empty(rank, in_features))
        self.lora_B = nn.Parameter(torch.empty(out_features, rank))
        self.dropout = nn.Dropout(p=dropout) if dropout > 0.0 else nn.Identity()
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.kaiming_uniform_(self.lora_A, a=math.sqrt(5))
        nn.init.zeros_(self.lora_B)

    @property
    def dropout_p(self) -> float:
        return self.dropout.p if isinstance(self.dropout, nn.Dropout) else 0.0

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        from ..ops import lora_branch  # noqa: PLC0415

        return lora_branch(self.dropout(x), self.lora_A.t(), self.lora_B.t(), self.scaling, owner=self)

    def merge_weights(self) -> torch.Tensor:
        """The dense update `B A * scaling`, shape [out_features, in_features] (init-time / export only: plain torch)."""
        return (self.lora_B @ self.lora_A) * self.scaling


class LinearWithLoRA(nn.Module):
    """Frozen `nn.Linear` + adapter (lora_layer.py:91-178); one fused GEMM per call on CUDA."""

    def __init__(self, linear: nn.Linear, rank: int = 4, alpha: float = 1.0, dropout: float = 0.0):
        super().__init__()
        self.linear = linear
        for p in self.linear.parameters():
            p.requires_grad = False
        self.in_features, self.out_features = linear.in_features, linear.out_features
        self.lora = LoRALayer(linear.in_features, linear.out_features, rank=rank, alpha=alpha, dropout=dropout)

    @property
    def weight(self):
        return self.linear.weight

    @property
    def bias(self):
        return self.linear.bias

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        from ..ops import lora_linear  # noqa: PLC0415

        p = self.lora.dropout_p if self.training else 0.0
        return lora_linear(x, self.linear.weight, self.linear.bias, self.lora.lora_A.t(), self.lora.lora_B.t(),
                           self.lora.scaling, dropout_p=p)

    def merge_weights(self) -> nn.Linear:
        lin = self.linear
        merged = nn.Linear(lin.in_features, lin.out_features, bias=lin.bias is not None, device=lin.weight.device,
                           dtype=lin.weight.dtype)
        with torch.no_grad():
            merged.weight.copy_(lin.weight + self.lora.merge_weights().to(lin.weight.dtype))
            if lin.bias is not None:
                merged.bias.copy_(lin.bias)
        return merged
