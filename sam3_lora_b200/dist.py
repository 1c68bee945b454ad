"""Data-parallel plumbing for the hot path.

One process per GPU (torchrun / torch.distributed).  The batch shards over ranks; frozen weights are
replicated; the ONLY exchange per optimizer step is one sum-all-reduce of the flat LoRA gradient
buffer (<= 50 MB fp32), followed by the 1/world scaling of DDP
(reference DDP semantics: sam3_lora/train/native_trainer.py:322-340; sampler:
sam3/train/data/torch_dataset.py:31).  NCCL on GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

import os
from typing import List, Optional

import torch
import torch.distributed as dist


def init_from_env(backend: Optional[str] = None) -> tuple:
    """(rank, local_rank, world).  Initialises the default process group when WORLD_SIZE > 1."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29531")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend, rank=rank, world_size=world, **kw)
    return rank, local, world


class LoRAGradAllReducer:
    """vit.ViT.grad_hook: averages the flat LoRA gradient over ranks with a single collective, issued
    right after the last backward kernel so it overlaps the optimizer preparation."""

    def __init__(self, group=None, average: bool = True):
        self.group = group
        self.average = average
        self.calls = 0
        self.bytes = 0

    def __call__(self, flat_grad: torch.Tensor) -> torch.Tensor:
        if not dist.is_initialized():
            return flat_grad
        world = dist.get_world_size(self.group)
        if world == 1:
            return flat_grad
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=self.group)
        if self.average:
            flat_grad.mul_(1.0 / world)
        self.calls += 1
        self.bytes += flat_grad.numel() * flat_grad.element_size()
        return flat_grad


def shard_indices(n_items: int, rank: int, world: int, epoch: int = 0, shuffle: bool = True, seed: int = 0) -> List[int]:
    """DistributedSampler-style split: a common seeded permutation, padded by wrap-around so every rank
    gets ceil(n/world) items, rank r takes positions r, r+world, ..."""
    if shuffle:
        g = torch.Generator().manual_seed(seed + epoch)
        order = torch.randperm(n_items, generator=g).tolist()
    else:
        order = list(range(n_items))
    per = (n_items + world - 1) // world
    total = per * world
    order += order[: total - len(order)]
    return order[rank:total:world]
