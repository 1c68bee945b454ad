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
    """vit.ViT.grad_hook: averages the flat LoRA gradient over ranks.

    segments = 1: one collective over the whole buffer, issued right after the last backward kernel (it is exposed: nothing
    but the optimizer is left to overlap with).  segments = S > 1: the trunk runs its backward in S block ranges
    (engine.backward_segment); as soon as a range has written its slice of the flat gradient, that slice is all-reduced on a
    side stream while the next range computes, so only the last slice's collective (1/S of the bytes) is exposed.  Same
    result as DDP's bucketed overlap (sam3_lora/train/native_trainer.py:322-340): sum over ranks, divided by world."""

    def __init__(self, group=None, average: bool = True, segments: int = 1):
        self.group = group
        self.average = average
        self.segments = max(1, int(segments))
        self.calls = 0
        self.bytes = 0
        self._side = None

    def _active(self) -> bool:
        return dist.is_initialized() and dist.get_world_size(self.group) > 1

    def _reduce(self, t: torch.Tensor):
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        if self.average:
            t.mul_(1.0 / dist.get_world_size(self.group))
        self.calls += 1
        self.bytes += t.numel() * t.element_size()

    def __call__(self, flat_grad: torch.Tensor) -> torch.Tensor:
        if self._active():
            self._reduce(flat_grad)
        return flat_grad

    # ---- overlapped form (called by vit._TrunkFn.backward when segments > 1) ----
    def reduce_slice(self, flat_slice: torch.Tensor) -> None:
        """All-reduce one finished slice; on CUDA it runs on a side stream behind an event of the current stream."""
        if not self._active() or flat_slice.numel() == 0:
            return
        if not flat_slice.is_cuda:
            self._reduce(flat_slice)
            return
        if self._side is None:
            self._side = torch.cuda.Stream(device=flat_slice.device)
        ev = torch.cuda.Event()
        ev.record()
        with torch.cuda.stream(self._side):
            self._side.wait_event(ev)
            self._reduce(flat_slice)

    def finish(self) -> None:
        """The compute stream waits for every outstanding slice (call once after the last segment)."""
        if self._side is not None:
            torch.cuda.current_stream().wait_stream(self._side)


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
