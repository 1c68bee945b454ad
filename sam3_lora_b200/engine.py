"""Python handle on the native ViT trunk engine (sam3b_vit_* in include/sam3b.h).

PyTorch owns the device memory (two uint8 tensors: packed frozen weights, workspace); the C++
engine (csrc/engine.cpp) owns the kernel schedule.  No arithmetic happens in Python.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Dict, List, Sequence, Tuple

from . import _lib
from ._abi import i32, i64

LORA_BITS = {"q_proj": 1, "k_proj": 2, "v_proj": 4, "out_proj": 8, "fc1": 16, "fc2": 32}
BIT_NAMES = {v: k for k, v in LORA_BITS.items()}


class VitConfigC(C.Structure):
    _fields_ = [
        ("img_size", i32), ("patch_size", i32), ("in_chans", i32), ("embed_dim", i32), ("depth", i32),
        ("num_heads", i32), ("mlp_hidden", i32), ("window_size", i32),
        ("n_global", i32), ("global_blocks", i32 * 16),
        ("pos_side", i32), ("ln_eps", C.c_float), ("rope_theta", C.c_float),
        ("lora_rank", i32), ("lora_scaling", C.c_float), ("lora_targets", i32),
        ("dtype", i32), ("max_batch", i32),
    ]


class LoraEntryC(C.Structure):
    _fields_ = [("block", i32), ("target", i32), ("in_", i32), ("out", i32), ("rank", i32), ("a_off", i64), ("b_off", i64)]


@dataclass
class VitSpec:
    """Trunk hyper-parameters (defaults: sam3/model_builder.py:69-96)."""
    img_size: int = 1008
    patch_size: int = 14
    in_chans: int = 3
    embed_dim: int = 1024
    depth: int = 32
    num_heads: int = 16
    mlp_hidden: int = 4736
    window_size: int = 24
    global_blocks: Tuple[int, ...] = (7, 15, 23, 31)
    pretrain_img_size: int = 336
    ln_eps: float = 1e-5
    rope_theta: float = 10000.0

    @property
    def grid(self) -> int:
        return self.img_size // self.patch_size

    @property
    def pos_side(self) -> int:
        return self.pretrain_img_size // self.patch_size


@dataclass
class LoraEntry:
    block: int
    target: str
    in_features: int
    out_features: int
    rank: int
    a_off: int
    b_off: int


def _declare(lib):
    if getattr(lib, "_vit_declared", False):
        return
    P = C.POINTER
    lib.sam3b_vit_create.argtypes = [P(VitConfigC), P(C.c_void_p)]
    lib.sam3b_vit_create.restype = C.c_int
    lib.sam3b_vit_destroy.argtypes = [C.c_void_p]
    lib.sam3b_vit_destroy.restype = None
    for name in ("sam3b_vit_weight_bytes", "sam3b_vit_lora_numel"):
        getattr(lib, name).argtypes = [C.c_void_p]
        getattr(lib, name).restype = i64
    lib.sam3b_vit_workspace_bytes.argtypes = [C.c_void_p, i32, i32]
    lib.sam3b_vit_workspace_bytes.restype = i64
    lib.sam3b_vit_lora_count.argtypes = [C.c_void_p]
    lib.sam3b_vit_lora_count.restype = i32
    lib.sam3b_vit_lora_entry.argtypes = [C.c_void_p, i32, P(LoraEntryC)]
    lib.sam3b_vit_lora_entry.restype = C.c_int
    lib.sam3b_vit_bind.argtypes = [C.c_void_p, C.c_void_p, i64, C.c_void_p, i64, i32, i32]
    lib.sam3b_vit_bind.restype = C.c_int
    lib.sam3b_vit_load_base.argtypes = [C.c_void_p, P(C.c_void_p), i32, C.c_void_p]
    lib.sam3b_vit_load_base.restype = C.c_int
    lib.sam3b_vit_forward.argtypes = [C.c_void_p, C.c_void_p, i32, C.c_void_p, C.c_void_p, i32, C.c_void_p]
    lib.sam3b_vit_forward.restype = C.c_int
    lib.sam3b_vit_set_drop_path.argtypes = [C.c_void_p, C.c_void_p]
    lib.sam3b_vit_set_drop_path.restype = C.c_int
    lib.sam3b_vit_set_lora_dropout.argtypes = [C.c_void_p, C.c_float, C.c_uint32]
    lib.sam3b_vit_set_lora_dropout.restype = C.c_int
    lib.sam3b_vit_set_lora_dropout_dev.argtypes = [C.c_void_p, C.c_float, C.c_uint32, C.c_void_p]
    lib.sam3b_vit_set_lora_dropout_dev.restype = C.c_int
    lib.sam3b_vit_backward.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.sam3b_vit_backward.restype = C.c_int
    lib.sam3b_vit_backward_segment.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, i32, i32, C.c_void_p]
    lib.sam3b_vit_backward_segment.restype = C.c_int
    lib.sam3b_vit_lora_grad_range.argtypes = [C.c_void_p, i32, i32, P(i64), P(i64)]
    lib.sam3b_vit_lora_grad_range.restype = C.c_int
    lib._vit_declared = True


BASE_TENSOR_ORDER_HEAD = ("patch_embed.proj.weight", "pos_embed", "ln_pre.weight", "ln_pre.bias")
BASE_TENSOR_ORDER_BLOCK = ("norm1.weight", "norm1.bias", "attn.qkv.weight", "attn.qkv.bias", "attn.proj.weight",
                           "attn.proj.bias", "norm2.weight", "norm2.bias", "mlp.fc1.weight", "mlp.fc1.bias",
                           "mlp.fc2.weight", "mlp.fc2.bias")


def base_tensor_names(depth: int) -> List[str]:
    names = list(BASE_TENSOR_ORDER_HEAD)
    for i in range(depth):
        names += [f"blocks.{i}.{n}" for n in BASE_TENSOR_ORDER_BLOCK]
    return names


class VitEngine:
    """One engine per (spec, LoRA layout, operand dtype).  Device buffers are torch tensors."""

    def __init__(self, spec: VitSpec, *, lora_rank: int, lora_scaling: float, lora_targets: Sequence[str],
                 dtype="float16", max_batch: int = 8):
        import torch  # noqa: PLC0415

        self.lib = _lib.load()
        _declare(self.lib)
        self.spec = spec
        self.torch_dtype = getattr(torch, dtype) if isinstance(dtype, str) else dtype
        c = VitConfigC()
        c.img_size, c.patch_size, c.in_chans, c.embed_dim = spec.img_size, spec.patch_size, spec.in_chans, spec.embed_dim
        c.depth, c.num_heads, c.mlp_hidden, c.window_size = spec.depth, spec.num_heads, spec.mlp_hidden, spec.window_size
        c.n_global = len(spec.global_blocks)
        for i, g in enumerate(spec.global_blocks):
            c.global_blocks[i] = g
        c.pos_side, c.ln_eps, c.rope_theta = spec.pos_side, spec.ln_eps, spec.rope_theta
        bits = 0
        for t in lora_targets:
            if t not in LORA_BITS:
                raise _lib.Sam3bError(f"unknown LoRA target {t!r} (known: {sorted(LORA_BITS)})")
            bits |= LORA_BITS[t]
        c.lora_rank, c.lora_scaling, c.lora_targets = (lora_rank if bits else 0), float(lora_scaling), bits
        c.dtype, c.max_batch = _lib.torch_dtype_code(self.torch_dtype), max_batch
        h = C.c_void_p()
        _lib.check(self.lib.sam3b_vit_create(C.byref(c), C.byref(h)))
        self._h = h
        self.max_batch = max_batch
        self.lora_numel = int(self.lib.sam3b_vit_lora_numel(h))
        self.entries: List[LoraEntry] = []
        for i in range(self.lib.sam3b_vit_lora_count(h)):
            e = LoraEntryC()
            _lib.check(self.lib.sam3b_vit_lora_entry(h, i, C.byref(e)))
            self.entries.append(LoraEntry(e.block, BIT_NAMES[e.target], e.in_, e.out, e.rank, e.a_off, e.b_off))
        self.weight_buf = None
        self.work_buf = None
        self._bound = None

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                self.lib.sam3b_vit_destroy(h)
            except Exception:  # noqa: BLE001
                pass

    # ---- memory ------------------------------------------------------------------------------
    def weight_bytes(self) -> int:
        return int(self.lib.sam3b_vit_weight_bytes(self._h))

    def workspace_bytes(self, batch: int, training: bool) -> int:
        return int(self.lib.sam3b_vit_workspace_bytes(self._h, batch, int(training)))

    @staticmethod
    def _aligned(nbytes: int, device):
        import torch  # noqa: PLC0415

        raw = torch.empty(nbytes + 1024, dtype=torch.uint8, device=device)
        off = (-raw.data_ptr()) % 1024
        return raw[off:off + nbytes]

    def bind(self, device, batch: int, training: bool):
        if self._bound == (str(device), batch, training) and self.work_buf is not None:
            return
        if self.weight_buf is None or str(self.weight_buf.device) != str(device):
            self.weight_buf = self._aligned(self.weight_bytes(), device)
            self._loaded = False
        need = self.workspace_bytes(batch, training)
        if self.work_buf is None or self.work_buf.numel() < need or str(self.work_buf.device) != str(device):
            self.work_buf = None  # release before re-allocating
            self.work_buf = self._aligned(need, device)
        _lib.check(self.lib.sam3b_vit_bind(self._h, self.weight_buf.data_ptr(), self.weight_buf.numel(),
                                           self.work_buf.data_ptr(), self.work_buf.numel(), batch, int(training)))
        self._bound = (str(device), batch, training)

    def load_base(self, tensors: Dict[str, "object"]):
        """tensors: name -> fp32 CUDA tensor, names as in base_tensor_names()."""
        names = base_tensor_names(self.spec.depth)
        keep = []
        arr = (C.c_void_p * len(names))()
        for i, n in enumerate(names):
            t = tensors[n]
            if not t.is_cuda:
                raise _lib.Sam3bError(f"{n}: base weights must live on the GPU (no CPU fallback)")
            t = t.detach().float().contiguous()
            keep.append(t)
            arr[i] = t.data_ptr()
        _lib.check(self.lib.sam3b_vit_load_base(self._h, arr, len(names), _lib.current_stream()))
        import torch  # noqa: PLC0415

        torch.cuda.current_stream().synchronize()  # `keep` may be temporaries
        self._loaded = True

    # ---- compute -----------------------------------------------------------------------------
    def forward(self, img, lora_flat, out, save_for_backward: bool):
        _lib.check(self.lib.sam3b_vit_forward(self._h, img.data_ptr(), img.shape[0], _lib.ptr(lora_flat), out.data_ptr(),
                                              int(save_for_backward), _lib.current_stream()))

    def set_drop_path(self, scales):
        """scales: fp32 CUDA tensor [depth, 2, batch] (0 or 1/keep) or None; the caller keeps it alive until backward."""
        self._drop_scales = scales
        _lib.check(self.lib.sam3b_vit_set_drop_path(self._h, _lib.ptr(scales)))

    def set_lora_dropout(self, p: float, seed: int, seed_dev=None):
        """seed_dev: optional int32 CUDA tensor (1 element) whose value the kernels add to `seed` — rewritten per step by the
        caller so that CUDA-graph replays draw fresh masks; the caller keeps it alive."""
        self._drop_seed_dev = seed_dev
        if seed_dev is None:
            _lib.check(self.lib.sam3b_vit_set_lora_dropout(self._h, float(p), int(seed) & 0xFFFFFFFF))
        else:
            _lib.check(self.lib.sam3b_vit_set_lora_dropout_dev(self._h, float(p), int(seed) & 0xFFFFFFFF, seed_dev.data_ptr()))

    def backward(self, gout, grad_flat):
        _lib.check(self.lib.sam3b_vit_backward(self._h, gout.data_ptr(), _lib.ptr(grad_flat), _lib.current_stream()))

    def backward_segment(self, gout, grad_flat, block_hi: int, block_lo: int):
        """Blocks [block_hi .. block_lo] of the backward (consecutive calls from depth - 1 down to 0; only the first reads gout).
        On return (stream order) grad_flat[slice(*self.grad_range(block_hi, block_lo))] is final."""
        _lib.check(self.lib.sam3b_vit_backward_segment(self._h, _lib.ptr(gout), _lib.ptr(grad_flat), block_hi, block_lo,
                                                       _lib.current_stream()))

    def grad_range(self, block_hi: int, block_lo: int) -> Tuple[int, int]:
        lo, hi = i64(), i64()
        _lib.check(self.lib.sam3b_vit_lora_grad_range(self._h, block_hi, block_lo, C.byref(lo), C.byref(hi)))
        return int(lo.value), int(hi.value)

    def segments(self, n: int) -> List[Tuple[int, int]]:
        """n (or fewer) consecutive block ranges (hi, lo) covering depth - 1 .. 0, earlier ranges first."""
        depth = self.spec.depth
        n = max(1, min(n, depth))
        cuts = [depth - (depth * k) // n for k in range(n + 1)]          # depth, ..., 0
        return [(cuts[k] - 1, cuts[k + 1]) for k in range(n) if cuts[k] > cuts[k + 1]]
