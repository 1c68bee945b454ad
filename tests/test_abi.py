"""The C ABI: libsam3b.so loads without a GPU and exports every function include/sam3b.h declares;
the ctypes declarations cover the same set (no compute calls here)."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


def _declared():
    header = (ROOT / "include" / "sam3b.h").read_text()
    return sorted(set(re.findall(r"\b(sam3b_[a-z0-9_]+)\s*\(", header)))


def test_library_exports_every_declared_symbol():
    from sam3_lora_b200 import _lib

    lib = _lib.load()
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), n
    assert lib.sam3b_abi_version() == 3       # 2: backward segments, device dropout seed, polygons; 3: attention dropout bits
    assert lib.sam3b_launch_count() == 0


def test_ctypes_signatures_cover_the_header():
    from sam3_lora_b200 import _abi, _lib, engine

    lib = _lib.load()
    engine._declare(lib)
    covered = set(_abi.SIGNATURES) | {"sam3b_gemm", "sam3b_last_error", "sam3b_abi_version", "sam3b_launch_count"}
    covered |= {n for n in _declared() if n.startswith("sam3b_vit_")}
    assert set(_declared()) <= covered
    for n in _declared():
        assert getattr(lib, n).argtypes is not None or n in ("sam3b_last_error", "sam3b_abi_version", "sam3b_launch_count"), n


def test_struct_layouts_match_header_field_order():
    from sam3_lora_b200 import _lib
    from sam3_lora_b200._abi import AttnDesc, LoraSite, MatcherDesc
    from sam3_lora_b200.engine import LoraEntryC, VitConfigC

    header = (ROOT / "include" / "sam3b.h").read_text()

    def fields(struct):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (struct, struct), header, re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        out = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            names = decl.split(",")
            first = names[0].split()[-1]
            out.append(first.strip("*").split("[")[0])
            out += [n.strip().strip("*").split("[")[0] for n in names[1:]]
        return out

    def py(cls):
        return [n.rstrip("_") for n, *_ in cls._fields_]

    assert fields("sam3b_gemm_desc") == py(_lib.GemmDesc)
    assert fields("sam3b_attn_desc") == py(AttnDesc)
    assert fields("sam3b_lora_site") == py(LoraSite)
    assert fields("sam3b_vit_config") == py(VitConfigC)
    assert fields("sam3b_lora_entry") == py(LoraEntryC)
    assert fields("sam3b_matcher_desc") == py(MatcherDesc)


def test_neck_matcher_loss_and_input_paths_fail_loudly_without_gpu():
    """Rows a8 / f1 / f2 / f4 have no CPU fallback either: CPU tensors raise Sam3bError before any kernel is launched."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import numpy as np
    import torch.nn as nn

    from sam3_lora_b200 import _lib, conv_ops
    from sam3_lora_b200.data import GpuPreprocessor
    from sam3_lora_b200.losses import mask_losses, sigmoid_focal_loss
    from sam3_lora_b200.matcher import BinaryHungarianMatcherV2

    with pytest.raises(_lib.Sam3bError):
        conv_ops.conv1x1_forward(torch.zeros(1, 8, 4, 4), nn.Conv2d(8, 8, 1).requires_grad_(False))
    with pytest.raises(_lib.Sam3bError):
        conv_ops.mask_einsum(torch.zeros(1, 8, 8), torch.zeros(1, 8, 4, 4))
    with pytest.raises(_lib.Sam3bError):
        mask_losses(torch.zeros(1, 4, 4), torch.zeros(1, 8, 8), 1.0)
    with pytest.raises(_lib.Sam3bError):
        sigmoid_focal_loss(torch.zeros(2, 8), torch.zeros(2, 8), 1.0)
    with pytest.raises(_lib.Sam3bError):
        BinaryHungarianMatcherV2()({"pred_logits": torch.zeros(1, 4, 1), "pred_boxes": torch.rand(1, 4, 4)},
                                   {"boxes_padded": torch.rand(1, 2, 4), "num_boxes": torch.tensor([2])})
    with pytest.raises(_lib.Sam3bError):
        GpuPreprocessor(64, device="cpu").image(np.zeros((8, 8, 3), np.uint8))


def test_hot_path_fails_loudly_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from sam3_lora_b200 import _lib
    from sam3_lora_b200.lora_layers import LoRALinear
    from sam3_lora_b200.vit import ViT

    lin = LoRALinear(torch.nn.Linear(16, 16), rank=4, alpha=8)
    with pytest.raises(_lib.Sam3bError):
        lin(torch.zeros(2, 16))
    v = ViT(img_size=224, embed_dim=128, depth=1, num_heads=2, mlp_ratio=4.75, window_size=8, global_att_blocks=(0,),
            pretrain_img_size=112)
    with pytest.raises(_lib.Sam3bError):
        v(torch.zeros(1, 3, 224, 224))


def test_engine_host_object_needs_no_gpu():
    """Layout queries (sizes, flat LoRA offsets) are pure host code."""
    from sam3_lora_b200.engine import VitEngine, VitSpec

    eng = VitEngine(VitSpec(), lora_rank=16, lora_scaling=2.0, lora_targets=["q_proj", "k_proj", "v_proj", "out_proj", "fc1", "fc2"])
    # 32 blocks x (4 x (1024*16 + 16*1024) + (1024*16 + 16*4736) + (4736*16 + 16*1024)) adapter parameters
    assert eng.lora_numel == 32 * (4 * 2 * 1024 * 16 + 2 * (1024 + 4736) * 16)
    assert len(eng.entries) == 32 * 6
    e0 = eng.entries[0]
    assert (e0.block, e0.target, e0.in_features, e0.out_features, e0.rank, e0.a_off) == (0, "q_proj", 1024, 1024, 16, 0)
    assert e0.b_off == 1024 * 16
    wb = eng.weight_bytes()
    assert 1.7e9 < wb < 2.1e9          # W and W^T in 16-bit for 446 M frozen parameters
    ws8 = eng.workspace_bytes(8, True)
    assert 4.5e10 < ws8 < 7.5e10       # ~55 GB of saved activations at batch 8 (fits 180 GB HBM3e)
    assert eng.workspace_bytes(8, False) < 5e9


def test_driver_build_entry_point_passes():
    """The driver's "does it build" check: make (a no-op when up to date) + ABI version against the header + every
    declared symbol exported."""
    import __graft_entry__ as entry

    entry.build()
