"""argtypes/restype declarations for the libsam3b.so entry points beyond sam3b_gemm.

Kept next to include/sam3b.h on purpose: tests/test_abi.py checks that every function the
header declares is exported by the library and declared here (or in _lib.py).
"""
from __future__ import annotations

import ctypes as C

i32, i64, f32, vp = C.c_int32, C.c_int64, C.c_float, C.c_void_p


class AttnDesc(C.Structure):
    _fields_ = [
        ("q", vp), ("ldq", i64), ("q_cols", i32), ("q_col0", i32),
        ("kv", vp), ("ldkv", i64), ("kv_cols", i32), ("k_col0", i32), ("v_col0", i32),
        ("nseg", i32), ("Lq", i32), ("Lk", i32), ("heads", i32), ("dtype", i32),
        ("scale", f32),
        ("O", vp), ("ldo", i64), ("o_col0", i32),
        ("lse2", vp),
        ("bias", vp), ("kpm", vp), ("drop_p", f32), ("drop_seed", C.c_uint32),
        ("dO", vp), ("lddo", i64), ("do_col0", i32),
        ("delta", vp),
        ("dq", vp), ("lddq", i64), ("dq_col0", i32),
        ("dkv", vp), ("lddkv", i64), ("dk_col0", i32), ("dv_col0", i32),
        ("rope", vp), ("rope_period", i32),
        ("drop_bits", vp), ("drop_bitsT", vp),
    ]


class LoraSite(C.Structure):
    _fields_ = [
        ("in_", i32), ("out_total", i32), ("n", i32), ("r", i32), ("rpad", i32),
        ("out_off", i32 * 3), ("out_len", i32 * 3),
        ("A", vp * 3), ("B", vp * 3),
    ]


class MatcherDesc(C.Structure):
    _fields_ = [
        ("B", i32), ("Q", i32), ("Tmax", i32), ("repeats", i32),
        ("logits", vp), ("pred_boxes", vp), ("tgt_boxes", vp), ("num_boxes", vp), ("out_valid", vp), ("tgt_valid", vp),
        ("w_class", f32), ("w_bbox", f32), ("w_giou", f32),
        ("focal", i32), ("stable", i32),
        ("alpha", f32), ("gamma", f32),
    ]


# name -> (restype, argtypes)
SIGNATURES: dict[str, tuple] = {
    "sam3b_layernorm_fwd": (C.c_int, [vp, vp, vp, f32, i32, i32, vp, i64, i32, vp, vp, vp]),
    "sam3b_layernorm_bwd": (C.c_int, [vp, i64, vp, vp, vp, vp, vp, i32, i32, vp, vp, i64, i32, vp]),
    "sam3b_cast_rows_16": (C.c_int, [vp, i32, i32, vp, i64, i32, vp]),
    "sam3b_cast_rows_16_scaled": (C.c_int, [vp, i32, i32, vp, i64, i32, vp, vp]),
    "sam3b_attention_fwd": (C.c_int, [C.POINTER(AttnDesc), vp]),
    "sam3b_attention_bwd": (C.c_int, [C.POINTER(AttnDesc), vp]),
    "sam3b_attention_dropout_bits": (C.c_int, [i32, i32, i32, f32, C.c_uint32, vp, vp, vp]),
    "sam3b_patch_gather": (C.c_int, [vp, i32, i32, i32, i32, i32, i32, vp, i64, i32, i32, vp]),
    "sam3b_tokens_to_nchw": (C.c_int, [vp, i32, i32, i32, i32, vp, vp]),
    "sam3b_nchw_to_tokens": (C.c_int, [vp, i32, i32, i32, i32, vp, vp, i64, i32, vp]),
    "sam3b_lora_pack": (C.c_int, [C.POINTER(LoraSite), vp, vp, i64, vp, vp, i64, i32, vp]),
    "sam3b_lora_unpack_grads": (C.c_int, [C.POINTER(LoraSite), vp, vp, C.POINTER(vp), C.POINTER(vp), vp]),
    "sam3b_dropout_rows16": (C.c_int, [vp, i64, i32, i32, vp, i64, f32, C.c_uint32, i32, vp]),
    "sam3b_focal_loss_fwd": (C.c_int, [vp, vp, i64, f32, f32, vp, vp, vp]),
    "sam3b_focal_loss_bwd": (C.c_int, [vp, vp, i64, f32, f32, vp, f32, vp, vp]),
    "sam3b_mask_loss_fwd": (C.c_int, [vp, i32, i32, i32, vp, i32, i32, i32, f32, f32, f32, vp, vp, vp, vp]),
    "sam3b_mask_loss_bwd": (C.c_int, [vp, i32, i32, i32, vp, i32, i32, i32, f32, f32, f32, vp, vp, vp, vp]),
    "sam3b_matcher": (C.c_int, [C.POINTER(MatcherDesc), vp, vp, vp, vp]),
    "sam3b_resample_coeffs": (C.c_int, [i32, i32, vp, vp]),
    "sam3b_image_resize_normalize": (C.c_int, [vp, i32, i32, i32, vp, vp, i32, vp, vp, i32, vp, vp, f32, f32, vp]),
    "sam3b_rle_masks_nearest": (C.c_int, [vp, vp, vp, i32, i32, vp, vp]),
    "sam3b_poly_crossings": (C.c_int, [vp, vp, i32, i64, vp, vp]),
    "sam3b_poly_masks_nearest": (C.c_int, [vp, i64, vp, vp, i32, i32, vp, vp]),
    "sam3b_adamw_step": (C.c_int, [vp, vp, vp, vp, i64, f32, f32, f32, f32, f32, i32, f32, vp]),
    "sam3b_grad_scale": (C.c_int, [vp, i64, f32, vp, vp]),
    "sam3b_scale_cast": (C.c_int, [vp, i32, vp, i32, i64, i32, vp, i32, vp]),
    "sam3b_transpose_cast": (C.c_int, [vp, i32, vp, i32, i32, i32, i32, i32, vp, vp]),
    "sam3b_im2col3x3": (C.c_int, [vp, i32, i32, i32, i32, vp, i64, vp]),
    "sam3b_conv3x3_supported": (C.c_int, [i32, i32, i32, i32]),
    "sam3b_conv3x3": (C.c_int, [vp, i32, i32, i32, i32, vp, i32, vp, vp, i64, i32, i32, vp]),
    "sam3b_pixel_shuffle2": (C.c_int, [vp, i32, i32, i32, i32, i32, vp, i32, vp]),
    "sam3b_pixel_unshuffle2": (C.c_int, [vp, vp, i32, i32, i32, i32, vp, i32, vp]),
    "sam3b_maxpool2_fwd": (C.c_int, [vp, i32, i32, i32, i32, vp, i32, vp]),
    "sam3b_maxpool2_bwd": (C.c_int, [vp, vp, i32, i32, i32, i32, vp, vp, i32, vp]),
    "sam3b_upsample_add": (C.c_int, [vp, i32, i32, vp, i32, i32, i32, i32, vp, i32, vp]),
    "sam3b_upsample_add_bwd": (C.c_int, [vp, i32, i32, i32, i32, i32, i32, vp, i32, vp]),
    "sam3b_groupnorm_stats": (C.c_int, [vp, i32, i32, i32, i32, f32, vp, vp, vp]),
    "sam3b_groupnorm_relu_fwd": (C.c_int, [vp, vp, vp, vp, i32, i32, i32, i32, vp, i32, i32, vp]),
    "sam3b_groupnorm_relu_bwd": (C.c_int, [vp, vp, vp, vp, vp, i32, i32, i32, i32, vp, vp, i32, vp]),
}


def declare(lib: C.CDLL) -> None:
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = argtypes
