"""ctypes binding of libsam3b.so (C ABI declared in include/sam3b.h).

The product path has no CPU fallback: if the shared library is missing, or a CUDA entry point
is called without a GPU, the call raises.  PyTorch is used only to own device memory and
streams; every kernel behind these entry points is hand-written sm_100a CUDA.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "libsam3b.so"

F16, BF16 = 0, 1
EPI_STORE16, EPI_QKV_ROPE, EPI_RESIDUAL_F32, EPI_GELU, EPI_DGELU, EPI_ATOMIC_F32, EPI_STORE32, EPI_ADDMASK16 = range(8)


class Sam3bError(RuntimeError):
    pass


class GemmDesc(C.Structure):
    _fields_ = [
        ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
        ("A", C.c_void_p), ("lda", C.c_int64), ("a_mn", C.c_int32),
        ("B", C.c_void_p), ("ldb", C.c_int64), ("b_mn", C.c_int32),
        ("dtype", C.c_int32), ("epilogue", C.c_int32),
        ("C", C.c_void_p), ("ldc", C.c_int64),
        ("C2", C.c_void_p), ("ldc2", C.c_int64),
        ("bias", C.c_void_p),
        ("residual", C.c_void_p), ("ldres", C.c_int64), ("res_row_mod", C.c_int32),
        ("aux", C.c_void_p), ("ldaux", C.c_int64),
        ("rope", C.c_void_p), ("rope_period", C.c_int32), ("rope_cols", C.c_int32),
        ("alpha", C.c_float),
        ("splitk", C.c_int32), ("c_trans", C.c_int32), ("bn", C.c_int32),
        ("dbg_lbo", C.c_int32), ("dbg_sbo", C.c_int32), ("max_ctas", C.c_int32), ("cta_pair", C.c_int32),
        ("row_scale", C.c_void_p), ("rows_per_scale", C.c_int32),
        ("drop_p", C.c_float), ("drop_seed", C.c_uint32),
    ]


_lib = None


def load() -> C.CDLL:
    """Load libsam3b.so (built in-tree by __graft_entry__.build() / csrc/Makefile)."""
    global _lib
    if _lib is not None:
        return _lib
    path = Path(os.environ.get("SAM3B_LIB", LIB_PATH))
    if not path.exists():
        raise Sam3bError(
            f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback for the hot path)")
    lib = C.CDLL(str(path))
    lib.sam3b_last_error.restype = C.c_char_p
    lib.sam3b_abi_version.restype = C.c_int
    lib.sam3b_launch_count.restype = C.c_int64
    lib.sam3b_gemm.argtypes = [C.POINTER(GemmDesc), C.c_void_p]
    lib.sam3b_gemm.restype = C.c_int
    _declare_optional(lib)
    _lib = lib
    return lib


def _declare_optional(lib: C.CDLL) -> None:
    """Signatures of the remaining entry points (declared as they are added to sam3b.h)."""
    from . import _abi  # noqa: PLC0415  (keeps this file small)

    _abi.declare(lib)


def launch_count() -> int:
    return int(load().sam3b_launch_count())


def check(rc: int) -> None:
    if rc != 0:
        msg = load().sam3b_last_error()
        raise Sam3bError(f"libsam3b error {rc}: {msg.decode() if msg else '?'}")


def ptr(t) -> int | None:
    """Device pointer of a torch tensor (None passes NULL)."""
    return None if t is None else t.data_ptr()


def current_stream() -> int:
    import torch  # noqa: PLC0415

    return torch.cuda.current_stream().cuda_stream


def torch_dtype_code(dt) -> int:
    import torch  # noqa: PLC0415

    if dt == torch.float16:
        return F16
    if dt == torch.bfloat16:
        return BF16
    raise Sam3bError(f"unsupported operand dtype {dt}")


def gemm(A, B, C_out, *, epilogue=EPI_STORE16, a_mn=False, b_mn=False, M=None, N=None, K=None,
         bias=None, residual=None, res_row_mod=0, aux=None, rope=None, rope_period=1, rope_cols=0,
         C2=None, alpha=1.0, splitk=1, c_trans=False, bn=0, dbg_lbo=0, dbg_sbo=0, max_ctas=0, cta_pair=0, row_scale=None, rows_per_scale=1, drop_p=0.0, drop_seed=0):
    """C = epilogue(alpha * A @ B^T).  A:[M,K] (or [K,M] if a_mn), B:[N,K] (or [K,N] if b_mn).

    Tensors may be column-slices of wider buffers: leading dimensions are taken from stride(0).
    """
    lib = load()
    if M is None:
        M = A.shape[1] if a_mn else A.shape[0]
    if K is None:
        K = A.shape[0] if a_mn else A.shape[1]
    if N is None:
        N = B.shape[1] if b_mn else B.shape[0]
    d = GemmDesc()
    d.M, d.N, d.K = M, N, K
    d.A, d.lda, d.a_mn = ptr(A), A.stride(0), int(a_mn)
    d.B, d.ldb, d.b_mn = ptr(B), B.stride(0), int(b_mn)
    d.dtype = torch_dtype_code(A.dtype)
    d.epilogue = epilogue
    d.C, d.ldc = ptr(C_out), C_out.stride(0)
    if C2 is not None:
        d.C2, d.ldc2 = ptr(C2), C2.stride(0)
    d.bias = ptr(bias)
    if residual is not None:
        d.residual, d.ldres, d.res_row_mod = ptr(residual), residual.stride(0), res_row_mod
    if aux is not None:
        d.aux, d.ldaux = ptr(aux), aux.stride(0)
    if rope is not None:
        d.rope, d.rope_period, d.rope_cols = ptr(rope), rope_period, rope_cols
    d.alpha = alpha
    d.splitk, d.c_trans, d.bn = splitk, int(c_trans), bn
    d.dbg_lbo, d.dbg_sbo, d.max_ctas, d.cta_pair = dbg_lbo, dbg_sbo, max_ctas, cta_pair
    d.row_scale, d.rows_per_scale = ptr(row_scale), rows_per_scale
    d.drop_p, d.drop_seed = drop_p, drop_seed
    check(lib.sam3b_gemm(C.byref(d), current_stream()))
    return C_out


# ----------------------------------------------------------------------------------------------
# thin wrappers over the remaining entry points (tensors in, tensors out)
# ----------------------------------------------------------------------------------------------
def layernorm_fwd(x, gamma, beta, eps, y16, mean, rstd):
    """x: fp32 [rows, D] contiguous; y16: 16-bit [rows, >=D] (may be a slice of a wider buffer)."""
    rows, D = x.shape
    check(load().sam3b_layernorm_fwd(ptr(x), ptr(gamma), ptr(beta), eps, rows, D, ptr(y16), y16.stride(0),
                                     torch_dtype_code(y16.dtype), ptr(mean), ptr(rstd), current_stream()))


def layernorm_bwd(dy16, x, mean, rstd, gamma, dres, dx, dx16=None):
    rows, D = x.shape
    check(load().sam3b_layernorm_bwd(ptr(dy16), dy16.stride(0), ptr(x), ptr(mean), ptr(rstd), ptr(gamma), ptr(dres),
                                     rows, D, ptr(dx), ptr(dx16), 0 if dx16 is None else dx16.stride(0),
                                     torch_dtype_code(dy16.dtype), current_stream()))


def cast_rows_16(x, y16, scale=None):
    """y16[:, :D] = (16-bit)(x * *scale); scale: optional 1-element fp32 device tensor."""
    rows, D = x.shape
    check(load().sam3b_cast_rows_16_scaled(ptr(x), rows, D, ptr(y16), y16.stride(0), torch_dtype_code(y16.dtype), ptr(scale),
                                           current_stream()))


GRAD_SCALE_TARGET = 256.0
ALL_ROWS = 1 << 30   # rows_per_scale that maps every GEMM row to row_scale[0]


def grad_scale(g, target: float = GRAD_SCALE_TARGET):
    """Device-resident fp32 [s, 1/s, scratch, -] with s = 2^floor(log2(target / max|g|)) (1 for g == 0).  Every backward
    chain here is linear in the incoming gradient, so it runs on s*g (fp16 operands neither overflow nor flush a
    mean-reduced loss gradient to zero) and the last kernel multiplies by 1/s; nothing is read back to the host."""
    import torch  # noqa: PLC0415

    scale = torch.empty(4, device=g.device, dtype=torch.float32)
    check(load().sam3b_grad_scale(ptr(g), g.numel(), float(target), ptr(scale), current_stream()))
    return scale


def _attn_desc(qkv, seg_len, D, heads, O, lse2):
    """ViT-style descriptor: q | k | v are column blocks of one buffer, segments of seg_len rows."""
    from ._abi import AttnDesc  # noqa: PLC0415

    d = AttnDesc()
    d.q, d.ldq, d.q_cols, d.q_col0 = ptr(qkv), qkv.stride(0), 3 * D, 0
    d.kv, d.ldkv, d.kv_cols, d.k_col0, d.v_col0 = ptr(qkv), qkv.stride(0), 3 * D, D, 2 * D
    d.nseg, d.Lq, d.Lk, d.heads = qkv.shape[0] // seg_len, seg_len, seg_len, heads
    d.dtype = torch_dtype_code(qkv.dtype)
    d.scale = (D // heads) ** -0.5
    d.O, d.ldo, d.o_col0 = ptr(O), O.stride(0), 0
    d.lse2 = ptr(lse2)
    return d


def attention_fwd(qkv, seg_len, D, heads, O, lse2):
    d = _attn_desc(qkv, seg_len, D, heads, O, lse2)
    check(load().sam3b_attention_fwd(C.byref(d), current_stream()))


def attention_bwd(qkv, seg_len, D, heads, O, lse2, dO, delta, dqkv, rope, rope_period):
    d = _attn_desc(qkv, seg_len, D, heads, O, lse2)
    d.dO, d.lddo, d.do_col0 = ptr(dO), dO.stride(0), 0
    d.delta = ptr(delta)
    d.dq, d.lddq, d.dq_col0 = ptr(dqkv), dqkv.stride(0), 0
    d.dkv, d.lddkv, d.dk_col0, d.dv_col0 = ptr(dqkv), dqkv.stride(0), D, 2 * D
    d.rope, d.rope_period = ptr(rope), rope_period
    check(load().sam3b_attention_bwd(C.byref(d), current_stream()))


def mha_desc(q, kv, nseg, Lq, Lk, heads, scale, O, lse2, *, q_col0=0, k_col0=0, v_col0=None, bias=None, kpm=None,
             drop_p=0.0, drop_seed=0):
    """General descriptor (cross attention, padded 32-wide heads, attn_mask / key_padding_mask / dropout)."""
    from ._abi import AttnDesc  # noqa: PLC0415

    d = AttnDesc()
    d.q, d.ldq, d.q_cols, d.q_col0 = ptr(q), q.stride(0), q.shape[1], q_col0
    d.kv, d.ldkv, d.kv_cols, d.k_col0 = ptr(kv), kv.stride(0), kv.shape[1], k_col0
    d.v_col0 = heads * 64 if v_col0 is None else v_col0
    d.nseg, d.Lq, d.Lk, d.heads = nseg, Lq, Lk, heads
    d.dtype = torch_dtype_code(q.dtype)
    d.scale = scale
    d.O, d.ldo, d.o_col0 = ptr(O), O.stride(0), 0
    d.lse2 = ptr(lse2)
    d.bias, d.kpm, d.drop_p, d.drop_seed = ptr(bias), ptr(kpm), drop_p, drop_seed
    return d


# Above this many scores per (segment, head) the attention-dropout mask is generated once per call as keep-bits (both
# orientations) and read by the three kernels, instead of each of them hashing every score.
DROPOUT_BITS_MIN_SCORES = 1 << 16


def attention_dropout_bits(d):
    """Fills d.drop_bits / d.drop_bitsT when the call qualifies (dropout on, Lq and Lk multiples of 32, large problem);
    returns the two int32 tensors (keep them alive until the backward has run) or None."""
    import torch  # noqa: PLC0415

    if d.drop_p <= 0.0 or d.Lq % 32 or d.Lk % 32 or d.Lq * d.Lk < DROPOUT_BITS_MIN_SCORES:
        return None
    n_bh = d.nseg * d.heads
    dev = torch.device("cuda", torch.cuda.current_device())
    pitch = lambda n: (n // 32 + 7) // 8 * 8          # noqa: E731  (attn_bits_pitch: 32-byte rows)
    bits = torch.empty(n_bh * d.Lq * pitch(d.Lk), device=dev, dtype=torch.int32)
    bitsT = torch.empty(n_bh * d.Lk * pitch(d.Lq), device=dev, dtype=torch.int32)
    check(load().sam3b_attention_dropout_bits(n_bh, d.Lq, d.Lk, float(d.drop_p), int(d.drop_seed) & 0xFFFFFFFF, ptr(bits), ptr(bitsT),
                                              current_stream()))
    d.drop_bits, d.drop_bitsT = ptr(bits), ptr(bitsT)
    return bits, bitsT


def mha_fwd(d):
    check(load().sam3b_attention_fwd(C.byref(d), current_stream()))


def mha_bwd(d, dO, delta, dq, dkv, *, dk_col0=0, dv_col0=None):
    d.dO, d.lddo, d.do_col0 = ptr(dO), dO.stride(0), 0
    d.delta = ptr(delta)
    d.dq, d.lddq, d.dq_col0 = ptr(dq), dq.stride(0), 0
    d.dkv, d.lddkv, d.dk_col0 = ptr(dkv), dkv.stride(0), dk_col0
    d.dv_col0 = d.heads * 64 if dv_col0 is None else dv_col0
    check(load().sam3b_attention_bwd(C.byref(d), current_stream()))


def patch_gather(img, P, ws, out16, Kpad):
    B, Cc, H, W = img.shape
    check(load().sam3b_patch_gather(ptr(img), B, Cc, H, W, P, ws, ptr(out16), out16.stride(0), Kpad,
                                    torch_dtype_code(out16.dtype), current_stream()))


def tokens_to_nchw(x, B, G, ws, D, out):
    check(load().sam3b_tokens_to_nchw(ptr(x), B, G, ws, D, ptr(out), current_stream()))


def nchw_to_tokens(g, B, G, ws, D, dx, dx16):
    check(load().sam3b_nchw_to_tokens(ptr(g), B, G, ws, D, ptr(dx), ptr(dx16), 0 if dx16 is None else dx16.stride(0),
                                      0 if dx16 is None else torch_dtype_code(dx16.dtype), current_stream()))


def make_lora_site(in_features, out_total, r, rpad, adapters):
    """adapters: list of (out_off, out_len, A_tensor[in,r], B_tensor[r,out_len]) (1..3 entries)."""
    from ._abi import LoraSite  # noqa: PLC0415

    s = LoraSite()
    s.in_, s.out_total, s.n, s.r, s.rpad = in_features, out_total, len(adapters), r, rpad
    for i, (off, ln, A, B) in enumerate(adapters):
        s.out_off[i], s.out_len[i] = off, ln
        s.A[i] = ptr(A)
        s.B[i] = ptr(B)
    return s


def lora_pack(site, down_T, w_ext, up_pack, wt_ext, dtype):
    check(load().sam3b_lora_pack(C.byref(site), ptr(down_T), ptr(w_ext), 0 if w_ext is None else w_ext.stride(0),
                                 ptr(up_pack), ptr(wt_ext), 0 if wt_ext is None else wt_ext.stride(0),
                                 torch_dtype_code(dtype), current_stream()))


def lora_unpack_grads(site, dA_pack, dB_pack, dA_list, dB_list):
    n = site.n
    arrA = (C.c_void_p * 3)(*([ptr(t) for t in dA_list] + [None] * (3 - n)))
    arrB = (C.c_void_p * 3)(*([ptr(t) for t in dB_list] + [None] * (3 - n)))
    check(load().sam3b_lora_unpack_grads(C.byref(site), ptr(dA_pack), ptr(dB_pack), arrA, arrB, current_stream()))


def adamw_step(p, g, m, v, lr, beta1, beta2, eps, weight_decay, step, grad_scale=1.0):
    check(load().sam3b_adamw_step(ptr(p), ptr(g), ptr(m), ptr(v), p.numel(), lr, beta1, beta2, eps, weight_decay, step,
                                  grad_scale, current_stream()))


def dropout_rows16(x16, out16, p, seed):
    rows, cols = x16.shape
    check(load().sam3b_dropout_rows16(ptr(x16), x16.stride(0), rows, cols, ptr(out16), out16.stride(0), p, seed,
                                      torch_dtype_code(x16.dtype), current_stream()))
