"""Neck / pixel decoder / mask head on the sm_100a kernels (row a8 of the hot-path table).

Reference semantics: `Sam3DualViTDetNeck.forward` (sam3/model/necks.py:100-125), `PixelDecoder.forward`
(sam3/model/maskformer_segmentation.py:203-219), `MaskPredictor.forward` (:28-51) and the 1x1 heads of
`UniversalSegmentationHead` (:270-273, 322-336).

How it runs here
  * activations are channels-last 16-bit between kernels; every convolution is a tcgen05 GEMM (`_lib.gemm`):
    ConvTranspose2d(2,2) = GEMM + pixel shuffle, Conv 1x1 = GEMM, Conv 3x3 = 16-bit im2col + GEMM, the mask einsum =
    one GEMM per image.  GroupNorm statistics, module inputs/outputs and all returned gradients are fp32.
  * each module is ONE autograd.Function with an explicit forward and backward schedule (like the trunk engine): the
    convolution weights are frozen under `apply_lora_to_model` (lora_layers.py:181-183), so only data gradients exist
    and nothing but the GELU / GroupNorm / max-pool inputs is saved.
  * a backward chain runs on gradients multiplied by a power-of-two scale derived on the device from max|grad|
    (`grad_scale`), so fp16 operands neither overflow nor flush to zero; the last kernel multiplies by 1/scale.
  * module outputs are fp32 tensors of logical shape [B, C, H, W] stored channels-last (a permuted view of the NHWC
    buffer the last GEMM wrote); channels-last inputs are consumed without a transpose.
There is no CPU path: CPU tensors raise Sam3bError.
"""
from __future__ import annotations

import weakref
from typing import List, Optional, Sequence

import torch

from . import _lib as L
from . import ops

_BIG = 1 << 30   # rows_per_scale that maps every row to scale element 0


def _dt():
    return ops._OPERAND_DTYPE


def _code(t: torch.Tensor) -> int:
    return 1 if t.dtype == torch.float32 else 0


def _dtc(t: Optional[torch.Tensor] = None) -> int:
    return L.torch_dtype_code(_dt() if t is None or t.dtype == torch.float32 else t.dtype)


def _st() -> int:
    return L.current_stream()


# --------------------------------------------------------------------------------------------------------------
# thin kernel wrappers (tensors in, tensors out)
# --------------------------------------------------------------------------------------------------------------
def grad_scale(g: torch.Tensor, target: float = L.GRAD_SCALE_TARGET) -> torch.Tensor:
    """See _lib.grad_scale.  g: any dense fp32 tensor."""
    return L.grad_scale(g, target)


def scale_cast(src: torch.Tensor, dst: torch.Tensor, scale: Optional[torch.Tensor] = None, accumulate: bool = False):
    n = src.numel()
    if n % 4:
        raise L.Sam3bError(f"scale_cast: {n} elements (must be a multiple of 4)")
    L.check(L.load().sam3b_scale_cast(L.ptr(src), _code(src), L.ptr(dst), _code(dst), n,
                                      L.torch_dtype_code(src.dtype if _code(src) == 0 else (dst.dtype if _code(dst) == 0 else _dt())),
                                      L.ptr(scale), int(accumulate), _st()))
    return dst


def transpose_cast(src: torch.Tensor, dst: torch.Tensor, batch: int, R: int, Cc: int, scale: Optional[torch.Tensor] = None):
    """src [batch, R, Cc] -> dst [batch, Cc, R] (both dense)."""
    dt16 = src.dtype if _code(src) == 0 else (dst.dtype if _code(dst) == 0 else _dt())
    L.check(L.load().sam3b_transpose_cast(L.ptr(src), _code(src), L.ptr(dst), _code(dst), batch, R, Cc,
                                          L.torch_dtype_code(dt16), L.ptr(scale), _st()))
    return dst


def _dense_any(x: torch.Tensor) -> torch.Tensor:
    """x as a dense tensor in either NCHW or channels-last order (no copy when it already is)."""
    if x.dtype != torch.float32:
        x = x.float()
    if x.is_contiguous() or (x.dim() == 4 and x.permute(0, 2, 3, 1).is_contiguous()):
        return x
    return x.contiguous()


def to_nhwc16(x: torch.Tensor, scale: Optional[torch.Tensor] = None) -> torch.Tensor:
    """[B, C, H, W] fp32 (NCHW or channels-last storage) -> [B, H, W, C] 16-bit, times *scale."""
    if not x.is_cuda:
        raise L.Sam3bError(f"neck / mask-head input is on {x.device}: this path has no CPU fallback")
    x = _dense_any(x)
    B, Cc, H, W = x.shape
    out = torch.empty(B, H, W, Cc, device=x.device, dtype=_dt())
    if x.permute(0, 2, 3, 1).is_contiguous():
        scale_cast(x, out, scale)
    else:
        transpose_cast(x, out, B, Cc, H * W, scale)
    return out


def nhwc32_as_nchw(buf: torch.Tensor) -> torch.Tensor:
    """[B, H, W, C] buffer -> logical [B, C, H, W] view (channels-last storage)."""
    return buf.permute(0, 3, 1, 2)


def im2col3x3(x16: torch.Tensor) -> torch.Tensor:
    B, H, W, Cc = x16.shape
    col = torch.empty(B * H * W, 9 * Cc, device=x16.device, dtype=x16.dtype)
    L.check(L.load().sam3b_im2col3x3(L.ptr(x16), B, H, W, Cc, L.ptr(col), col.stride(0), _st()))
    return col


def pixel_shuffle2(u16: torch.Tensor, B: int, H: int, W: int, Cc: int, gelu: bool = False) -> torch.Tensor:
    out = torch.empty(B, 2 * H, 2 * W, Cc, device=u16.device, dtype=u16.dtype)
    L.check(L.load().sam3b_pixel_shuffle2(L.ptr(u16), B, H, W, Cc, int(gelu), L.ptr(out), _dtc(u16), _st()))
    return out


def pixel_unshuffle2(dy16: torch.Tensor, h16: Optional[torch.Tensor] = None) -> torch.Tensor:
    B, H2, W2, Cc = dy16.shape
    out = torch.empty(B * (H2 // 2) * (W2 // 2), 4 * Cc, device=dy16.device, dtype=dy16.dtype)
    L.check(L.load().sam3b_pixel_unshuffle2(L.ptr(dy16), L.ptr(h16), B, H2 // 2, W2 // 2, Cc, L.ptr(out), _dtc(dy16), _st()))
    return out


def maxpool2_fwd(x16: torch.Tensor) -> torch.Tensor:
    B, H, W, Cc = x16.shape
    y = torch.empty(B, H // 2, W // 2, Cc, device=x16.device, dtype=x16.dtype)
    L.check(L.load().sam3b_maxpool2_fwd(L.ptr(x16), B, H, W, Cc, L.ptr(y), _dtc(x16), _st()))
    return y


def maxpool2_bwd(x16: torch.Tensor, dy16: torch.Tensor, inv_scale: Optional[torch.Tensor], dx32: torch.Tensor):
    B, H, W, Cc = x16.shape
    L.check(L.load().sam3b_maxpool2_bwd(L.ptr(x16), L.ptr(dy16), B, H, W, Cc, L.ptr(inv_scale), L.ptr(dx32), _dtc(x16), _st()))


def upsample_add(prev16: torch.Tensor, cur16: torch.Tensor) -> torch.Tensor:
    B, H, W, Cc = cur16.shape
    out = torch.empty_like(cur16)
    L.check(L.load().sam3b_upsample_add(L.ptr(prev16), prev16.shape[1], prev16.shape[2], L.ptr(cur16), B, H, W, Cc, L.ptr(out),
                                        _dtc(cur16), _st()))
    return out


def upsample_add_bwd(dout16: torch.Tensor, h: int, w: int) -> torch.Tensor:
    B, H, W, Cc = dout16.shape
    dprev = torch.empty(B, h, w, Cc, device=dout16.device, dtype=dout16.dtype)
    L.check(L.load().sam3b_upsample_add_bwd(L.ptr(dout16), B, H, W, Cc, h, w, L.ptr(dprev), _dtc(dout16), _st()))
    return dprev


def groupnorm_stats(x32: torch.Tensor, B: int, HW: int, Cc: int, G: int, eps: float) -> torch.Tensor:
    work = torch.empty(3 * B * G, device=x32.device, dtype=torch.float64)
    stat = torch.empty(B, G, 2, device=x32.device, dtype=torch.float32)
    L.check(L.load().sam3b_groupnorm_stats(L.ptr(x32), B, HW, Cc, G, float(eps), L.ptr(work), L.ptr(stat), _st()))
    return stat


def groupnorm_relu_fwd(x32, stat, gamma, beta, B, HW, Cc, G, out: torch.Tensor):
    L.check(L.load().sam3b_groupnorm_relu_fwd(L.ptr(x32), L.ptr(stat), L.ptr(gamma), L.ptr(beta), B, HW, Cc, G, L.ptr(out),
                                              _code(out), _dtc(out), _st()))
    return out


def groupnorm_relu_bwd(dy16, x32, stat, gamma, beta, B, HW, Cc, G) -> torch.Tensor:
    work = torch.empty(3 * B * G, device=x32.device, dtype=torch.float64)
    dx = torch.empty(B * HW, Cc, device=x32.device, dtype=dy16.dtype)
    L.check(L.load().sam3b_groupnorm_relu_bwd(L.ptr(dy16), L.ptr(x32), L.ptr(stat), L.ptr(gamma), L.ptr(beta), B, HW, Cc, G,
                                              L.ptr(work), L.ptr(dx), _dtc(dy16), _st()))
    return dx


IMPLICIT_CONV = True   # tests flip this to compare the TMA implicit-GEMM convolution with the im2col + GEMM path


def conv3x3(x16: torch.Tensor, w9: torch.Tensor, bias=None, out_f32: bool = False) -> torch.Tensor:
    """nn.Conv2d(C, Cout, 3, padding=1) on channels-last 16-bit x16 [B,H,W,C] -> [B*H*W, Cout] (16-bit or fp32).
    w9 [Cout, 9C] with k = (ky, kx, c).  Implicit GEMM (4-D TMA boxes, no patch matrix) when the shape allows it."""
    B, H, W, Cc = x16.shape
    co = w9.shape[0]
    lib = L.load()
    if IMPLICIT_CONV and lib.sam3b_conv3x3_supported(H, W, Cc, co):
        out = torch.empty(B * H * W, co, device=x16.device, dtype=torch.float32 if out_f32 else x16.dtype)
        L.check(lib.sam3b_conv3x3(L.ptr(x16), B, H, W, Cc, L.ptr(w9), co, L.ptr(bias), L.ptr(out), out.stride(0), int(out_f32),
                                  _dtc(x16), _st()))
        return out
    col = im2col3x3(x16)
    return _gemm32(col, w9, bias) if out_f32 else _gemm16(col, w9, bias)


def _gemm16(a16, w16, bias=None):
    out = torch.empty(a16.shape[0], w16.shape[0], device=a16.device, dtype=a16.dtype)
    return L.gemm(a16, w16, out, epilogue=L.EPI_STORE16, bias=bias)


def _gemm32(a16, w16, bias=None, inv_scale=None, out=None):
    if out is None:
        out = torch.empty(a16.shape[0], w16.shape[0], device=a16.device, dtype=torch.float32)
    return L.gemm(a16, w16, out, epilogue=L.EPI_STORE32, bias=bias, row_scale=inv_scale, rows_per_scale=_BIG)


def _gemm_acc32(a16, w16, acc32, inv_scale):
    """acc32 += inv_scale * a16 @ w16^T  (fp32 residual epilogue, in place)."""
    return L.gemm(a16, w16, acc32, epilogue=L.EPI_RESIDUAL_F32, residual=acc32, row_scale=inv_scale, rows_per_scale=_BIG, bn=256)


# --------------------------------------------------------------------------------------------------------------
# frozen-weight packing (cached per parameter tensor; re-packed when the tensor is modified in place or moved)
# --------------------------------------------------------------------------------------------------------------
class _PackCache:
    _cache: dict = {}

    @classmethod
    def get(cls, tensors: Sequence[torch.Tensor], kind: str, builder):
        ident = (kind,) + tuple(id(t) for t in tensors)
        key = tuple((t._version, t.data_ptr(), tuple(t.shape)) for t in tensors) + (_dt(),)
        hit = cls._cache.get(ident)
        if hit is not None and hit[0] == key:
            return hit[1]
        for t in tensors:
            if t.requires_grad:
                raise L.Sam3bError("neck / mask-head convolutions are frozen on this path (apply_lora_to_model freezes them, "
                                   "lora_layers.py:181-183); a trainable conv weight has no gradient kernel here")
        val = builder()
        if hit is None:
            weakref.finalize(tensors[0], cls._cache.pop, ident, None)
        cls._cache[ident] = (key, val)
        return val


def _pad8(n: int) -> int:
    return (n + 7) // 8 * 8


def pack_conv1x1(conv: torch.nn.Conv2d):
    """(W [Np, Cin], W^T [Cin, Np], bias [Np] fp32) with Np = out channels padded to 8 (zero rows)."""
    W, b = conv.weight, conv.bias

    def build():
        co, ci = W.shape[0], W.shape[1]
        npad = _pad8(co)
        w = torch.zeros(npad, ci, device=W.device, dtype=_dt())
        w[:co] = W.detach().reshape(co, ci)
        bias = torch.zeros(npad, device=W.device, dtype=torch.float32)
        if b is not None:
            bias[:co] = b.detach().float()
        return w, w.t().contiguous(), bias
    return _PackCache.get([W] + ([b] if b is not None else []), "c1", build)


def pack_conv3x3(conv: torch.nn.Conv2d):
    """(W9 [Cout, 9 Cin] with k = (ky, kx, ci);  Wg [Cin, 9 Cout] = flipped/transposed taps for the data gradient; bias)."""
    W, b = conv.weight, conv.bias

    def build():
        co, ci = W.shape[0], W.shape[1]
        w9 = W.detach().permute(0, 2, 3, 1).reshape(co, 9 * ci).to(_dt()).contiguous()
        wg = W.detach().flip(2, 3).permute(1, 2, 3, 0).reshape(ci, 9 * co).to(_dt()).contiguous()
        bias = b.detach().float().contiguous() if b is not None else None
        return w9, wg, bias
    return _PackCache.get([W] + ([b] if b is not None else []), "c3", build)


def pack_deconv2x2(conv: torch.nn.ConvTranspose2d):
    """(Wd [(di,dj,co), ci];  WdT [ci, (di,dj,co)];  bias repeated over (di,dj))."""
    W, b = conv.weight, conv.bias

    def build():
        ci, co = W.shape[0], W.shape[1]
        wd = W.detach().permute(2, 3, 1, 0).reshape(4 * co, ci).to(_dt()).contiguous()
        wdt = W.detach().permute(0, 2, 3, 1).reshape(ci, 4 * co).to(_dt()).contiguous()
        bias = b.detach().float().repeat(4).contiguous() if b is not None else None
        return wd, wdt, bias
    return _PackCache.get([W] + ([b] if b is not None else []), "d2", build)


# --------------------------------------------------------------------------------------------------------------
# SimpleFPN neck (necks.py:40-125)
# --------------------------------------------------------------------------------------------------------------
class _NeckFn(torch.autograd.Function):
    """x [B, C, H, W] -> one [B, d_model, s*H, s*W] map per branch; `branches` = list of nn.Sequential as the reference builds."""

    @staticmethod
    def forward(ctx, x, branches):
        ctx.set_materialize_grads(False)      # unused pyramid levels arrive as None and cost nothing in backward
        B, Cc, H, W = x.shape
        x16 = to_nhwc16(x)
        M = B * H * W
        outs, saved = [], []
        for seq in branches:
            names = [n for n, _ in seq.named_children()]
            t, h, w, c = x16, H, W, Cc
            h0 = None
            if "dconv_2x2_0" in names:        # scale 4: deconv -> GELU -> deconv   (necks.py:42-56)
                wd, _, bias = pack_deconv2x2(seq.dconv_2x2_0)
                h0 = _gemm16(t.view(M, c), wd, bias)
                c = wd.shape[0] // 4
                t = pixel_shuffle2(h0, B, h, w, c, gelu=True)
                h, w = 2 * h, 2 * w
                wd, _, bias = pack_deconv2x2(seq.dconv_2x2_1)
                u = _gemm16(t.view(-1, c), wd, bias)
                c = wd.shape[0] // 4
                t = pixel_shuffle2(u, B, h, w, c)
                h, w = 2 * h, 2 * w
            elif "dconv_2x2" in names:        # scale 2   (necks.py:57-62)
                wd, _, bias = pack_deconv2x2(seq.dconv_2x2)
                u = _gemm16(t.view(M, c), wd, bias)
                c = wd.shape[0] // 4
                t = pixel_shuffle2(u, B, h, w, c)
                h, w = 2 * h, 2 * w
            elif "maxpool_2x2" in names:      # scale 0.5 (necks.py:65-70)
                t = maxpool2_fwd(t)
                h, w = h // 2, w // 2
            w1, _, b1 = pack_conv1x1(seq.conv_1x1)
            d = seq.conv_1x1.out_channels
            c1 = _gemm16(t.view(-1, c), w1, b1)
            if w1.shape[0] != d:
                c1 = c1[:, :d].contiguous()
            w9, _, b9 = pack_conv3x3(seq.conv_3x3)
            out = conv3x3(c1.view(B, h, w, d), w9, b9, out_f32=True)
            outs.append(nhwc32_as_nchw(out.view(B, h, w, seq.conv_3x3.out_channels)))
            saved.append(h0)
        ctx.branches = branches
        ctx.dims = (B, Cc, H, W)
        ctx.n_h0 = [s is not None for s in saved]
        ctx.save_for_backward(x16, *[s for s in saved if s is not None])
        return tuple(outs)

    @staticmethod
    def backward(ctx, *gouts):
        B, Cc, H, W = ctx.dims
        x16, *h0s = ctx.saved_tensors
        h0s = list(h0s)
        M = B * H * W
        dx32 = torch.zeros(B, H, W, Cc, device=x16.device, dtype=torch.float32)
        for seq, g, has_h0 in zip(ctx.branches, gouts, ctx.n_h0):
            h0 = h0s.pop(0) if has_h0 else None
            if g is None:
                continue
            names = [n for n, _ in seq.named_children()]
            g = _dense_any(g)
            sc = grad_scale(g)
            s_in, s_out = sc[0:1], sc[1:2]
            dy16 = to_nhwc16(g, s_in)
            _, wg, _ = pack_conv3x3(seq.conv_3x3)
            dc1 = conv3x3(dy16, wg)                                   # [Mb, d_model]
            w1, w1t, _ = pack_conv1x1(seq.conv_1x1)
            if w1.shape[0] != dc1.shape[1]:
                dc1 = torch.nn.functional.pad(dc1, (0, w1.shape[0] - dc1.shape[1]))
            if "dconv_2x2_0" in names:
                dt = _gemm16(dc1, w1t)                                # [M2, C/4] at 4H x 4W
                c2 = dt.shape[1]
                _, wdt1, _ = pack_deconv2x2(seq.dconv_2x2_1)
                dg0 = _gemm16(pixel_unshuffle2(dt.view(B, 4 * H, 4 * W, c2)), wdt1)      # [M1, C/2] at 2H x 2W
                c1c = dg0.shape[1]
                _, wdt0, _ = pack_deconv2x2(seq.dconv_2x2_0)
                _gemm_acc32(pixel_unshuffle2(dg0.view(B, 2 * H, 2 * W, c1c), h0), wdt0, dx32.view(M, Cc), s_out)
            elif "dconv_2x2" in names:
                dt = _gemm16(dc1, w1t)
                _, wdt, _ = pack_deconv2x2(seq.dconv_2x2)
                _gemm_acc32(pixel_unshuffle2(dt.view(B, 2 * H, 2 * W, dt.shape[1])), wdt, dx32.view(M, Cc), s_out)
            elif "maxpool_2x2" in names:
                dt = _gemm16(dc1, w1t)
                maxpool2_bwd(x16, dt.view(B, H // 2, W // 2, Cc), s_out, dx32)
            else:
                _gemm_acc32(dc1, w1t, dx32.view(M, Cc), s_out)
        dx = torch.empty(B, Cc, H, W, device=x16.device, dtype=torch.float32)
        transpose_cast(dx32, dx, B, H * W, Cc)
        return dx, None


def neck_forward(x: torch.Tensor, branches) -> List[torch.Tensor]:
    return list(_NeckFn.apply(x, list(branches)))


# --------------------------------------------------------------------------------------------------------------
# pixel decoder (maskformer_segmentation.py:172-219)
# --------------------------------------------------------------------------------------------------------------
class _PixelDecoderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, convs, norms, shared_conv, *feats):
        prev = to_nhwc16(feats[-1])
        fpn = feats[:-1][::-1]
        B = prev.shape[0]
        ys, stats, shapes = [], [], [tuple(prev.shape)]
        out = None
        for li, f in enumerate(fpn):
            k = 0 if shared_conv else li
            cur = to_nhwc16(f)
            _, H, W, Cc = cur.shape
            s16 = upsample_add(prev, cur)
            w9, _, b9 = pack_conv3x3(convs[k])
            y32 = conv3x3(s16, w9, b9, out_f32=True)                                        # [B*H*W, C] fp32
            G = norms[k].num_groups
            stat = groupnorm_stats(y32, B, H * W, Cc, G, norms[k].eps)
            last = li == len(fpn) - 1
            o = torch.empty(B, H, W, Cc, device=y32.device, dtype=torch.float32 if last else prev.dtype)
            groupnorm_relu_fwd(y32, stat, norms[k].weight.detach().float().contiguous(), norms[k].bias.detach().float().contiguous(),
                               B, H * W, Cc, G, o)
            ys.append(y32); stats.append(stat); shapes.append((B, H, W, Cc))
            prev, out = o, o
        ctx.cfg = (convs, norms, shared_conv, shapes)
        ctx.save_for_backward(*ys, *stats)
        return nhwc32_as_nchw(out)

    @staticmethod
    def backward(ctx, g):
        convs, norms, shared_conv, shapes = ctx.cfg
        n = len(shapes) - 1
        saved = ctx.saved_tensors
        ys, stats = saved[:n], saved[n:]
        g = _dense_any(g)
        sc = grad_scale(g)
        s_in, s_out = sc[0:1], sc[1:2]
        dy16 = to_nhwc16(g, s_in)
        grads_fpn = []
        for li in range(n - 1, -1, -1):
            k = 0 if shared_conv else li
            B, H, W, Cc = shapes[li + 1]
            G = norms[k].num_groups
            dconv = groupnorm_relu_bwd(dy16.view(B * H * W, Cc), ys[li], stats[li], norms[k].weight.detach().float().contiguous(),
                                       norms[k].bias.detach().float().contiguous(), B, H * W, Cc, G)
            _, wg, _ = pack_conv3x3(convs[k])
            ds16 = conv3x3(dconv.view(B, H, W, Cc), wg)                                      # d(curr + up(prev)), scaled
            dcur = torch.empty(B, H, W, Cc, device=ds16.device, dtype=torch.float32)
            scale_cast(ds16, dcur, s_out)
            grads_fpn.append(nhwc32_as_nchw(dcur))
            _, h, w, _ = shapes[li]
            dy16 = upsample_add_bwd(ds16.view(B, H, W, Cc), h, w)
        dlast = torch.empty(*shapes[0], device=dy16.device, dtype=torch.float32)
        scale_cast(dy16, dlast, s_out)
        # feats order: fpn_feats (= reversed processing order) then the coarsest level
        return (None, None, None, *grads_fpn, nhwc32_as_nchw(dlast))


def pixel_decoder_forward(feats: Sequence[torch.Tensor], convs, norms, shared_conv: bool) -> torch.Tensor:
    if len(feats) == 1:   # a single level: the reference's loop body never runs and it returns backbone_feats[-1]
        return feats[-1]
    # the reference relies on broadcasting when the backbone maps have batch 1 and the encoder map one entry per query
    # (maskformer_segmentation.py:118-120, 209): make the batch explicit (autograd sums the expanded gradient back)
    bmax = max(f.shape[0] for f in feats)
    feats = [f.expand(bmax, -1, -1, -1) if (f.shape[0] == 1 and bmax > 1) else f for f in feats]
    return _PixelDecoderFn.apply(list(convs), list(norms), bool(shared_conv), *feats)


# --------------------------------------------------------------------------------------------------------------
# 1x1 heads and the mask einsum (maskformer_segmentation.py:23-51, 270-273)
# --------------------------------------------------------------------------------------------------------------
class _Conv1x1Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, conv):
        w, _, bias = pack_conv1x1(conv)
        x16 = to_nhwc16(x)
        B, H, W, Cc = x16.shape
        out = _gemm32(x16.view(-1, Cc), w, bias)
        ctx.conv = conv
        ctx.dims = (B, H, W, Cc)
        co = conv.out_channels
        return nhwc32_as_nchw(out.view(B, H, W, -1)[..., :co])

    @staticmethod
    def backward(ctx, g):
        B, H, W, Cc = ctx.dims
        w, wt, _ = pack_conv1x1(ctx.conv)
        g = _dense_any(g)
        sc = grad_scale(g)
        dy16 = to_nhwc16(g, sc[0:1]).view(B * H * W, -1)
        if dy16.shape[1] != w.shape[0]:
            dy16 = torch.nn.functional.pad(dy16, (0, w.shape[0] - dy16.shape[1]))
        dx = _gemm32(dy16, wt, None, sc[1:2])
        return nhwc32_as_nchw(dx.view(B, H, W, Cc)), None


def conv1x1_forward(x: torch.Tensor, conv: torch.nn.Conv2d) -> torch.Tensor:
    return _Conv1x1Fn.apply(x, conv)


class _MaskEinsumFn(torch.autograd.Function):
    """einsum("bqc,bchw->bqhw"): one [Q, C] x [HW, C]^T GEMM per image, fp32 logits out."""

    @staticmethod
    def forward(ctx, me, pix):
        if not me.is_cuda:
            raise L.Sam3bError(f"mask einsum operand is on {me.device}: this path has no CPU fallback")
        B, Q, Cc = me.shape
        inst16 = to_nhwc16(pix)                                    # [B, H, W, C]
        _, H, W, _ = inst16.shape
        if Cc % 8 or (H * W) % 8:
            raise L.Sam3bError(f"mask einsum: C={Cc} and H*W={H * W} must be multiples of 8")
        me16 = torch.empty(B, Q, Cc, device=me.device, dtype=_dt())
        scale_cast(me.detach().float().contiguous(), me16)
        out = torch.empty(B, Q, H * W, device=me.device, dtype=torch.float32)
        i2 = inst16.view(B, H * W, Cc)
        for b in range(B):
            L.gemm(me16[b], i2[b], out[b], epilogue=L.EPI_STORE32)
        ctx.save_for_backward(me16, inst16)
        ctx.pix_grad, ctx.me_grad = pix.requires_grad, me.requires_grad
        return out.view(B, Q, H, W)

    @staticmethod
    def backward(ctx, g):
        me16, inst16 = ctx.saved_tensors
        B, Q, Cc = me16.shape
        _, H, W, _ = inst16.shape
        HW = H * W
        if Q % 8:
            raise L.Sam3bError(f"mask einsum backward: the query count {Q} must be a multiple of 8")
        g = g.float().contiguous()
        sc = grad_scale(g)
        dm16 = torch.empty(B, Q, HW, device=g.device, dtype=me16.dtype)
        scale_cast(g, dm16, sc[0:1])
        dpix = dme = None
        if ctx.needs_input_grad[1]:
            dm16t = torch.empty(B, HW, Q, device=g.device, dtype=me16.dtype)
            transpose_cast(dm16, dm16t, B, Q, HW)
            me16t = torch.empty(B, Cc, Q, device=g.device, dtype=me16.dtype)
            transpose_cast(me16, me16t, B, Q, Cc)
            dp = torch.empty(B, HW, Cc, device=g.device, dtype=torch.float32)
            for b in range(B):
                _gemm32(dm16t[b], me16t[b], None, sc[1:2], out=dp[b])
            dpix = nhwc32_as_nchw(dp.view(B, H, W, Cc))
        if ctx.needs_input_grad[0]:
            inst16t = torch.empty(B, Cc, HW, device=g.device, dtype=me16.dtype)
            transpose_cast(inst16.view(B, HW, Cc), inst16t, B, HW, Cc)
            dme = torch.zeros(B, Q, Cc, device=g.device, dtype=torch.float32)
            kb = (HW + 63) // 64
            tiles = ((Q + 127) // 128) * ((Cc + 63) // 64)
            sk = max(1, min(kb, 296 // max(1, tiles)))
            for b in range(B):
                L.gemm(dm16[b], inst16t[b], dme[b], epilogue=L.EPI_ATOMIC_F32, splitk=sk, bn=64)
            dme = dme * sc[1:2]
        return dme, dpix


def mask_einsum(mask_embed: torch.Tensor, pixel_embed: torch.Tensor) -> torch.Tensor:
    """mask_embed [B, Q, C], pixel_embed [B, C, H, W] -> [B, Q, H, W] fp32."""
    return _MaskEinsumFn.apply(mask_embed, pixel_embed)
