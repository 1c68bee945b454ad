"""Row a8 on the GPU: conv.cu kernels against PyTorch, the neck / pixel decoder / mask head modules against the golden vectors
of the real reference (tests/golden/seg_small.npz) and against the CPU oracle at SAM3's channel widths.
Tolerances (fp16 operands, fp32 accumulation): forward rel-L2 <= 2e-3, input gradients rel-L2 <= 5e-3."""
import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from tests.helpers import GOLDEN, rel_l2

pytestmark = pytest.mark.gpu
SCALES = (4.0, 2.0, 1.0, 0.5)
DEV = "cuda:0"


class _Trunk(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.channel_list = [dim]

    def forward(self, x):
        return [x]


class _NoPos(nn.Module):
    def forward(self, x):
        return torch.zeros_like(x)


def _golden():
    z = np.load(GOLDEN / "seg_small.npz")
    return {k: torch.from_numpy(z[k]) for k in z.files}


# ------------------------------------------------------------------------------------------------------------
# kernels
# ------------------------------------------------------------------------------------------------------------
def test_casts_transposes_and_grad_scale():
    from sam3_lora_b200 import conv_ops as CO

    g = torch.Generator(device=DEV).manual_seed(0)
    x = torch.randn(3, 40, 72, device=DEV, generator=g) * 1e-7
    sc = CO.grad_scale(x)
    torch.cuda.synchronize()
    s = sc[0].item()
    assert s == 2.0 ** np.floor(np.log2(256.0 / x.abs().max().item())) and sc[1].item() == 1.0 / s
    z = CO.grad_scale(torch.zeros(64, device=DEV))
    assert z[0].item() == 1.0 and z[1].item() == 1.0
    # f32 -> 16 with scale, transposed
    y16 = torch.empty(3, 72, 40, device=DEV, dtype=torch.float16)
    CO.transpose_cast(x, y16, 3, 40, 72, sc[0:1])
    assert torch.equal(y16, (x * s).transpose(1, 2).half())
    # 16 -> f32 transposed back with 1/scale
    back = torch.empty(3, 40, 72, device=DEV)
    CO.transpose_cast(y16, back, 3, 72, 40, sc[1:2])
    assert rel_l2(back, x) < 1e-3
    # 16 -> 16 and f32 -> f32
    t16 = torch.empty(3, 40, 72, device=DEV, dtype=torch.float16)
    CO.transpose_cast(y16, t16, 3, 72, 40)
    assert torch.equal(t16, y16.transpose(1, 2))
    t32 = torch.empty(3, 72, 40, device=DEV)
    CO.transpose_cast(x, t32, 3, 40, 72)
    assert torch.equal(t32, x.transpose(1, 2).contiguous())
    # contiguous casts, accumulate
    a = torch.randn(1000, device=DEV, generator=g)
    h = torch.empty(1000, device=DEV, dtype=torch.bfloat16)
    CO.scale_cast(a, h)
    assert torch.equal(h, a.bfloat16())
    acc = torch.ones(1000, device=DEV)
    CO.scale_cast(h, acc, sc[1:2], accumulate=True)
    assert torch.allclose(acc, 1 + h.float() / s, rtol=1e-6)


def test_im2col_shuffle_pool_upsample_against_torch():
    from sam3_lora_b200 import conv_ops as CO

    g = torch.Generator(device=DEV).manual_seed(1)
    B, H, W, Cc = 2, 6, 10, 16
    x = torch.randn(B, H, W, Cc, device=DEV, generator=g).half()
    col = CO.im2col3x3(x)
    ref = F.unfold(x.permute(0, 3, 1, 2).float(), 3, padding=1)                 # [B, C*9, HW] with k = (c, ky, kx)
    ref = ref.view(B, Cc, 9, H * W).permute(0, 3, 2, 1).reshape(B * H * W, 9 * Cc)
    assert torch.equal(col.float(), ref)
    # pixel shuffle: columns (di, dj, c)
    u = torch.randn(B * H * W, 4 * Cc, device=DEV, generator=g).half()
    sh = CO.pixel_shuffle2(u, B, H, W, Cc)
    ref = u.view(B, H, W, 2, 2, Cc).permute(0, 1, 3, 2, 4, 5).reshape(B, 2 * H, 2 * W, Cc)
    assert torch.equal(sh, ref)
    assert torch.equal(CO.pixel_unshuffle2(sh), u)
    shg = CO.pixel_shuffle2(u, B, H, W, Cc, gelu=True)
    assert rel_l2(shg.float(), F.gelu(ref.float())) < 6e-4
    dy = torch.randn(B, 2 * H, 2 * W, Cc, device=DEV, generator=g).half()
    un = CO.pixel_unshuffle2(dy, u)
    uf = u.float().requires_grad_(True)
    F.gelu(uf).backward(CO.pixel_unshuffle2(dy).float())
    assert rel_l2(un.float(), uf.grad) < 6e-4
    # max pool with ties (values on a coarse grid) : first maximum wins, like ATen
    xt = (torch.randint(0, 3, (B, H, W, Cc), device=DEV, generator=g).float()).half()
    xr = xt.permute(0, 3, 1, 2).float().requires_grad_(True)
    pr = F.max_pool2d(xr, 2, 2)
    assert torch.equal(CO.maxpool2_fwd(xt).permute(0, 3, 1, 2).float(), pr)
    dyp = torch.randn(B, H // 2, W // 2, Cc, device=DEV, generator=g).half()
    pr.backward(dyp.permute(0, 3, 1, 2).float())
    dx = torch.ones(B, H, W, Cc, device=DEV)
    half = torch.full((1,), 0.5, device=DEV)
    CO.maxpool2_bwd(xt, dyp, half, dx)
    assert torch.allclose(dx, 1 + 0.5 * xr.grad.permute(0, 2, 3, 1), atol=1e-6)
    # nearest up-sample + add and its adjoint
    prev = torch.randn(B, H // 2, W // 2, Cc, device=DEV, generator=g).half()
    out = CO.upsample_add(prev, x)
    ref = x.float() + F.interpolate(prev.permute(0, 3, 1, 2).float(), size=(H, W), mode="nearest").permute(0, 2, 3, 1)
    assert torch.equal(out, ref.half())
    adj = CO.upsample_add_bwd(dy[:, :H, :W].contiguous(), H // 2, W // 2)
    ref = dy[:, :H, :W].float().view(B, H // 2, 2, W // 2, 2, Cc).sum((2, 4))
    assert rel_l2(adj.float(), ref) < 1e-3


@pytest.mark.parametrize("B,H,W,Cc,co", [(2, 20, 24, 64, 40), (1, 16, 8, 128, 256), (3, 37, 40, 256, 256), (2, 72, 72, 256, 264)])
def test_implicit_gemm_conv3x3_against_torch_and_im2col(B, H, W, Cc, co):
    """TMA implicit-GEMM conv (zero-filled 4-D boxes) vs F.conv2d on the same 16-bit-rounded operands, fp32 and 16-bit
    outputs, image heights that are not a multiple of the 16-row tile, Cout above one 256-wide N tile, and vs im2col + GEMM."""
    from sam3_lora_b200 import conv_ops as CO

    torch.backends.cudnn.allow_tf32 = False          # the PyTorch side of the comparison must be real fp32
    g = torch.Generator(device=DEV).manual_seed(11)
    x = torch.randn(B, H, W, Cc, device=DEV, generator=g).half()
    wgt = (torch.randn(co, Cc, 3, 3, device=DEV, generator=g) * (9 * Cc) ** -0.5)
    bias = torch.randn(co, device=DEV, generator=g)
    w9 = wgt.permute(0, 2, 3, 1).reshape(co, 9 * Cc).half().contiguous()
    ref = F.conv2d(x.permute(0, 3, 1, 2).float(), w9.float().view(co, 3, 3, Cc).permute(0, 3, 1, 2), bias, padding=1)
    ref = ref.permute(0, 2, 3, 1).reshape(B * H * W, co)
    assert CO.L.load().sam3b_conv3x3_supported(H, W, Cc, co)
    y32 = CO.conv3x3(x, w9, bias, out_f32=True)
    assert rel_l2(y32, ref) < 1e-5                   # identical operands; fp32 accumulation order differs over K = 9C
    y16 = CO.conv3x3(x, w9, bias)
    assert y16.dtype == torch.float16 and rel_l2(y16.float(), ref) < 6e-4
    CO.IMPLICIT_CONV = False
    try:
        y_im2col = CO.conv3x3(x, w9, bias, out_f32=True)
    finally:
        CO.IMPLICIT_CONV = True
    assert rel_l2(y32, y_im2col) < 1e-5
    xb = x.bfloat16()
    yb = CO.conv3x3(xb, w9.bfloat16(), None, out_f32=True)
    refb = F.conv2d(xb.permute(0, 3, 1, 2).float(), w9.bfloat16().float().view(co, 3, 3, Cc).permute(0, 3, 1, 2), None, padding=1)
    assert rel_l2(yb, refb.permute(0, 2, 3, 1).reshape(B * H * W, co)) < 1e-5


@pytest.mark.parametrize("Cc,HW", [(32, 4 * 36), (256, 20 * 27)])
def test_groupnorm_relu_forward_backward(Cc, HW):
    from sam3_lora_b200 import conv_ops as CO

    g = torch.Generator(device=DEV).manual_seed(2)
    B, G = 3, 8
    x = torch.randn(B, HW, Cc, device=DEV, generator=g) * 2 + 0.5
    gamma = 1 + 0.3 * torch.randn(Cc, device=DEV, generator=g)
    beta = 0.3 * torch.randn(Cc, device=DEV, generator=g)
    stat = CO.groupnorm_stats(x, B, HW, Cc, G, 1e-5)
    y = torch.empty(B, HW, Cc, device=DEV)
    CO.groupnorm_relu_fwd(x, stat, gamma, beta, B, HW, Cc, G, y)
    xr = x.transpose(1, 2).clone().requires_grad_(True)                             # [B, C, HW]
    ref = F.relu(F.group_norm(xr, G, gamma, beta, 1e-5))
    assert rel_l2(y, ref.transpose(1, 2)) < 2e-6
    y16 = torch.empty(B, HW, Cc, device=DEV, dtype=torch.float16)
    CO.groupnorm_relu_fwd(x, stat, gamma, beta, B, HW, Cc, G, y16)
    assert rel_l2(y16.float(), ref.transpose(1, 2)) < 6e-4
    dy = torch.randn(B, HW, Cc, device=DEV, generator=g).half()
    ref.backward(dy.float().transpose(1, 2))
    dx = CO.groupnorm_relu_bwd(dy.view(B * HW, Cc), x, stat, gamma, beta, B, HW, Cc, G)
    assert rel_l2(dx.float().view(B, HW, Cc), xr.grad.transpose(1, 2)) < 1e-3


# ------------------------------------------------------------------------------------------------------------
# modules vs the reference's golden vectors
# ------------------------------------------------------------------------------------------------------------
def _frozen(m):
    for p in m.parameters():
        p.requires_grad_(False)
    return m


def test_neck_matches_reference_golden():
    from sam3_lora_b200.necks import Sam3DualViTDetNeck

    z = _golden()
    neck = Sam3DualViTDetNeck(_Trunk(64), _NoPos(), d_model=32, scale_factors=SCALES)
    missing, unexpected = neck.load_state_dict({k[len("neck.param."):]: v for k, v in z.items() if k.startswith("neck.param.")})
    assert not missing and not unexpected
    neck = _frozen(neck).to(DEV)
    x = z["neck.x"].to(DEV).requires_grad_(True)
    feats, pos, s2, _ = neck(x)
    assert s2 is None and len(feats) == 4
    for i, f in enumerate(feats):
        assert f.shape == z[f"neck.out{i}"].shape and f.dtype == torch.float32
        assert rel_l2(f.detach().cpu(), z[f"neck.out{i}"]) < 2e-3, i
    sum((f * z[f"neck.cot{i}"].to(DEV)).sum() for i, f in enumerate(feats)).backward()
    assert rel_l2(x.grad.cpu(), z["neck.dx"]) < 5e-3
    # one branch at a time (the others receive no gradient)
    for i in range(4):
        x2 = z["neck.x"].to(DEV).requires_grad_(True)
        fi = neck(x2)[0][i]
        (fi * z[f"neck.cot{i}"].to(DEV)).sum().backward()
        assert rel_l2(x2.grad.cpu(), z[f"neck.dx{i}"]) < 5e-3, i


def _seg_modules(z, d=32):
    from sam3_lora_b200.maskformer_segmentation import PixelDecoder, UniversalSegmentationHead

    head = UniversalSegmentationHead(d, 2, PixelDecoder(d, 2))
    sd = {k[len("seg.param."):]: v for k, v in z.items() if k.startswith("seg.param.")}
    missing, unexpected = head.load_state_dict(sd)
    assert not missing and not unexpected
    return _frozen(head).to(DEV)


def test_pixel_decoder_heads_and_mask_einsum_match_reference_golden():
    from sam3_lora_b200 import conv_ops as CO

    z = _golden()
    head = _seg_modules(z)
    feats = [z[f"seg.feat{i}"].to(DEV).requires_grad_(True) for i in range(3)]
    q = z["seg.queries"].to(DEV).requires_grad_(True)
    pix = head.pixel_decoder(feats)
    assert rel_l2(pix.detach().cpu(), z["seg.pixel_embed"]) < 2e-3
    inst = CO.conv1x1_forward(pix, head.instance_seg_head)
    masks = head.mask_predictor(q, inst)
    sem = CO.conv1x1_forward(pix, head.semantic_seg_head)
    assert masks.dtype == torch.float32 and masks.shape == z["seg.masks"].shape
    assert rel_l2(masks.detach().cpu(), z["seg.masks"]) < 2e-3        # mask logits (north-star tolerance 1e-3 rel, checked below)
    assert (masks.detach().cpu() - z["seg.masks"]).abs().max() < 1e-3 * z["seg.masks"].abs().max() * 4
    assert rel_l2(sem.detach().cpu(), z["seg.semantic"]) < 2e-3
    ((masks * z["seg.cot_masks"].to(DEV)).sum() + (sem * z["seg.cot_semantic"].to(DEV)).sum()).backward()
    # Gradients cross two GroupNorm+ReLU stages and the ReLUs of the mask-embedding MLP.  With 16-bit GEMM operands a
    # pre-activation within ~1e-3 of zero can land on the other side of the ReLU than in the fp32 reference, which flips
    # that element's whole gradient contribution (about 0.1 % of the elements -> a few % in rel-L2; the reference itself
    # shows the same effect under its TF32 setting).  So: (1) a loose bound against the exact fp32 golden, and (2) the
    # tight bound against the oracle evaluated with the same operand rounding (identical ReLU masks).
    from oracle import seg_oracle as SO
    p = {k[len("seg.param."):]: v for k, v in z.items() if k.startswith("seg.param.")}
    fr = [z[f"seg.feat{i}"].clone().requires_grad_(True) for i in range(3)]
    qr = z["seg.queries"].clone().requires_grad_(True)
    mr, sr = SO.seg_head(fr, qr, p, operand_dtype=torch.float16)
    ((mr * z["seg.cot_masks"]).sum() + (sr * z["seg.cot_semantic"]).sum()).backward()
    err = {"fwd_masks_vs_rounded_oracle": rel_l2(masks.detach().cpu(), mr.detach()),
           "fwd_sem_vs_rounded_oracle": rel_l2(sem.detach().cpu(), sr.detach()),
           "dq_vs_golden": rel_l2(q.grad.cpu(), z["seg.dqueries"]), "dq_vs_rounded_oracle": rel_l2(q.grad.cpu(), qr.grad)}
    for i in range(3):
        err[f"dfeat{i}_vs_golden"] = rel_l2(feats[i].grad.cpu(), z[f"seg.dfeat{i}"])
        err[f"dfeat{i}_vs_rounded_oracle"] = rel_l2(feats[i].grad.cpu(), fr[i].grad)
    print("seg golden errors:", {k: f"{v:.2e}" for k, v in err.items()})
    for k, v in err.items():
        # the query gradient crosses the MLP's ReLUs on only 2 x 8 x 32 hidden values: ONE flipped unit is 4 % in rel-L2
        bound = 5e-4 if k.startswith("fwd") else (5e-3 if k.endswith("rounded_oracle") else (1.5e-1 if k.startswith("dq") else 6e-2))
        assert v < bound, (k, err)
    # decoder-layer form
    ml = head.mask_predictor(z["seg.queries_layers"].to(DEV), inst.detach())
    assert rel_l2(ml.cpu(), z["seg.masks_layers"]) < 2e-3


def test_universal_head_forward_wiring_matches_oracle():
    """UniversalSegmentationHead.forward: per-query gather of the backbone maps, encoder tokens as the coarsest level."""
    from oracle import seg_oracle as SO

    z = _golden()
    head = _seg_modules(z)
    p = {k[len("seg.param."):]: v for k, v in z.items() if k.startswith("seg.param.")}
    g = torch.Generator().manual_seed(5)
    feats = [torch.randn(2, 32, 32, 32, generator=g), torch.randn(2, 32, 16, 16, generator=g), torch.randn(2, 32, 8, 8, generator=g)]
    enc = torch.randn(64 + 5, 3, 32, generator=g)                     # [tokens (+5 prompt tokens), queries-batch 3, C]
    ids = torch.tensor([1, 0, 1])
    q = torch.randn(2, 3, 8, 32, generator=g)                         # [layers, B, Q, C]; aux_masks False -> last layer
    out = head([f.to(DEV) for f in feats], q.to(DEV), ids.to(DEV), encoder_hidden_states=enc.to(DEV))
    rf = [f[ids] for f in feats]
    rf[-1] = enc.permute(1, 2, 0)[..., :64].reshape(-1, 32, 8, 8)
    masks, sem = SO.seg_head(rf, q[-1], p)
    assert rel_l2(out["pred_masks"].cpu(), masks) < 2e-3
    assert rel_l2(out["semantic_seg"].cpu(), sem) < 2e-3
    assert out["presence_logit"] is None


# ------------------------------------------------------------------------------------------------------------
# SAM3 channel widths vs the oracle; tiny gradients survive the fp16 chain
# ------------------------------------------------------------------------------------------------------------
def test_neck_and_mask_head_at_sam3_widths_against_oracle():
    from oracle import seg_oracle as SO
    from sam3_lora_b200.maskformer_segmentation import PixelDecoder, UniversalSegmentationHead
    from sam3_lora_b200.necks import Sam3DualViTDetNeck
    from sam3_lora_b200 import conv_ops as CO

    dim, d, B, G0, Q = 1024, 256, 2, 12, 16
    g = torch.Generator().manual_seed(6)
    # ---- neck: forward and input gradient for mean-reduced cotangents (~1e-9 per element, far below fp16's 6e-8) ----
    pn = SO.make_neck_params(dim, d, SCALES, seed=4)
    neck = Sam3DualViTDetNeck(_Trunk(dim), _NoPos(), d_model=d, scale_factors=SCALES)
    neck.load_state_dict(pn)
    neck = _frozen(neck).to(DEV)
    x = torch.randn(B, dim, G0, G0, generator=g)
    xr = x.clone().requires_grad_(True)
    ref = SO.neck(xr, pn, SCALES)
    xq = x.clone().requires_grad_(True)
    refq = SO.neck(xq, pn, SCALES, operand_dtype=torch.float16)     # same max-pool arg-max as the 16-bit activations
    xg = x.to(DEV).requires_grad_(True)
    feats = neck(xg)[0]
    cots = [torch.randn(r.shape, generator=g) / r.numel() * 1e-4 for r in ref]
    for f, r, rq in zip(feats, ref, refq):
        assert rel_l2(f.detach().cpu(), r.detach()) < 2e-3
        assert rel_l2(f.detach().cpu(), rq.detach()) < 5e-4
    sum((r * c).sum() for r, c in zip(ref, cots)).backward()
    sum((r * c).sum() for r, c in zip(refq, cots)).backward()
    sum((f * c.to(DEV)).sum() for f, c in zip(feats, cots)).backward()
    assert xr.grad.abs().max() < 1e-6
    e_exact, e_round = rel_l2(xg.grad.cpu(), xr.grad), rel_l2(xg.grad.cpu(), xq.grad)
    print(f"neck dx: vs exact fp32 oracle {e_exact:.2e} (max-pool near-ties), vs operand-rounded oracle {e_round:.2e}")
    assert e_exact < 3e-2 and e_round < 5e-3
    # ---- pixel decoder + heads + einsum on the SAME inputs, against the oracle with the same operand rounding ----
    ps = SO.make_seg_params(d, 2, seed=7)
    head = UniversalSegmentationHead(d, 2, PixelDecoder(d, 2))
    head.load_state_dict(ps)
    head = _frozen(head).to(DEV)
    fin = [f.detach().cpu().contiguous() for f in feats[:3]]
    fr = [f.clone().requires_grad_(True) for f in fin]
    fg = [f.to(DEV).requires_grad_(True) for f in fin]
    q = torch.randn(B, Q, d, generator=g)
    qr, qg = q.clone().requires_grad_(True), q.to(DEV).requires_grad_(True)
    masks_x, _ = SO.seg_head(fin, q, ps)                                              # exact fp32 reference arithmetic
    masks_r, sem_r = SO.seg_head(fr, qr, ps, operand_dtype=torch.float16)
    pix = head.pixel_decoder(fg)
    masks = head.mask_predictor(qg, CO.conv1x1_forward(pix, head.instance_seg_head))
    sem = CO.conv1x1_forward(pix, head.semantic_seg_head)
    assert rel_l2(masks.detach().cpu(), masks_x) < 2e-3
    assert rel_l2(masks.detach().cpu(), masks_r.detach()) < 5e-4
    assert rel_l2(sem.detach().cpu(), sem_r.detach()) < 5e-4
    cm = torch.randn(masks_r.shape, generator=g) / masks_r.numel() * 1e-4
    cs = torch.randn(sem_r.shape, generator=g) / sem_r.numel() * 1e-4
    ((masks_r * cm).sum() + (sem_r * cs).sum()).backward()
    ((masks * cm.to(DEV)).sum() + (sem * cs.to(DEV)).sum()).backward()
    errs = [rel_l2(a.grad.cpu(), b.grad) for a, b in zip(fg, fr)] + [rel_l2(qg.grad.cpu(), qr.grad)]
    print("seg head input-gradient errors vs operand-rounded oracle:", [f"{e:.2e}" for e in errs])
    assert max(errs) < 5e-3, errs


def test_pixel_decoder_broadcasts_batch_one_backbone_maps():
    """UniversalSegmentationHead with a single image and several prompts: backbone maps have batch 1, the encoder map has
    one entry per prompt, and the reference adds them by broadcasting (maskformer_segmentation.py:118-120, 209)."""
    from oracle import seg_oracle as SO
    from sam3_lora_b200.maskformer_segmentation import PixelDecoder

    d = 32
    ps = SO.make_seg_params(d, 2, seed=41)
    pd = PixelDecoder(d, 2)
    pd.load_state_dict({k[len("pixel_decoder."):]: v for k, v in ps.items() if k.startswith("pixel_decoder.")})
    pd = _frozen(pd).to(DEV)
    g = torch.Generator().manual_seed(42)
    feats = [torch.randn(1, d, 32, 32, generator=g), torch.randn(1, d, 16, 16, generator=g), torch.randn(3, d, 8, 8, generator=g)]
    fr = [f.clone().requires_grad_(True) for f in feats]
    fg = [f.to(DEV).requires_grad_(True) for f in feats]
    ref = SO.pixel_decoder(fr, ps, prefix="pixel_decoder.", operand_dtype=torch.float16)
    out = pd(fg)
    assert out.shape == ref.shape == (3, d, 32, 32)
    assert rel_l2(out.detach().cpu(), ref.detach()) < 5e-4
    cot = torch.randn(ref.shape, generator=g)
    (ref * cot).sum().backward()
    (out * cot.to(DEV)).sum().backward()
    for a, b in zip(fg, fr):
        assert a.grad.shape == b.grad.shape and rel_l2(a.grad.cpu(), b.grad) < 5e-3


def test_cpu_tensors_are_rejected():
    from sam3_lora_b200 import _lib
    from sam3_lora_b200.necks import Sam3DualViTDetNeck

    neck = _frozen(Sam3DualViTDetNeck(_Trunk(64), _NoPos(), d_model=32))
    with pytest.raises(_lib.Sam3bError):
        neck(torch.randn(1, 64, 8, 8))
