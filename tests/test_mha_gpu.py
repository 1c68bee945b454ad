"""Row a7: fused MultiheadAttention (E=256, 8 heads, head_dim 32 as padded 64-wide heads) vs the oracle:
self / cross attention, additive attn_mask, key_padding_mask, dropout on the probabilities, LoRA on q/k/v/out."""
import pytest
import torch

from oracle import mha_oracle as MO
from tests.helpers import rel_l2

pytestmark = pytest.mark.gpu

FWD_TOL, GRAD_TOL = 2e-3, 6e-3


def _make(E=256, H=8, batch_first=True, dropout=0.0, lora=True, seed=0):
    from sam3_lora_b200.lora_layers import LoRAConfig, apply_lora_to_model
    from sam3_lora_b200.mha import MultiheadAttention

    torch.manual_seed(seed)
    m = MultiheadAttention(E, H, dropout=dropout, batch_first=batch_first)
    torch.nn.init.normal_(m.in_proj_bias, std=0.1)
    torch.nn.init.normal_(m.out_proj.bias, std=0.1)
    holder = torch.nn.Module()
    holder.attn = m
    if lora:
        apply_lora_to_model(holder, LoRAConfig(rank=8, alpha=16, target_modules=["q_proj", "k_proj", "v_proj", "out_proj"]))
        for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
            torch.nn.init.normal_(getattr(m, n).lora.lora_B, std=0.05)
    return holder.cuda(), m


def _oracle_params(m):
    p = {"in_proj_weight": m.in_proj_weight.detach().cpu(), "in_proj_bias": m.in_proj_bias.detach().cpu(),
         "out_proj.weight": m._out_linear.weight.detach().cpu(), "out_proj.bias": m._out_linear.bias.detach().cpu()}
    leaves = {}
    for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
        mod = getattr(m, n, None)
        if mod is not None and hasattr(mod, "lora"):
            for ab in ("lora_A", "lora_B"):
                t = getattr(mod.lora, ab).detach().cpu().clone().requires_grad_(True)
                p[f"{n}.lora.{ab}"] = t
                leaves[f"{n}.lora.{ab}"] = t
    return p, leaves


def _compare(m, q, k, v, attn_mask=None, kpm=None, dropout=None, batch_first=True, g_scale=1.0, share=None):
    m.dropout_seed_override = dropout[1] if dropout else None
    qc, kc, vc = (t.clone().cuda().requires_grad_(True) for t in (q, k, v))
    if share in ("qk", "qkv"):      # the module sees ONE tensor object for these arguments (encoder: q = k = x + pos)
        kc = qc
    if share in ("kv", "qkv"):
        vc = kc
    out = m(qc, kc, vc, key_padding_mask=None if kpm is None else kpm.cuda(), attn_mask=None if attn_mask is None else attn_mask.cuda())[0]
    g = torch.randn(out.shape, generator=torch.Generator().manual_seed(3)) * g_scale
    out.backward(g.cuda())
    p, leaves = _oracle_params(m)
    qo, ko, vo = (t.clone().requires_grad_(True) for t in (q, k, v))
    if share in ("qk", "qkv"):
        ko = qo
    if share in ("kv", "qkv"):
        vo = ko
    bf = (lambda t: t) if batch_first else (lambda t: t.transpose(0, 1))
    lo = getattr(m, "q_proj", None)
    scaling = lo.lora.scaling if lo is not None and hasattr(lo, "lora") else 1.0
    ref = MO.mha_forward(bf(qo), bf(ko), bf(vo), p, m.num_heads, attn_mask=attn_mask, key_padding_mask=kpm, scaling=scaling,
                         dropout=dropout)
    ref = bf(ref)
    ref.backward(g)
    assert rel_l2(out.detach().cpu(), ref.detach()) < FWD_TOL
    assert rel_l2(qc.grad.cpu(), qo.grad) < GRAD_TOL
    if kc is not qc:
        assert rel_l2(kc.grad.cpu(), ko.grad) < GRAD_TOL
    if vc is not kc and vc is not qc:
        assert rel_l2(vc.grad.cpu(), vo.grad) < GRAD_TOL
    for n, t in leaves.items():
        mod, ab = n.split(".lora.")
        got = getattr(getattr(m, mod).lora, ab).grad.cpu()
        assert rel_l2(got, t.grad) < GRAD_TOL, n
    assert m.in_proj_weight.grad is None


def test_self_attention_encoder_pattern_with_dropout():
    holder, m = _make(dropout=0.1)
    m.train()
    gen = torch.Generator().manual_seed(0)
    x, pos = torch.randn(2, 576, 256, generator=gen), torch.randn(2, 576, 256, generator=gen)
    _compare(m, x + pos, x + pos, x, dropout=(0.1, 4242))


def test_cross_attention_to_prompt_with_key_padding_mask():
    holder, m = _make()
    m.eval()
    gen = torch.Generator().manual_seed(1)
    q, mem = torch.randn(2, 300, 256, generator=gen), torch.randn(2, 33, 256, generator=gen)
    kpm = torch.zeros(2, 33, dtype=torch.bool)
    kpm[0, 12:] = True
    kpm[1, 30:] = True
    _compare(m, q, mem, mem, kpm=kpm)


def test_decoder_cross_attention_with_additive_bias_seq_first():
    holder, m = _make(batch_first=False)
    m.eval()
    gen = torch.Generator().manual_seed(2)
    q, mem = torch.randn(201, 2, 256, generator=gen), torch.randn(640, 2, 256, generator=gen)
    bias = torch.randn(2 * 8, 201, 640, generator=gen)
    _compare(m, q, mem, mem, attn_mask=bias, batch_first=False)


# ---- the real DETR sizes: 72 x 72 = 5184 image tokens, 200 + 200 + 1 queries, 32 + 1 prompt tokens -----------
def test_encoder_self_attention_at_5184_tokens_with_dropout():
    """TransformerEncoderLayer.forward_pre's self-attention (encoder.py:176-186): q = k = x + pos, v = x over the 5184
    image tokens, attention dropout 0.1 active (same stateless mask in the oracle)."""
    holder, m = _make(dropout=0.1)
    m.train()
    gen = torch.Generator().manual_seed(10)
    x, pos = torch.randn(1, 5184, 256, generator=gen), torch.randn(1, 5184, 256, generator=gen)
    _compare(m, x + pos, x + pos, x, dropout=(0.1, 777))


def test_decoder_image_cross_attention_401x5184_with_box_rpb_bias():
    """TransformerDecoderLayer's image cross-attention (decoder.py:156-175): 401 queries (200 o2o + 200 o2m + presence)
    against 5184 memory tokens with the additive fp32 box-RPB bias [B*8, 401, 5184] (decoder.py:331-408, 516-524),
    sequence-first layout."""
    holder, m = _make(batch_first=False)
    m.eval()
    gen = torch.Generator().manual_seed(11)
    q, mem = torch.randn(401, 1, 256, generator=gen), torch.randn(5184, 1, 256, generator=gen)
    bias = torch.randn(8, 401, 5184, generator=gen) * 2.0
    _compare(m, q, mem, mem, attn_mask=bias, batch_first=False)


def test_image_to_prompt_cross_attention_with_loss_sized_gradients():
    """Encoder cross-attention of the 5184 image tokens to the 33 prompt tokens (encoder.py:188-198) with cotangents of the
    size a real loss hands back (1e-8): dK / dV are sums over all 5184 queries; with the backward running on a power-of-two
    multiple of the cotangent they overflowed fp16 before the scale was bounded by 32768 / Lq (found by the a9 step)."""
    holder, m = _make()
    m.eval()
    gen = torch.Generator().manual_seed(12)
    q, mem = torch.randn(1, 5184, 256, generator=gen), torch.randn(1, 33, 256, generator=gen)
    kpm = torch.zeros(1, 33, dtype=torch.bool)
    kpm[0, 20:] = True
    _compare(m, q, mem, mem, kpm=kpm, g_scale=1e-8)


@pytest.mark.parametrize("pattern", ["encoder_self", "cross_shared_kv", "text_qkv", "all_distinct", "decoder_self_seq_first"])
def test_adapter_free_fused_path_matches_oracle(pattern):
    """Without adapters the module runs _FusedMHAFn: 16-bit between the kernels, q|k (or k|v, or q|k|v) projected by one
    GEMM when they share their input tensor, gradients of a shared input summed by one dgrad GEMM."""
    gen = torch.Generator().manual_seed(20)
    if pattern == "text_qkv":                         # CLIP-style resblock: 1024 wide, 16 heads of 64, causal mask, x, x, x
        holder, m = _make(E=1024, H=16, lora=False)
        m.eval()
        x = torch.randn(2, 32, 1024, generator=gen)
        causal = torch.full((32, 32), float("-inf")).triu_(1)
        _compare(m, x, x, x, attn_mask=causal, share="qkv")
        assert m._packed16
        return
    if pattern == "decoder_self_seq_first":
        holder, m = _make(lora=False, batch_first=False, dropout=0.1)
        m.train()
        x, pos = torch.randn(201, 2, 256, generator=gen), torch.randn(201, 2, 256, generator=gen)
        _compare(m, x + pos, x + pos, x, dropout=(0.1, 99), batch_first=False, share="qk")
        return
    holder, m = _make(lora=False)
    m.eval()
    if pattern == "encoder_self":
        x, pos = torch.randn(2, 576, 256, generator=gen), torch.randn(2, 576, 256, generator=gen)
        _compare(m, x + pos, x + pos, x, share="qk")
    elif pattern == "cross_shared_kv":
        q, mem = torch.randn(2, 300, 256, generator=gen), torch.randn(2, 33, 256, generator=gen)
        kpm = torch.zeros(2, 33, dtype=torch.bool)
        kpm[0, 12:] = True
        _compare(m, q, mem, mem, kpm=kpm, share="kv", g_scale=1e-7)
    else:
        q, k, v = (torch.randn(2, n, 256, generator=gen) for n in (130, 70, 70))
        _compare(m, q, k, v)
    # the composed path gives the same numbers (same kernels, fp32 detours): switch and compare outputs
    m.fused = False
    x = torch.randn(2, 64, 256, generator=gen).cuda()
    a = m(x, x, x + 1.0)[0]
    m.fused = True
    b = m(x, x, x + 1.0)[0]
    assert rel_l2(b.cpu(), a.cpu()) < 1e-3


def test_dropout_keep_bits_equal_the_oracle_mask_in_both_orientations():
    """sam3b_attention_dropout_bits: bit (k & 31) of bits[(bh*Lq + q)*pitch(Lk) + k/32] and bit (q & 31) of
    bitsT[(bh*Lk + k)*pitch(Lq) + q/32] are the oracle's keep decisions (same stateless hash as the inline evaluation)."""
    import numpy as np

    from sam3_lora_b200 import _lib as L
    from sam3_lora_b200._abi import AttnDesc

    nseg, H, Lq, Lk, p, seed = 2, 3, 288, 320, 0.3, 4321
    d = AttnDesc()
    d.nseg, d.heads, d.Lq, d.Lk, d.drop_p, d.drop_seed = nseg, H, Lq, Lk, p, seed
    old = L.DROPOUT_BITS_MIN_SCORES
    L.DROPOUT_BITS_MIN_SCORES = 1
    try:
        bits, bitsT = L.attention_dropout_bits(d)
    finally:
        L.DROPOUT_BITS_MIN_SCORES = old
    torch.cuda.synchronize()
    pitch = lambda n: (n // 32 + 7) // 8 * 8          # noqa: E731
    ref = (MO.attn_drop_scale_mask(seed, nseg, H, Lq, Lk, p) > 0).reshape(nseg * H, Lq, Lk).numpy()
    b = bits.cpu().numpy().view(np.uint32).reshape(nseg * H, Lq, pitch(Lk))[:, :, :Lk // 32]
    got = ((b[..., None] >> np.arange(32, dtype=np.uint32)) & 1).reshape(nseg * H, Lq, Lk).astype(bool)
    assert np.array_equal(got, ref)
    bt = bitsT.cpu().numpy().view(np.uint32).reshape(nseg * H, Lk, pitch(Lq))[:, :, :Lq // 32]
    gotT = ((bt[..., None] >> np.arange(32, dtype=np.uint32)) & 1).reshape(nseg * H, Lk, Lq).astype(bool)
    assert np.array_equal(gotT, ref.transpose(0, 2, 1))
    assert abs(ref.mean() - (1 - p)) < 5e-3


def test_no_adapters_matches_torch_module_directly():
    from sam3_lora_b200.mha import replace_torch_mha

    torch.manual_seed(5)
    holder = torch.nn.Module()
    holder.attn = torch.nn.MultiheadAttention(256, 8, batch_first=True)
    ref_mod = holder.attn
    x = torch.randn(2, 130, 256)
    ref = ref_mod.eval()(x, x, x, need_weights=False)[0]
    assert replace_torch_mha(holder) == 1
    holder = holder.cuda().eval()
    out = holder.attn(x.cuda(), x.cuda(), x.cuda())[0]
    assert rel_l2(out.cpu(), ref.detach()) < FWD_TOL
