"""The MHA oracle is pinned against torch.nn.MultiheadAttention (the dependency that owns this arithmetic; the
reference calls it through MultiheadAttentionWrapper with need_weights=False, sam3/model/model_misc.py:31-34)."""
import torch

from oracle import mha_oracle as MO
from tests.helpers import rel_max


def _params(mha):
    return {"in_proj_weight": mha.in_proj_weight.detach(), "in_proj_bias": mha.in_proj_bias.detach(),
            "out_proj.weight": mha.out_proj.weight.detach(), "out_proj.bias": mha.out_proj.bias.detach()}


def test_oracle_matches_torch_mha_self_and_cross_with_masks():
    torch.manual_seed(0)
    E, H, B = 256, 8, 2
    mha = torch.nn.MultiheadAttention(E, H, dropout=0.1, batch_first=True).eval()
    torch.nn.init.normal_(mha.in_proj_bias, std=0.1)
    torch.nn.init.normal_(mha.out_proj.bias, std=0.1)
    x = torch.randn(B, 70, E)
    pos = torch.randn(B, 70, E)
    ref = mha(x + pos, x + pos, x, need_weights=False)[0]                # encoder self-attention pattern (q = k = x + pos, v = x)
    assert rel_max(MO.mha_forward(x + pos, x + pos, x, _params(mha), H), ref) < 2e-6
    mem = torch.randn(B, 33, E)
    kpm = torch.zeros(B, 33, dtype=torch.bool)
    kpm[0, 20:] = True
    bias = torch.randn(B * H, 70, 33)
    ref = mha(x, mem, mem, key_padding_mask=kpm, attn_mask=bias, need_weights=False)[0]   # decoder cross-attn with additive bias
    assert rel_max(MO.mha_forward(x, mem, mem, _params(mha), H, attn_mask=bias, key_padding_mask=kpm), ref) < 2e-6


def test_oracle_seq_first_wrapper_equivalence_and_dropout_mask_statistics():
    torch.manual_seed(1)
    E, H = 256, 8
    mha = torch.nn.MultiheadAttention(E, H, batch_first=False).eval()
    q = torch.randn(40, 3, E)
    kv = torch.randn(50, 3, E)
    ref = mha(q, kv, kv, need_weights=False)[0]
    got = MO.mha_forward(q.transpose(0, 1), kv.transpose(0, 1), kv.transpose(0, 1), _params(mha), H).transpose(0, 1)
    assert rel_max(got, ref) < 2e-6
    m = MO.attn_drop_scale_mask(123, 2, 8, 64, 96, 0.1)
    assert abs((m > 0).float().mean().item() - 0.9) < 0.01 and torch.allclose(m[m > 0], torch.tensor(1 / 0.9))
