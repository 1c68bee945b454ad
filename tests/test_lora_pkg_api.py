"""`sam3_lora.lora` surface (package-layout adapters, sam3_lora/lora/lora_layer.py:16-178, lora_utils.py:14-289).
CPU: names / shapes / state-dict keys / merge arithmetic, compared with the reference's own module when it is installed under
baseline/_ref.  GPU: the fused forward / backward against F.linear(x, W + B A * s)."""
import sys

import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from sam3_lora_b200 import lora as L
from sam3_lora_b200 import sam3_bridge


class _Attn(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.q_proj, self.k_proj, self.v_proj, self.out_proj = (nn.Linear(d, d) for _ in range(4))


class _Layer(nn.Module):
    def __init__(self, d, ff):
        super().__init__()
        self.self_attn = nn.MultiheadAttention(d, 4)
        self.cross_attn = _Attn(d)
        self.linear1, self.linear2 = nn.Linear(d, ff), nn.Linear(ff, d)
        self.norm = nn.LayerNorm(d)
        self.head = nn.Linear(d, 3)


class _Toy(nn.Module):
    def __init__(self, d=32, ff=64):
        super().__init__()
        self.encoder = nn.ModuleList([_Layer(d, ff) for _ in range(2)])
        self.proj = nn.Linear(d, d)


def _ref():
    root = sam3_bridge.reference_root()
    if root is None:
        return None
    if str(root) not in sys.path:
        sys.path.insert(0, str(root))
    import importlib

    return importlib.import_module("sam3_lora.lora")


@pytest.mark.parametrize("targets", [None, ["all"], ["q_proj", "v_proj"], ["linear1"], ["self_attn"]])
def test_injection_selects_the_same_modules_and_keys_as_the_reference(targets):
    torch.manual_seed(0)
    mine = L.inject_lora_into_model(_Toy(), L.LoRAConfig(rank=4, alpha=8.0, target_modules=targets), verbose=False)
    sd = L.get_lora_state_dict(mine)
    assert sd and all(k.endswith((".lora.lora_A", ".lora.lora_B")) for k in sd)
    for k, v in sd.items():
        mod = mine.get_submodule(k.rsplit(".lora.", 1)[0])
        want = (4, mod.in_features) if k.endswith("lora_A") else (mod.out_features, 4)
        assert tuple(v.shape) == want
    assert all(p.requires_grad for p in mine.proj.parameters())          # injection does not freeze the rest of the model
    assert all(not p.requires_grad for n, m in mine.named_modules() if isinstance(m, L.LinearWithLoRA) for p in m.linear.parameters())
    R = _ref()
    if R is None:
        pytest.skip("reference not installed under baseline/_ref: structural checks only")
    theirs = R.inject_lora_into_model(_Toy(), R.LoRAConfig(rank=4, alpha=8.0, target_modules=targets), verbose=False)
    ref_sd = R.get_lora_state_dict(theirs)
    assert set(sd) == set(ref_sd)
    assert {k: tuple(v.shape) for k, v in sd.items()} == {k: tuple(v.shape) for k, v in ref_sd.items()}
    assert len(L.get_lora_parameters(mine)) == len(R.get_lora_parameters(theirs))


def test_state_dict_round_trip_and_merge_arithmetic():
    torch.manual_seed(1)
    m = L.inject_lora_into_model(_Toy(), L.LoRAConfig(rank=4, alpha=8.0), verbose=False)
    for p in L.get_lora_parameters(m):
        nn.init.normal_(p, std=0.1)
    sd = {k: v.clone() for k, v in L.get_lora_state_dict(m).items()}
    m2 = L.inject_lora_into_model(_Toy(), L.LoRAConfig(rank=4, alpha=8.0), verbose=False)
    L.load_lora_state_dict(m2, sd)
    for k, v in L.get_lora_state_dict(m2).items():
        assert torch.equal(v, sd[k])
    w = m.encoder[0].linear1
    assert isinstance(w, L.LinearWithLoRA) and w.weight is w.linear.weight and w.bias is w.linear.bias
    expect = w.linear.weight + (w.lora.lora_B @ w.lora.lora_A) * (8.0 / 4)
    L.merge_lora_weights(m)
    merged = m.encoder[0].linear1
    assert type(merged) is nn.Linear and torch.allclose(merged.weight, expect, atol=1e-7)
    assert not any(isinstance(x, L.LinearWithLoRA) for x in m.modules())
    R = _ref()
    if R is not None:       # same numbers as the reference's merge on the same factors
        t = R.inject_lora_into_model(_Toy(), R.LoRAConfig(rank=4, alpha=8.0), verbose=False)
        t.load_state_dict({k: v for k, v in m2.state_dict().items()}, strict=True)
        R.merge_lora_weights(t)
        L.merge_lora_weights(m2)
        for (k1, v1), (k2, v2) in zip(sorted(t.state_dict().items()), sorted(m2.state_dict().items())):
            assert k1 == k2 and torch.allclose(v1, v2, atol=1e-7), k1
    with pytest.raises(ValueError):
        L.load_lora_state_dict(L.inject_lora_into_model(_Toy(), L.LoRAConfig(rank=8), verbose=False), sd)


def test_cpu_forward_raises_no_fallback():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    w = L.LinearWithLoRA(nn.Linear(64, 64), rank=4)
    with pytest.raises(RuntimeError):
        w(torch.randn(2, 64))


@pytest.mark.gpu
@pytest.mark.parametrize("p_drop", [0.0])
def test_fused_forward_backward_matches_dense_formula(p_drop):
    torch.manual_seed(2)
    lin = nn.Linear(256, 384).cuda()
    w = L.LinearWithLoRA(lin, rank=16, alpha=32.0, dropout=p_drop).cuda()
    with torch.no_grad():
        w.lora.lora_B.normal_(std=0.05)
    x = torch.randn(3, 130, 256, device="cuda", requires_grad=True)
    g = torch.randn(3, 130, 384, device="cuda")
    y = w(x)
    y.backward(g)
    got = (y.detach(), x.grad.clone(), w.lora.lora_A.grad.clone(), w.lora.lora_B.grad.clone())
    A, B = w.lora.lora_A.detach().clone().requires_grad_(True), w.lora.lora_B.detach().clone().requires_grad_(True)
    x2 = x.detach().clone().requires_grad_(True)
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        ref = F.linear(x2, lin.weight + (B @ A) * 2.0, lin.bias)
        ref.backward(g)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    from tests.helpers import rel_l2

    assert tuple(w.lora.lora_A.grad.shape) == (16, 256) and tuple(w.lora.lora_B.grad.shape) == (384, 16)
    assert rel_l2(got[0], ref.detach()) < 1e-3
    assert rel_l2(got[1], x2.grad) < 2e-3
    assert rel_l2(got[2], A.grad) < 3e-3 and rel_l2(got[3], B.grad) < 3e-3


@pytest.mark.gpu
def test_package_layout_adapters_inside_the_trunk_engine():
    """inject_lora_into_model(target 'fc') on a model that holds the native trunk: the engine copies the [r, in] / [out, r]
    factors into its own layout and hands back gradients in the package layout."""
    from sam3_lora_b200.vit import ViT
    from tests.helpers import rel_l2

    torch.manual_seed(3)
    kw = dict(img_size=224, embed_dim=128, depth=2, num_heads=2, mlp_ratio=4.75, window_size=8, global_att_blocks=(1,),
              pretrain_img_size=112, drop_path_rate=0.0, max_batch=2)
    a = ViT(**kw).cuda()
    b = ViT(**kw).cuda()
    b.load_state_dict(a.state_dict())
    from sam3_lora_b200.lora_layers import LoRAConfig as RootCfg, apply_lora_to_model

    L.inject_lora_into_model(a, L.LoRAConfig(rank=4, alpha=8.0, target_modules=["fc1", "fc2"]), verbose=False)
    for p in a.parameters():
        p.requires_grad = False
    for p in L.get_lora_parameters(a):
        p.requires_grad = True
        nn.init.normal_(p, std=0.05)
    a.cuda()
    apply_lora_to_model(b, RootCfg(rank=4, alpha=8, target_modules=["fc1", "fc2"], strict_reference_names=True))
    b.cuda()
    with torch.no_grad():
        for i in range(2):
            for t in ("fc1", "fc2"):
                src, dst = getattr(a.blocks[i].mlp, t).lora, getattr(b.blocks[i].mlp, t).lora
                dst.lora_A.copy_(src.lora_A.t())
                dst.lora_B.copy_(src.lora_B.t())
    img = torch.randn(2, 3, 224, 224, device="cuda")
    g = torch.randn(2, 128, 16, 16, device="cuda") * 0.1
    ya = a(img)[0]
    ya.backward(g)
    yb = b(img)[0]
    yb.backward(g)
    assert torch.equal(ya, yb)
    for i in range(2):
        for t in ("fc1", "fc2"):
            src, dst = getattr(a.blocks[i].mlp, t).lora, getattr(b.blocks[i].mlp, t).lora
            assert src.lora_A.grad.shape == src.lora_A.shape
            assert rel_l2(src.lora_A.grad.t(), dst.lora_A.grad) < 1e-5 and rel_l2(src.lora_B.grad.t(), dst.lora_B.grad) < 1e-5
