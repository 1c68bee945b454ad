"""Generates tests/golden/vit_small.npz by running the REAL reference code.

Run in the build container only (needs /root/reference; it does not exist on the GPU box):

    python tests/golden/make_golden.py

What is executed from the reference (unmodified, imported from /root/reference):
  * sam3/model/vitdet.py  `ViT` (PatchEmbed, get_abs_pos tiling, ln_pre, Block, Attention with
    compute_axial_cis / apply_rotary_enc, window_partition/unpartition, F.scaled_dot_product_attention)
  * lora_layers.py        `LoRALinear`, `LoRAConfig`, `apply_lora_to_model`, `save_lora_weights`
Third-party pieces the reference imports but that are not installed here are stubbed with their
published definitions: timm.layers.{Mlp, DropPath, trunc_normal_}.

The ViT is instantiated at a reduced size (same code path, smaller hyper-parameters) so the fixture
stays ~1 MB: 224x224 image, patch 14 -> 16x16 tokens, window 8 (4 windows of 64 tokens),
embed 128, 2 heads (head_dim 64), depth 2 with block 1 global, MLP hidden 608.
Adapters: fc1/fc2 through the reference's own apply_lora_to_model; q/k/v/out ("north-star"
aliasing) by composing the reference's LoRALinear over row-slices of the fused qkv.
"""
from __future__ import annotations

import sys
import types
from pathlib import Path

import numpy as np
import torch
import torch.nn as nn

ROOT = Path(__file__).resolve().parents[2]
REF = Path("/root/reference")
sys.path.insert(0, str(ROOT))


def install_stubs():
    """timm.layers stand-ins (timm is not installed; definitions follow timm's public ones)."""
    timm = types.ModuleType("timm")
    layers = types.ModuleType("timm.layers")

    class DropPath(nn.Module):
        def __init__(self, drop_prob: float = 0.0, scale_by_keep: bool = True):
            super().__init__()
            self.drop_prob, self.scale_by_keep = drop_prob, scale_by_keep

        def forward(self, x):
            if self.drop_prob == 0.0 or not self.training:
                return x
            keep = 1 - self.drop_prob
            shape = (x.shape[0],) + (1,) * (x.ndim - 1)
            mask = x.new_empty(shape).bernoulli_(keep)
            if keep > 0.0 and self.scale_by_keep:
                mask.div_(keep)
            return x * mask

    class Mlp(nn.Module):
        def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, norm_layer=None,
                     bias=True, drop=0.0, use_conv=False):
            super().__init__()
            out_features = out_features or in_features
            hidden_features = hidden_features or in_features
            drops = drop if isinstance(drop, tuple) else (drop, drop)
            self.fc1 = nn.Linear(in_features, hidden_features, bias=bias)
            self.act = act_layer()
            self.drop1 = nn.Dropout(drops[0])
            self.norm = norm_layer(hidden_features) if norm_layer is not None else nn.Identity()
            self.fc2 = nn.Linear(hidden_features, out_features, bias=bias)
            self.drop2 = nn.Dropout(drops[1])

        def forward(self, x):
            return self.drop2(self.fc2(self.norm(self.drop1(self.act(self.fc1(x))))))

    layers.DropPath, layers.Mlp, layers.trunc_normal_ = DropPath, Mlp, nn.init.trunc_normal_
    timm.layers = layers
    sys.modules["timm"], sys.modules["timm.layers"] = timm, layers
    # import sam3.model.vitdet without executing sam3/__init__.py (which pulls the whole model zoo)
    for name, path in (("sam3", REF / "sam3"), ("sam3.model", REF / "sam3" / "model")):
        pkg = types.ModuleType(name)
        pkg.__path__ = [str(path)]
        sys.modules[name] = pkg


class SlicedQKV(nn.Module):
    """Fused-qkv replacement built only from reference classes: three nn.Linear row-slices of the
    original qkv, each wrapped in the reference's LoRALinear, concatenated back."""

    def __init__(self, qkv: nn.Linear, LoRALinear, rank, alpha):
        super().__init__()
        D = qkv.in_features
        mods = []
        for i in range(3):
            lin = nn.Linear(D, D, bias=True)
            lin.weight.data.copy_(qkv.weight.data[i * D:(i + 1) * D])
            lin.bias.data.copy_(qkv.bias.data[i * D:(i + 1) * D])
            mods.append(LoRALinear(lin, rank=rank, alpha=alpha, dropout=0.0))
        self.q_proj, self.k_proj, self.v_proj = mods

    def forward(self, x):
        return torch.cat([self.q_proj(x), self.k_proj(x), self.v_proj(x)], dim=-1)


def main():
    install_stubs()
    sys.path.insert(0, str(REF))
    from sam3.model.vitdet import ViT  # noqa: PLC0415  (reference)
    import lora_layers as ref_lora  # noqa: PLC0415  (reference)

    from oracle.vit_oracle import LoRASpec, ViTConfig, lora_keys, make_params  # noqa: PLC0415

    cfg = ViTConfig(img_size=224, patch_size=14, embed_dim=128, depth=2, num_heads=2, mlp_hidden=608, window_size=8,
                    global_att_blocks=(1,), pretrain_img_size=112)
    spec = LoRASpec(rank=4, alpha=8.0)
    params = make_params(cfg, spec, seed=1234)
    params = {k: v.half().float() for k, v in params.items()}  # fp16-exact values -> compact fixture
    g = torch.Generator().manual_seed(99)
    img = torch.randn(1, 3, 224, 224, generator=g).half().float()
    gout = torch.randn(1, 128, 16, 16, generator=g).half().float()

    torch.manual_seed(0)
    vit = ViT(img_size=224, pretrain_img_size=112, patch_size=14, embed_dim=128, depth=2, num_heads=2,
              mlp_ratio=4.75, norm_layer="LayerNorm", drop_path_rate=0.0, qkv_bias=True, use_abs_pos=True,
              tile_abs_pos=True, global_att_blocks=(1,), rel_pos_blocks=(), use_rope=True, use_interp_rope=True,
              window_size=8, pretrain_use_cls_token=True, retain_cls_token=False, ln_pre=True, ln_post=False,
              return_interm_layers=False, bias_patch_embed=False, use_act_checkpoint=False)
    base = {k: v for k, v in params.items() if ".lora." not in k}
    missing, unexpected = vit.load_state_dict(base, strict=False)
    assert not unexpected, unexpected
    assert all(k.endswith("freqs_cis") for k in missing), missing

    # adapters: fc1/fc2 through the reference's own injection
    cfg_l = ref_lora.LoRAConfig(rank=spec.rank, alpha=spec.alpha, dropout=0.0, target_modules=["fc1", "fc2"])
    ref_lora.apply_lora_to_model(vit, cfg_l)
    # q/k/v/out: reference LoRALinear composed over the fused qkv / proj
    for blk in vit.blocks:
        blk.attn.qkv = SlicedQKV(blk.attn.qkv, ref_lora.LoRALinear, spec.rank, spec.alpha)
        blk.attn.proj = ref_lora.LoRALinear(blk.attn.proj, rank=spec.rank, alpha=spec.alpha, dropout=0.0)
    # load the adapter values (same key scheme as save_lora_weights: "{module}.lora.lora_A")
    name_map = {}
    for i in range(cfg.depth):
        for t in ("q_proj", "k_proj", "v_proj"):
            name_map[f"blocks.{i}.attn.{t}"] = f"blocks.{i}.attn.qkv.{t}"
        name_map[f"blocks.{i}.attn.out_proj"] = f"blocks.{i}.attn.proj"
        name_map[f"blocks.{i}.mlp.fc1"] = f"blocks.{i}.mlp.fc1"
        name_map[f"blocks.{i}.mlp.fc2"] = f"blocks.{i}.mlp.fc2"
    mods = dict(vit.named_modules())
    for ours, theirs in name_map.items():
        lay = mods[theirs].lora
        lay.lora_A.data.copy_(params[f"{ours}.lora.lora_A"])
        lay.lora_B.data.copy_(params[f"{ours}.lora.lora_B"])
    assert ref_lora.count_parameters(vit)["trainable_parameters"] == sum(params[k].numel() for k in lora_keys(params))

    vit.eval()  # drop_path / dropout identity, no activation checkpointing
    blocks_out = []
    hooks = [b.register_forward_hook(lambda m, i, o: blocks_out.append(o.detach().clone())) for b in vit.blocks]
    pre = []
    hooks.append(vit.ln_pre.register_forward_hook(lambda m, i, o: pre.append(o.detach().clone())))
    out = vit(img)[0]
    (out * gout).sum().backward()
    for h in hooks:
        h.remove()

    # the reference's own serializer defines the on-disk adapter layout
    tmp = Path("/tmp/_golden_lora.pt")
    ref_lora.save_lora_weights(vit, str(tmp))
    saved = torch.load(tmp, weights_only=False)
    saved_keys = sorted(saved.keys())
    saved_shapes = {k: tuple(v.shape) for k, v in saved.items()}

    out_npz = {}
    for k, v in params.items():
        out_npz["param:" + k] = v.numpy().astype(np.float16)
    out_npz["img"] = img.numpy().astype(np.float16)
    out_npz["gout"] = gout.numpy().astype(np.float16)
    out_npz["out"] = out.detach().numpy()
    out_npz["ln_pre_out"] = pre[0].numpy()
    for i, b in enumerate(blocks_out):
        out_npz[f"block{i}_out"] = b.numpy()
    for ours, theirs in name_map.items():
        lay = mods[theirs].lora
        out_npz[f"grad:{ours}.lora.lora_A"] = lay.lora_A.grad.numpy()
        out_npz[f"grad:{ours}.lora.lora_B"] = lay.lora_B.grad.numpy()
    out_npz["ref_saved_keys"] = np.array(saved_keys)
    out_npz["ref_saved_shapes"] = np.array([str(saved_shapes[k]) for k in saved_keys])
    dst = ROOT / "tests" / "golden" / "vit_small.npz"
    np.savez_compressed(dst, **out_npz)
    print(f"wrote {dst} ({dst.stat().st_size / 1e6:.2f} MB); out rms {out.pow(2).mean().sqrt().item():.4f}")

    # ---- LoRALinear micro-golden (reference lora_layers.py:58-91 forward + autograd) ------------
    torch.manual_seed(7)
    lin = nn.Linear(96, 80)
    ll = ref_lora.LoRALinear(lin, rank=8, alpha=16, dropout=0.0)
    ll.lora.lora_B.data.normal_(0, 0.05)
    x = torch.randn(5, 96)
    y = ll(x)
    gy = torch.randn_like(y)
    xg = x.clone().requires_grad_(True)
    yy = ll(xg)
    (yy * gy).sum().backward()
    np.savez_compressed(ROOT / "tests" / "golden" / "lora_linear.npz", W=lin.weight.detach().numpy(),
                        b=lin.bias.detach().numpy(), A=ll.lora.lora_A.detach().numpy(), B=ll.lora.lora_B.detach().numpy(),
                        x=x.numpy(), y=y.detach().numpy(), gy=gy.numpy(), dx=xg.grad.numpy(),
                        dA=ll.lora.lora_A.grad.numpy(), dB=ll.lora.lora_B.grad.numpy(), scaling=np.float32(ll.lora.scaling))
    print("wrote lora_linear.npz")


if __name__ == "__main__":
    main()
