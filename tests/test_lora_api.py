"""Host-side LoRA surface (no GPU): same API, freezing rule, name matching, component gates and
state-dict layout as the reference's lora_layers.py (golden keys/shapes come from the reference's
own save_lora_weights, recorded in tests/golden/vit_small.npz)."""
import math

import pytest
import torch
import torch.nn as nn

from sam3_lora_b200 import lora_layers as LL
from sam3_lora_b200.vit import ViT
from tests.helpers import load_small_golden


class Toy(nn.Module):
    def __init__(self):
        super().__init__()
        self.vision_backbone = nn.ModuleDict({"fc1": nn.Linear(8, 16), "q_proj": nn.Linear(8, 8), "out_proj": nn.Linear(8, 8)})
        self.transformer = nn.ModuleDict({"encoder": nn.ModuleDict({"q_proj": nn.Linear(8, 8)}),
                                          "decoder": nn.ModuleDict({"v_proj": nn.Linear(8, 8)})})
        self.segmentation_head = nn.ModuleDict({"q_proj": nn.Linear(8, 8)})
        self.other = nn.Linear(8, 8)


def test_config_defaults_and_to_dict_keys_match_reference():
    c = LL.LoRAConfig()
    assert (c.rank, c.alpha, c.dropout) == (8, 16, 0.0)
    assert c.target_modules == {"q_proj", "k_proj", "v_proj", "out_proj"}
    assert list(c.to_dict().keys()) == ["rank", "alpha", "dropout", "target_modules", "apply_to_vision_encoder",
                                        "apply_to_text_encoder", "apply_to_geometry_encoder", "apply_to_detr_encoder",
                                        "apply_to_detr_decoder", "apply_to_mask_decoder"]
    assert c.apply_to_geometry_encoder is False and c.apply_to_mask_decoder is False and c.apply_to_vision_encoder is True


def test_lora_layer_init_shapes_and_scaling():
    torch.manual_seed(0)
    l = LL.LoRALayer(64, 32, rank=4, alpha=8, dropout=0.1)
    assert l.lora_A.shape == (64, 4) and l.lora_B.shape == (4, 32)
    assert l.scaling == 2.0 and isinstance(l.dropout, nn.Dropout)
    assert torch.count_nonzero(l.lora_B) == 0
    # kaiming_uniform_(a=sqrt(5)) on an [in, r] tensor: fan_in = r -> bound = 1/sqrt(r)
    assert l.lora_A.abs().max() <= 1 / math.sqrt(4) + 1e-6 and l.lora_A.std() > 0.1
    assert isinstance(LL.LoRALayer(8, 8).dropout, nn.Identity)


def test_apply_freezes_everything_and_wraps_by_basename():
    m = Toy()
    LL.apply_lora_to_model(m, LL.LoRAConfig(rank=2, alpha=4, target_modules=["q_proj", "v_proj", "out_proj", "fc1"]))
    assert isinstance(m.vision_backbone["fc1"], LL.LoRALinear)
    assert isinstance(m.vision_backbone["q_proj"], LL.LoRALinear)
    assert isinstance(m.vision_backbone["out_proj"], nn.Linear)       # out_proj is never wrapped (lora_layers.py:194-196)
    assert isinstance(m.transformer["encoder"]["q_proj"], LL.LoRALinear)
    assert isinstance(m.transformer["decoder"]["v_proj"], LL.LoRALinear)
    assert isinstance(m.segmentation_head["q_proj"], nn.Linear)       # mask-decoder gate is off by default
    assert isinstance(m.other, nn.Linear)
    trainable = [n for n, p in m.named_parameters() if p.requires_grad]
    assert trainable and all(".lora.lora_" in n for n in trainable)
    stats = LL.count_parameters(m)
    assert stats["trainable_parameters"] == sum(p.numel() for p in LL.get_lora_parameters(m))
    assert set(stats) == {"total_parameters", "trainable_parameters", "trainable_percentage"}


def test_component_gates():
    m = Toy()
    LL.apply_lora_to_model(m, LL.LoRAConfig(rank=2, target_modules=["q_proj", "v_proj", "fc1"], apply_to_vision_encoder=False,
                                            apply_to_detr_decoder=False, apply_to_mask_decoder=True))
    assert isinstance(m.vision_backbone["fc1"], nn.Linear) and isinstance(m.vision_backbone["q_proj"], nn.Linear)
    assert isinstance(m.transformer["decoder"]["v_proj"], nn.Linear)
    assert isinstance(m.transformer["encoder"]["q_proj"], LL.LoRALinear)
    assert isinstance(m.segmentation_head["q_proj"], LL.LoRALinear)   # alias of the mask_decoder gate
    m2 = Toy()
    LL.apply_lora_to_model(m2, LL.LoRAConfig(rank=2, target_modules=["q_proj"], apply_to_mask_decoder=False,
                                             strict_reference_names=True))
    assert isinstance(m2.segmentation_head["q_proj"], LL.LoRALinear)  # strict mode: reference gate sees no "mask_decoder" substring


def _small_vit():
    return ViT(img_size=224, embed_dim=128, depth=2, num_heads=2, mlp_ratio=4.75, window_size=8, global_att_blocks=(1,),
               pretrain_img_size=112)


def test_vit_state_dict_layout_matches_reference_save_format():
    g = load_small_golden()
    v = _small_vit()
    LL.apply_lora_to_model(v, LL.LoRAConfig(rank=4, alpha=8, target_modules=["q_proj", "k_proj", "v_proj", "out_proj", "fc1", "fc2"]))
    ours = LL.lora_state_dict(v)
    # golden: keys/shapes written by the reference's save_lora_weights on the reference-built model, where the
    # sliced projections live at attn.qkv.{q,k,v}_proj and attn.proj (make_golden.py); map to the virtual names
    ref = {}
    for k, shp in zip(g["ref_saved_keys"], g["ref_saved_shapes"]):
        k = k.replace(".attn.qkv.", ".attn.").replace(".attn.proj.", ".attn.out_proj.")
        ref[k] = eval(shp)  # noqa: S307 - literal tuple written by make_golden.py
    assert set(ours) == set(ref)
    for k, t in ours.items():
        assert tuple(t.shape) == ref[k] and t.dtype == torch.float32
    # mlp keys are byte-for-byte the reference's own names
    assert "blocks.0.mlp.fc1.lora.lora_A" in ours and ours["blocks.0.mlp.fc1.lora.lora_A"].shape == (128, 4)
    assert ours["blocks.0.mlp.fc1.lora.lora_B"].shape == (4, 608)
    sd = v.state_dict()
    assert "blocks.0.mlp.fc1.original_layer.weight" in sd     # frozen base moves under original_layer (lora_layers.py:74)
    assert "blocks.0.attn.qkv.weight" in sd and "blocks.0.attn.freqs_cis" in sd
    assert sd["blocks.1.attn.freqs_cis"].shape == (256, 32) and sd["blocks.0.attn.freqs_cis"].dtype == torch.complex64


def test_save_load_roundtrip(tmp_path):
    v = _small_vit()
    cfg = LL.LoRAConfig(rank=4, alpha=8, target_modules=["q_proj", "v_proj", "fc2"])
    LL.apply_lora_to_model(v, cfg)
    for p in LL.get_lora_parameters(v):
        nn.init.normal_(p)
    path = tmp_path / "last_lora_weights.pt"
    LL.save_lora_weights(v, str(path))
    blob = torch.load(path)
    assert all(isinstance(t, torch.Tensor) and t.dtype == torch.float32 for t in blob.values())
    v2 = _small_vit()
    LL.apply_lora_to_model(v2, cfg)
    LL.load_lora_weights(v2, str(path))
    a, b = LL.lora_state_dict(v), LL.lora_state_dict(v2)
    assert set(a) == set(b) and all(torch.equal(a[k], b[k]) for k in a)


def test_strict_reference_names_reproduces_readme_census():
    """README.md:225-226: full_lora_config on the native model -> 'Applied LoRA to 64 modules',
    11,796,480 trainable parameters (fc1/fc2 of the 32 ViT blocks, r=32) — SURVEY.md fact 5."""
    with torch.device("meta"):
        trunk = ViT()
    root = nn.Module()
    root.backbone = nn.Module()
    root.backbone.vision_backbone = nn.Module()
    root.backbone.vision_backbone.trunk = trunk
    cfg = LL.LoRAConfig(rank=32, alpha=64, dropout=0.1, target_modules=["q_proj", "k_proj", "v_proj", "out_proj", "fc1", "fc2"],
                        apply_to_geometry_encoder=True, apply_to_mask_decoder=True, strict_reference_names=True)
    LL.apply_lora_to_model(root, cfg)
    n_wrapped = sum(isinstance(m, LL.LoRALinear) for m in root.modules())
    assert n_wrapped == 64
    assert LL.count_parameters(root)["trainable_parameters"] == 11_796_480
    # north-star aliasing adds the q/k/v/out adapters of the fused projections
    with torch.device("meta"):
        trunk2 = ViT()
    LL.apply_lora_to_model(trunk2, LL.LoRAConfig(rank=16, alpha=32, target_modules=["q_proj", "k_proj", "v_proj", "out_proj"]))
    assert LL.count_parameters(trunk2)["trainable_parameters"] == 32 * 4 * 2 * 1024 * 16   # 4.19 M


def test_fused_mha_gets_q_k_v_virtual_and_wrapped_out_proj():
    from sam3_lora_b200.mha import MultiheadAttention, replace_torch_mha

    root = nn.Module()
    root.transformer = nn.Module()
    root.transformer.encoder = nn.Module()
    root.transformer.encoder.self_attn = nn.MultiheadAttention(256, 8, dropout=0.1)
    assert replace_torch_mha(root) == 1
    m = root.transformer.encoder.self_attn
    assert isinstance(m, MultiheadAttention) and m.dropout == 0.1 and not m.batch_first
    keys = set(m.state_dict())
    assert {"in_proj_weight", "in_proj_bias", "out_proj.weight", "out_proj.bias"} <= keys
    LL.apply_lora_to_model(root, LL.LoRAConfig(rank=8, alpha=16, target_modules=["q_proj", "k_proj", "v_proj", "out_proj"]))
    sd = LL.lora_state_dict(root)
    pre = "transformer.encoder.self_attn."
    assert set(sd) == {pre + f"{n}.lora.lora_{ab}" for n in ("q_proj", "k_proj", "v_proj", "out_proj") for ab in "AB"}
    assert sd[pre + "q_proj.lora.lora_A"].shape == (256, 8) and sd[pre + "out_proj.lora.lora_B"].shape == (8, 256)
    assert isinstance(m.out_proj, LL.LoRALinear)
    # strict mode keeps the reference's behaviour: nothing on a packed-in_proj attention, out_proj never wrapped
    root2 = nn.Module()
    root2.self_attn = nn.MultiheadAttention(256, 8)
    replace_torch_mha(root2)
    LL.apply_lora_to_model(root2, LL.LoRAConfig(rank=8, target_modules=["q_proj", "out_proj"], strict_reference_names=True))
    assert LL.lora_state_dict(root2) == {}
    # detr gate off -> untouched
    root3 = nn.Module()
    root3.transformer = nn.Module()
    root3.transformer.decoder = nn.Module()
    root3.transformer.decoder.cross_attn = MultiheadAttention(256, 8)
    LL.apply_lora_to_model(root3, LL.LoRAConfig(rank=8, target_modules=["q_proj"], apply_to_detr_decoder=False))
    assert LL.lora_state_dict(root3) == {}
