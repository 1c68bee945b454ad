"""The CPU oracle is pinned against vectors produced by the real reference code
(tests/golden/make_golden.py ran sam3/model/vitdet.py `ViT` + lora_layers.py `LoRALinear`)."""
import numpy as np
import torch

from oracle import vit_oracle as O
from tests.helpers import GOLDEN, load_small_golden, rel_max


def test_oracle_forward_matches_reference_vit():
    g = load_small_golden()
    out, blocks = O.vit_forward(g["img"], g["params"], g["cfg"], g["spec"].scaling, return_blocks=True)
    assert rel_max(blocks[0], g["ln_pre_out"]) < 2e-6          # patch embed + tiled abs pos + ln_pre
    assert rel_max(blocks[1], g["blocks"][0]) < 5e-6           # window block (9 -> 4 windows of 64 tokens)
    assert rel_max(blocks[2], g["blocks"][1]) < 5e-6           # global block (interpolated rope)
    assert rel_max(out, g["out"]) < 5e-6


def test_oracle_lora_grads_match_reference_autograd():
    g = load_small_golden()
    out, grads = O.train_step_reference(g["img"], g["params"], g["cfg"], g["spec"], g["gout"])
    assert set(grads) == set(g["grads"])
    for k, ref in g["grads"].items():
        assert grads[k].shape == ref.shape
        assert rel_max(grads[k], ref) < 2e-5, k


def test_oracle_fp64_agrees_with_fp32():
    g = load_small_golden()
    p64 = {k: v.double() for k, v in g["params"].items()}
    out64 = O.vit_forward(g["img"].double(), p64, g["cfg"], g["spec"].scaling)
    assert rel_max(out64.float(), g["out"]) < 5e-6


def test_lora_delta_matches_reference_loralinear():
    z = np.load(GOLDEN / "lora_linear.npz")
    t = {k: torch.from_numpy(z[k]) for k in z.files if k != "scaling"}
    s = float(z["scaling"])
    x = t["x"].clone().requires_grad_(True)
    A = t["A"].clone().requires_grad_(True)
    B = t["B"].clone().requires_grad_(True)
    y = x @ t["W"].T + t["b"] + O.lora_delta(x, A, B, s)
    assert rel_max(y, t["y"]) < 1e-6
    (y * t["gy"]).sum().backward()
    assert rel_max(x.grad, t["dx"]) < 1e-5
    assert rel_max(A.grad, t["dA"]) < 1e-5
    assert rel_max(B.grad, t["dB"]) < 1e-5


def test_rope_angles_match_reference_buffer_shape_and_scale():
    cfg = O.ViTConfig()
    w = O.rope_angles_for_block(cfg, is_global=False)
    gl = O.rope_angles_for_block(cfg, is_global=True)
    assert w.shape == (576, 32) and gl.shape == (5184, 32)
    # global rope is the window rope sampled at 1/3 positions: token (row 3, col 6) == window (1, 2)
    assert torch.allclose(gl[3 * 72 + 6], w[1 * 24 + 2], atol=1e-12)


def test_oracle_drop_path_matches_timm_semantics():
    """x + DropPath(branch): per-sample mask/keep on the branch only (timm DropPath, used at vitdet.py:610-611)."""
    g = load_small_golden()
    cfg = g["cfg"]
    img2 = torch.cat([g["img"], g["img"].flip(-1)], dim=0)
    full = O.vit_forward(img2, g["params"], cfg, g["spec"].scaling, return_blocks=True)[1]
    drop = torch.ones(cfg.depth, 2, 2)
    drop[0, 0, 1] = 0.0          # sample 1 skips block 0's attention branch
    drop[1, 1, 0] = 1.0 / 0.9    # sample 0 keeps block 1's MLP branch, scaled by 1/keep
    out, blocks = O.vit_forward(img2, g["params"], cfg, g["spec"].scaling, return_blocks=True, drop_scales=drop)
    # sample 0, block 0 untouched
    assert torch.equal(blocks[1][0], full[1][0])
    # sample 1, block 0: attention branch removed -> x_mid == x_in, then MLP on it
    x_in = blocks[0][1:2]
    D = cfg.embed_dim
    p = g["params"]
    h = torch.nn.functional.layer_norm(x_in, (D,), p["blocks.0.norm2.weight"], p["blocks.0.norm2.bias"], cfg.ln_eps)
    h1 = h @ p["blocks.0.mlp.fc1.weight"].T + p["blocks.0.mlp.fc1.bias"] + O.lora_delta(h, p["blocks.0.mlp.fc1.lora.lora_A"], p["blocks.0.mlp.fc1.lora.lora_B"], g["spec"].scaling)
    gl = torch.nn.functional.gelu(h1)
    h2 = gl @ p["blocks.0.mlp.fc2.weight"].T + p["blocks.0.mlp.fc2.bias"] + O.lora_delta(gl, p["blocks.0.mlp.fc2.lora.lora_A"], p["blocks.0.mlp.fc2.lora.lora_B"], g["spec"].scaling)
    assert rel_max(blocks[1][1:2], x_in + h2) < 1e-6
