"""Row a9 on the GPU: the reference's `Sam3Image` with the hot-path modules swapped for the native ones
(sam3_bridge.build_native_model) against tests/golden/sam3_step.npz, which the UNMODIFIED reference produced on the CPU in
fp32 (tests/golden/make_golden_sam3.py).  Weights and the batch are regenerated from their seeds; every random op is off.

The detector's wiring is the reference's own Python, imported from baseline/_ref (it travels to the GPU box); the test is
skipped when it is absent.  Tolerances are written next to each assert; all measured numbers go to parity_report.jsonl.
"""
from __future__ import annotations

from pathlib import Path

import numpy as np
import pytest
import torch

from tests.helpers import GOLDEN, rel_l2, rel_max
from tests.test_vit_engine_gpu import _report

pytestmark = pytest.mark.gpu

bridge = pytest.importorskip("sam3_lora_b200.sam3_bridge")
if bridge.reference_root() is None:
    pytest.skip("reference not installed under baseline/_ref (tools/install_reference.sh)", allow_module_level=True)
if not (GOLDEN / "sam3_step.npz").exists():
    pytest.skip("tests/golden/sam3_step.npz missing", allow_module_level=True)

from sam3_lora_b200 import sam3_step as step  # noqa: E402

RANK, ALPHA = 16, 32
TARGETS = ["q_proj", "k_proj", "v_proj", "out_proj", "fc1", "fc2"]


def _seed_adapters(model, seed=1):
    import zlib

    with torch.no_grad():
        for name, p in model.named_parameters():
            if ".lora." not in name:
                continue
            g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(name.encode())) & 0x7FFFFFFF)
            if name.endswith("lora_A"):
                p.copy_(((torch.rand(p.shape, generator=g) * 2 - 1) / p.shape[0] ** 0.5).to(p.device))
            else:
                p.copy_((torch.randn(p.shape, generator=g) * 0.02).to(p.device))


@pytest.fixture(scope="module")
def golden():
    z = np.load(GOLDEN / "sam3_step.npz")
    return {k: z[k] for k in z.files}


@pytest.fixture(scope="module")
def native_model():
    from sam3_lora_b200.lora_layers import LoRAConfig, apply_lora_to_model

    ref = bridge.build_reference_model("cpu", seed=0)
    model = bridge.build_native_model("cuda", reference_model=ref, max_batch=1)
    cfg = LoRAConfig(rank=RANK, alpha=ALPHA, dropout=0.0, target_modules=TARGETS, apply_to_vision_encoder=True,
                     apply_to_text_encoder=False, apply_to_geometry_encoder=False, apply_to_detr_encoder=False,
                     apply_to_detr_decoder=False, apply_to_mask_decoder=False, strict_reference_names=True)
    model = apply_lora_to_model(model, cfg).to("cuda")
    _seed_adapters(model)
    model.train()
    step.disable_stochastic(model)
    return model


def test_swapped_modules_are_native(native_model):
    from sam3_lora_b200.maskformer_segmentation import PixelDecoder
    from sam3_lora_b200.mha import MultiheadAttention
    from sam3_lora_b200.necks import Sam3DualViTDetNeck
    from sam3_lora_b200.vit import ViT

    m = native_model
    assert isinstance(m.backbone.vision_backbone, Sam3DualViTDetNeck)
    assert isinstance(m.backbone.vision_backbone.trunk, ViT)
    assert isinstance(m.segmentation_head.pixel_decoder, PixelDecoder)
    assert not any(isinstance(x, torch.nn.MultiheadAttention) for x in m.modules())
    assert sum(isinstance(x, MultiheadAttention) for x in m.modules()) >= 6 * 2 + 6 * 3 + 1


def test_forward_and_adapter_gradients_match_reference_golden(native_model, golden):
    """Mask logits / class logits / boxes of the final decoder step and the adapter gradients of the linear functional
    sum(pred_masks*Cm) + sum(pred_logits*Cl) + sum(pred_boxes*Cb) against the fp32 CPU reference.

    Bounds: trunk features rel-L2 <= 2e-3 (same bound as the trunk tests); mask logits rel-L2 <= 5e-3, rel-max <= 1e-2
    (the trunk error passes through the neck, 6 + 6 DETR layers with fp16-operand attention and the pixel decoder);
    class logits / boxes abs <= 5e-3; adapter gradients rel-L2 <= 2e-2 per tensor, <= 1e-2 median."""
    from tests.golden.make_golden_sam3 import cotangents  # same seeded cotangents as the generator

    model = native_model
    batch = step.move_to_device(step.collate(step.synthetic_datapoints(1, seed=0)), "cuda")
    feats = {}
    trunk = model.backbone.vision_backbone.trunk
    hook = trunk.register_forward_hook(lambda m, i, o: feats.__setitem__("f", o[-1].detach()))
    outputs_list = model(batch)
    hook.remove()
    fin = step.final_outputs(outputs_list)
    sel = golden["sel"].tolist()
    rep = {}
    rep["trunk_slice_rel_l2"] = rel_l2(feats["f"][0, ::64, ::6, ::6].cpu(), torch.from_numpy(golden["trunk_slice"]))
    for k in ("pred_logits", "pred_boxes", "presence_logit_dec", "semantic_seg"):
        got, ref = fin[k].detach().float().cpu(), torch.from_numpy(golden[k])
        rep[k + "_rel_l2"], rep[k + "_abs_max"] = rel_l2(got, ref), (got - ref).abs().max().item()
    for k, src in (("pred_masks_sel", "pred_masks"), ("pred_masks_o2m_sel", "pred_masks_o2m")):
        got, ref = fin[src][:, sel].detach().float().cpu(), torch.from_numpy(golden[k])
        rep[k + "_rel_l2"], rep[k + "_rel_max"] = rel_l2(got, ref), rel_max(got, ref)
    cot = cotangents({k: fin[k].detach().cpu() for k in ("pred_masks", "pred_logits", "pred_boxes")})
    lin = sum((fin[k] * cot[k].to("cuda")).sum() for k in cot)
    rep["linear_value_rel"] = abs(lin.item() - float(golden["linear_value"])) / abs(float(golden["linear_value"]))
    lin.backward()
    names = [str(n) for n in golden["grad_names"]]
    got_norm = {}
    full = {}
    for name, p in model.named_parameters():
        if ".lora." in name:
            assert p.grad is not None, name
            got_norm[name] = p.grad.norm().item()
            if "grad." + name in golden:
                full[name] = rel_l2(p.grad.detach().float().cpu(), torch.from_numpy(golden["grad." + name]))
    assert set(got_norm) == set(names)
    norm_err = [abs(got_norm[n] - r) / max(r, 1e-30) for n, r in zip(names, golden["grad_norms"])]
    rep["grad_norm_rel_max"], rep["grad_rel_l2_max"] = max(norm_err), max(full.values())
    rep["grad_rel_l2_median"] = sorted(full.values())[len(full) // 2]
    rep["grads_full"] = full
    _report("sam3_step_a9_vs_reference_golden", rep)
    assert rep["trunk_slice_rel_l2"] < 2e-3, rep
    assert rep["pred_masks_sel_rel_l2"] < 5e-3 and rep["pred_masks_sel_rel_max"] < 1e-2, rep
    assert rep["pred_masks_o2m_sel_rel_l2"] < 5e-3, rep
    assert rep["pred_logits_abs_max"] < 5e-3 and rep["pred_boxes_abs_max"] < 5e-3, rep
    assert rep["grad_rel_l2_max"] < 2e-2 and rep["grad_rel_l2_median"] < 1e-2, rep
    assert rep["grad_norm_rel_max"] < 2e-2, rep


def test_training_objective_runs_and_is_close_to_reference(native_model, golden):
    """The trainer's objective (GPU matcher + fused mask / focal losses inside the reference's Sam3LossWrapper) on the
    swapped model: finite, differentiable, and — matching being discrete — within 2 % of the reference's CPU value."""
    model = native_model
    for p in model.parameters():
        p.grad = None
    matcher, wrapper = step.build_objective(native=True)
    batch = step.move_to_device(step.collate(step.synthetic_datapoints(1, seed=0)), "cuda")
    loss, loss_dict = step.training_loss(model, batch, matcher, wrapper)
    assert torch.isfinite(loss)
    loss.backward()
    g = [p.grad for n, p in model.named_parameters() if ".lora." in n]
    assert all(x is not None and torch.isfinite(x).all() for x in g)
    assert sum(x.abs().sum().item() for x in g) > 0
    ref = dict(zip([str(n) for n in golden["loss_names"]], golden["loss_values"]))
    got = {k: float(v) for k, v in loss_dict.items() if isinstance(v, torch.Tensor) and v.numel() == 1 and k in ref}
    rel = {k: abs(got[k] - ref[k]) / max(abs(ref[k]), 1e-6) for k in got}
    from sam3.train.loss.loss_fns import CORE_LOSS_KEY

    _report("sam3_step_a9_objective", {"loss": float(loss), "ref_loss": ref.get(CORE_LOSS_KEY), "rel": rel})
    assert rel[CORE_LOSS_KEY] < 2e-2, (got, ref)
