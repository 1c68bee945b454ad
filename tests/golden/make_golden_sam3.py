"""Generates tests/golden/sam3_step.npz by running the REAL reference detector end to end (row a9).

Run in the build container only (needs the reference installed under baseline/_ref, tools/install_reference.sh):

    python tests/golden/make_golden_sam3.py            # ~5 min of CPU

Executed from the reference, unmodified: `build_sam3_image_model` (sam3/model_builder.py:558-641, load_from_HF=False),
`Sam3Image.forward` in train() mode (sam3/model/sam3_image.py:442-576), `collate_fn_api`, the reference's own
`lora_layers.apply_lora_to_model` (fc1/fc2 of the 32 trunk blocks = what the shipped configs adapt on this model), and —
for the objective — `Sam3LossWrapper` + `Boxes` / `IABCEMdetr` / `Masks` + `BinaryHungarianMatcherV2` (SciPy) exactly as
train_sam3_lora_native.py:743-793, 898-931 wires them (eager focal loss: the Triton kernels need a GPU).

Weights are synthetic and seeded per parameter NAME (sam3_bridge.seed_parameters; no checkpoint is reachable offline), the
batch is sam3_step.synthetic_datapoints(1, seed=0): both are regenerated bit-identically by the GPU test, so only outputs
are stored.  Every random op is off (sam3_step.disable_stochastic).  Stored (fp32):
  * pred_logits / pred_boxes / presence_logit_dec / semantic_seg, pred_masks[:, SEL] and pred_masks_o2m[:, SEL] of the
    final decoder step, a strided slice of the trunk feature map;
  * gradients of the LINEAR functional  sum(pred_masks * Cm) + sum(pred_logits * Cl) + sum(pred_boxes * Cb)  (seeded
    cotangents; BASELINE.md §3 variant ii) w.r.t. every adapter: all 128 norms, full tensors for blocks 0 and 31;
  * the training objective's loss dict (discrete matching makes it informational: compared only when indices agree).
"""
from __future__ import annotations

import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from sam3_lora_b200 import sam3_bridge as bridge  # noqa: E402
from sam3_lora_b200 import sam3_step as step  # noqa: E402

SEL = [0, 25, 50, 75, 100, 125, 150, 199]
RANK, ALPHA = 16, 32
TARGETS = ["q_proj", "k_proj", "v_proj", "out_proj", "fc1", "fc2"]


def seed_adapters(model, seed=1):
    """A ~ U(-1/sqrt(in), 1/sqrt(in)), B ~ N(0, 0.02) (non-zero so the adapter path carries signal), keyed by name."""
    import zlib

    with torch.no_grad():
        for name, p in model.named_parameters():
            if ".lora." not in name:
                continue
            g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(name.encode())) & 0x7FFFFFFF)
            if name.endswith("lora_A"):
                p.copy_((torch.rand(p.shape, generator=g) * 2 - 1) / p.shape[0] ** 0.5)
            else:
                p.copy_(torch.randn(p.shape, generator=g) * 0.02)


def cotangents(out, seed=7):
    g = torch.Generator().manual_seed(seed)
    return {k: torch.randn(out[k].shape, generator=g) for k in ("pred_masks", "pred_logits", "pred_boxes")}


def main():
    t0 = time.time()
    bridge.import_reference()
    import lora_layers as ref_lora  # the reference's root module (baseline/_ref/lora_layers.py)

    assert "baseline/_ref" in ref_lora.__file__ or "reference" in ref_lora.__file__, ref_lora.__file__
    model = bridge.build_reference_model("cpu", seed=0)
    cfg = ref_lora.LoRAConfig(rank=RANK, alpha=ALPHA, dropout=0.0, target_modules=TARGETS, apply_to_vision_encoder=True,
                              apply_to_text_encoder=False, apply_to_geometry_encoder=False, apply_to_detr_encoder=False,
                              apply_to_detr_decoder=False, apply_to_mask_decoder=False)
    model = ref_lora.apply_lora_to_model(model, cfg)
    seed_adapters(model)
    model.train()
    step.disable_stochastic(model)
    batch = step.collate(step.synthetic_datapoints(1, seed=0))
    out = {}
    feats = {}
    trunk = model.backbone.vision_backbone.trunk
    hook = trunk.register_forward_hook(lambda m, i, o: feats.__setitem__("f", o[-1].detach()))
    with bridge.cpu_compat():
        outputs_list = model(batch)
    hook.remove()
    fin = step.final_outputs(outputs_list)
    print("forward done", time.time() - t0)
    out["trunk_slice"] = feats["f"][0, ::64, ::6, ::6].numpy().copy()
    out["trunk_rms"] = np.float32(feats["f"].pow(2).mean().sqrt().item())
    for k in ("pred_logits", "pred_boxes", "presence_logit_dec", "semantic_seg"):
        out[k] = fin[k].detach().numpy().copy()
    out["pred_masks_sel"] = fin["pred_masks"][:, SEL].detach().numpy().copy()
    out["pred_masks_o2m_sel"] = fin["pred_masks_o2m"][:, SEL].detach().numpy().copy()
    out["pred_masks_rms"] = np.float32(fin["pred_masks"].pow(2).mean().sqrt().item())
    out["sel"] = np.asarray(SEL)
    cot = cotangents(fin)
    lin = sum((fin[k] * cot[k]).sum() for k in cot)
    out["linear_value"] = np.float64(lin.item())
    lin.backward()
    print("backward done", time.time() - t0)
    names, norms = [], []
    for name, p in model.named_parameters():
        if ".lora." in name:
            names.append(name)
            norms.append(p.grad.norm().item())
            if ".blocks.0." in name or ".blocks.31." in name:
                out["grad." + name] = p.grad.numpy().copy()
            p.grad = None
    out["grad_names"] = np.asarray(names)
    out["grad_norms"] = np.asarray(norms, dtype=np.float64)
    # the objective, as the trainer computes it
    matcher, wrapper = step.build_objective(native=False)
    with bridge.cpu_compat():
        loss, loss_dict = step.training_loss(model, batch, matcher, wrapper)
    print("objective done", time.time() - t0, float(loss))
    out["loss_names"] = np.asarray(sorted(k for k, v in loss_dict.items() if isinstance(v, torch.Tensor) and v.numel() == 1))
    out["loss_values"] = np.asarray([float(loss_dict[k]) for k in out["loss_names"]], dtype=np.float64)
    dst = Path(__file__).with_name("sam3_step.npz")
    np.savez_compressed(dst, **out)
    print(f"wrote {dst} ({dst.stat().st_size / 1e6:.2f} MB) in {time.time() - t0:.0f} s")


if __name__ == "__main__":
    main()
