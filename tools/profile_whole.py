"""torch.profiler summary of one whole-detector training step on the swapped model (GPU): top CUDA kernels / ops by time
and the forward time of each top-level sub-module (CUDA events around forward hooks).  Context for where the non-trunk time goes."""
import argparse
import contextlib
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from sam3_lora_b200 import sam3_bridge, sam3_step  # noqa: E402
from sam3_lora_b200.lora_layers import LoRAConfig, apply_lora_to_model  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--out", default="gpurun_out/profile_whole.txt")
    a = ap.parse_args()
    B = a.batch
    ref = sam3_bridge.build_reference_model("cpu", seed=0)
    model = sam3_bridge.build_native_model("cuda", reference_model=ref, max_batch=B)
    with contextlib.redirect_stdout(sys.stderr):
        model = apply_lora_to_model(model, LoRAConfig(rank=16, alpha=32, target_modules=["fc1", "fc2"], apply_to_text_encoder=False,
                                                      apply_to_detr_encoder=False, apply_to_detr_decoder=False,
                                                      strict_reference_names=True)).cuda()
    params = [p for p in model.parameters() if p.requires_grad]
    for p in params:
        if p.shape[0] == 16:
            torch.nn.init.normal_(p, std=0.02)
    model.train()
    matcher, wrapper = sam3_step.build_objective(native=True)
    opt = torch.optim.AdamW(params, lr=5e-5, fused=True)
    data = sam3_step.move_to_device(sam3_step.collate(sam3_step.synthetic_datapoints(B, seed=0)), "cuda")

    names = ["backbone.vision_backbone.trunk", "backbone.vision_backbone", "backbone.language_backbone", "geometry_encoder",
             "transformer.encoder", "transformer.decoder", "segmentation_head", "dot_prod_scoring"]
    ev = {}

    def pre(n):
        def f(m, i):
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            ev.setdefault(n, []).append([e, None])
        return f

    def post(n):
        def f(m, i, o):
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            ev[n][-1][1] = e
        return f

    hooks = []
    for n in names:
        m = model.get_submodule(n)
        hooks += [m.register_forward_pre_hook(pre(n)), m.register_forward_hook(post(n))]

    phases = {}

    def step(record=False):
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        marks[0].record()
        outputs = None
        with torch.profiler.record_function("phase:forward+matcher+loss"):
            loss, _ = sam3_step.training_loss(model, data, matcher, wrapper)
        marks[1].record()
        opt.zero_grad(set_to_none=True)
        with torch.profiler.record_function("phase:backward"):
            loss.backward()
        marks[2].record()
        with torch.profiler.record_function("phase:adamw"):
            opt.step()
        marks[3].record()
        torch.cuda.synchronize()
        if record:
            phases["forward+matcher+loss_ms"] = marks[0].elapsed_time(marks[1])
            phases["backward_ms"] = marks[1].elapsed_time(marks[2])
            phases["adamw_ms"] = marks[2].elapsed_time(marks[3])

    for _ in range(3):
        ev.clear()
        step()
    ev.clear()
    step(record=True)
    mods = {n: sum(a_.elapsed_time(b_) for a_, b_ in v if b_ is not None) for n, v in ev.items()}
    for h in hooks:
        h.remove()
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CPU, torch.profiler.ProfilerActivity.CUDA]) as prof:
        step()
    txt = prof.key_averages().table(sort_by="cuda_time_total", row_limit=60, max_name_column_width=90)
    Path(a.out).parent.mkdir(exist_ok=True)
    with open(a.out, "w") as f:
        f.write(json.dumps({"batch": B, "phases": phases, "module_forward_ms (incl. recompute calls during backward)": mods}, indent=1) + "\n\n")
        f.write(txt)
    print(json.dumps({"phases": phases, "modules": mods}, indent=1))


if __name__ == "__main__":
    main()
