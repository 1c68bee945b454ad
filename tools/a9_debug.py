"""Debug helper (GPU): run the swapped SAM3 detector once with forward hooks on every module and report the first module
whose output is not finite, with the magnitude of its inputs.  `--stochastic 0/1`, `--batch B`, `--parts trunk,neck,...`."""
import argparse
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from sam3_lora_b200 import sam3_bridge, sam3_step  # noqa: E402


def tensors(o):
    if isinstance(o, torch.Tensor):
        yield o
    elif isinstance(o, (list, tuple)):
        for x in o:
            yield from tensors(x)
    elif isinstance(o, dict):
        for x in o.values():
            yield from tensors(x)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--stochastic", type=int, default=1)
    ap.add_argument("--parts", default="trunk,neck,pixel_decoder,mha,matcher")
    ap.add_argument("--term", default="loss_bbox")
    a = ap.parse_args()
    ref = sam3_bridge.build_reference_model("cpu", seed=0)
    model = sam3_bridge.build_native_model("cuda", reference_model=ref, max_batch=a.batch, parts=[p for p in a.parts.split(",") if p])
    from sam3_lora_b200.lora_layers import LoRAConfig, apply_lora_to_model

    model = apply_lora_to_model(model, LoRAConfig(rank=8, alpha=16, target_modules=["fc1", "fc2"], strict_reference_names=True)).cuda()
    for p in model.parameters():
        if p.requires_grad and p.shape[0] == 8:
            torch.nn.init.normal_(p, std=0.02)
    model.train()
    if not a.stochastic:
        sam3_step.disable_stochastic(model)
    bad = []
    stats = []

    def hook(name):
        def fn(mod, inp, out):
            fin = all(torch.isfinite(t).all().item() for t in tensors(out) if t.is_floating_point())
            mx = max([t.detach().float().abs().max().item() for t in tensors(out) if t.is_floating_point() and t.numel()] or [0.0])
            stats.append((name, type(mod).__name__, mx))
            if not fin and not bad:
                imx = [t.detach().float().abs().max().item() for t in tensors(inp) if t.is_floating_point() and t.numel()]
                bad.append((name, type(mod).__name__, imx))
        return fn

    for n, m in model.named_modules():
        if n:
            m.register_forward_hook(hook(n))
    batch = sam3_step.move_to_device(sam3_step.collate(sam3_step.synthetic_datapoints(a.batch, seed=0)), "cuda")
    outs = None
    try:
        outs = model(batch)
    except Exception as e:  # noqa: BLE001
        print("forward raised:", type(e).__name__, str(e)[:300])
    print("first non-finite:", bad[:1])
    big = sorted(stats, key=lambda s: -s[2] if s[2] == s[2] else 0)[:15]
    print("largest |output| by module:")
    for s in big:
        print("   ", s)
    if bad:
        idx = [i for i, s in enumerate(stats) if s[0] == bad[0][0]][0]
        print("modules just before the first non-finite one:")
        for s in stats[max(0, idx - 12):idx + 1]:
            print("   ", s)


    if outs is None or bad:
        return
    # backward of ONE loss term with backward hooks on every module: the first module (in backward order) that turns a
    # finite grad_output into a non-finite grad_input is the culprit
    matcher, wrapper = sam3_step.build_objective(native=True)
    events = []

    def bhook(name):
        def fn(mod, gin, gout):
            go = [t for t in tensors(gout) if t is not None]
            gi = [t for t in tensors(gin) if t is not None]
            f_o = all(torch.isfinite(t).all().item() for t in go)
            f_i = all(torch.isfinite(t).all().item() for t in gi)
            mo = max([t.float().abs().max().item() for t in go if t.numel()] or [0.0])
            mi = max([t.float().abs().max().item() for t in gi if t.numel()] or [0.0])
            events.append((name, type(mod).__name__, f_o, f_i, mo, mi))
        return fn

    for n, m in model.named_modules():
        if n and not any(True for _ in m.children()):
            m.register_full_backward_hook(bhook(n))
    from sam3_lora_b200 import mha as MH, ops as OPS

    orig_b = MH._AttnCoreFn.backward
    seen = []

    def dbg_backward(ctx, gO):
        out = orig_b(ctx, gO)
        fin = [bool(torch.isfinite(t).all().item()) for t in out[:3]]
        meta = ctx.meta[:9]
        if not all(fin) or not torch.isfinite(gO).all():
            seen.append(("ATTN", meta, "gO finite", bool(torch.isfinite(gO).all().item()), "dq/dk/dv finite", fin,
                         "gO amax", float(gO.abs().max()), "nan counts", [int(torch.isnan(t).sum()) for t in out[:3]],
                         "numel", [t.numel() for t in out[:3]]))
        return out

    MH._AttnCoreFn.backward = staticmethod(dbg_backward)
    orig_l = OPS._LoRALinearFn.backward

    def dbg_lin(ctx, gy):
        out = orig_l(ctx, gy)
        bad_o = [i for i, t in enumerate(out) if isinstance(t, torch.Tensor) and not torch.isfinite(t).all()]
        if bad_o or not torch.isfinite(gy).all():
            seen.append(("LINEAR", ctx.meta[:4], "gy finite", bool(torch.isfinite(gy).all().item()), "bad outputs", bad_o, "gy amax", float(gy.abs().max())))
        return out

    OPS._LoRALinearFn.backward = staticmethod(dbg_lin)
    loss, loss_dict = sam3_step.training_loss(model, batch, matcher, wrapper)
    term = loss_dict[a.term] if a.term != "core" else loss
    print("term", a.term, float(term))
    params = [p for p in model.parameters() if p.requires_grad]
    g = torch.autograd.grad(term, params, allow_unused=True)
    print("non-finite adapter grads:", sum(1 for x in g if x is not None and not torch.isfinite(x).all()), "of", len(g))
    print("custom-function events (backward order), first 6:")
    for e in seen[:6]:
        print("   ", e)
    firsts = [e for e in events if e[2] and not e[3]]
    print("modules with finite grad_output but non-finite grad_input (backward order):")
    for e in firsts[:8]:
        print("   ", e)
    print("first 25 backward events:")
    for e in events[:25]:
        print("   ", e)
    nf = [i for i, e in enumerate(events) if not e[2] or not e[3]]
    if nf:
        print("events around the first non-finite one:")
        for e in events[max(0, nf[0] - 6):nf[0] + 3]:
            print("   ", e)


if __name__ == "__main__":
    main()
