"""Short workload for ncu captures: the first blocks of the SAM3 trunk (depth-1 window blocks + 1 global)
at batch 8, a few training steps through the native engine.  Never used for benchmark numbers."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from sam3_lora_b200.lora_layers import LoRAConfig, apply_lora_to_model, get_lora_parameters  # noqa: E402
from sam3_lora_b200.vit import ViT  # noqa: E402

depth = int(sys.argv[1]) if len(sys.argv) > 1 else 8
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
B = 8
torch.manual_seed(0)
model = ViT(depth=depth, global_att_blocks=(depth - 1,), max_batch=B)
apply_lora_to_model(model, LoRAConfig(rank=16, alpha=32, dropout=0.0,
                                      target_modules=["q_proj", "k_proj", "v_proj", "out_proj", "fc1", "fc2"]))
for p in get_lora_parameters(model):
    if p.shape[0] == 16:
        torch.nn.init.normal_(p, std=0.02)
model = model.cuda().train()
img = torch.randn(B, 3, 1008, 1008, device="cuda")
gout = torch.randn(B, 1024, 72, 72, device="cuda") * 1e-3
for _ in range(steps):
    out = model(img)[0]
    (out * gout).sum().backward()
torch.cuda.synchronize()
print("done")
