import sys, torch
sys.path.insert(0, '.')
from sam3_lora_b200 import conv_ops as CO
x = (torch.randn(8, 288, 288, 256, device='cuda') * 0.5).half()
w9 = (torch.randn(256, 9 * 256, device='cuda') * 0.02).half()
for _ in range(5):
    CO.conv3x3(x, w9, None, out_f32=False)
torch.cuda.synchronize()
