import os, sys
sys.path.insert(0, "/root/repo")
import torch
from noisynet_b200 import ops
dev = torch.device("cuda:0")
s_a = 1 / 3.0
x = torch.randint(0, 16, (512, 65, 14, 14), device=dev).float() * s_a
gy = torch.randn(512, 120, 10, 10, device=dev)
for _ in range(3):
    ops.conv_wgrad(gy, x, (120, 65, 5, 5), precision="bf16", a_code_scale=s_a)
torch.cuda.synchronize()
