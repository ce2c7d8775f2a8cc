"""Real-plane conv 64->64 3x3 on 256x64x128x128 for ncu captures: prof_real_conv.py [steps] [f32|bf16]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cplxmodule_b200 import conv_ops                           # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dt = torch.bfloat16 if (len(sys.argv) > 2 and sys.argv[2] == "bf16") else torch.float32
torch.manual_seed(0)
x = torch.randn(256, 64, 128, 128, device="cuda").to(dt)
w = (torch.randn(64, 64, 3, 3, device="cuda") / 24).to(dt)
with torch.no_grad():
    for _ in range(steps):
        y = conv_ops.real_convnd(2, x, w)
torch.cuda.synchronize()
print("ok", tuple(y.shape))
