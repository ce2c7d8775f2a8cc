"""BASELINE config 4 (CplxConv2d 64->64 3x3 on 256x64x128x128) for ncu captures.

usage: prof_conv.py [steps] [f32|bf16] [nchw|nhwc] [plain|vd]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cplxmodule_b200 import cplx                                      # noqa: E402
from cplxmodule_b200.nn import CplxConv2d                             # noqa: E402
from cplxmodule_b200.nn.relevance import CplxConv2dVD                 # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dtype = torch.bfloat16 if (len(sys.argv) > 2 and sys.argv[2] == "bf16") else torch.float32
nhwc = len(sys.argv) > 3 and sys.argv[3] == "nhwc"
vd = len(sys.argv) > 4 and sys.argv[4] == "vd"
torch.manual_seed(0)
conv = (CplxConv2dVD if vd else CplxConv2d)(64, 64, 3).cuda().train().to(dtype)
z = cplx.randn(256, 64, 128, 128, device="cuda").to(dtype)
if nhwc:
    z = cplx.Cplx(z.real.contiguous(memory_format=torch.channels_last),
                  z.imag.contiguous(memory_format=torch.channels_last))
with torch.no_grad():
    for _ in range(steps):
        y = conv(z)
torch.cuda.synchronize()
print("ok", tuple(y.shape))
