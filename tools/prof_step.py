"""A few headline steps (CplxLinearVD 4096->4096, B=4096, forward + KL) for ncu captures."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cplxmodule_b200 as cb                                          # noqa: E402
from cplxmodule_b200 import cplx                                      # noqa: E402
from cplxmodule_b200.nn.relevance import CplxLinearVD, penalties      # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 6
dtype = sys.argv[2] if len(sys.argv) > 2 else "f32"
cb.set_noise_mode(os.environ.get("NOISE", "torch"))
torch.manual_seed(0)
layer = CplxLinearVD(4096, 4096).cuda().train()
x = cplx.randn(4096, 4096, device="cuda")
if dtype == "bf16":
    layer = layer.bfloat16()
    x = x.to(torch.bfloat16)
with torch.no_grad():
    for _ in range(steps):
        y = layer(x)
        kl = sum(penalties(layer))
torch.cuda.synchronize()
print("ok", float(kl))
