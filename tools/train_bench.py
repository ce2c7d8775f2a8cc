"""Time one TRAINING step of the headline layer on one GPU: forward (train mode, fused noise)
+ KL + backward of  loss = <y, c> + C * KL  + Adam-free parameter touch (no optimizer)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cplxmodule_b200 import cplx                                      # noqa: E402
from cplxmodule_b200.nn.relevance import CplxLinearVD, penalties      # noqa: E402

B = D = 4096
for dt_name, dt in (("f32", torch.float32), ("bf16", torch.bfloat16)):
    torch.manual_seed(0)
    layer = CplxLinearVD(D, D).cuda().train().to(dt)
    x = cplx.randn(B, D, device="cuda").to(dt)
    c_re = torch.randn(B, D, device="cuda", dtype=dt)
    c_im = torch.randn(B, D, device="cuda", dtype=dt)

    def step():
        layer.zero_grad(set_to_none=True)
        y = layer(x)
        loss = (y.real * c_re).sum() + (y.imag * c_im).sum() + 1e-3 * sum(penalties(layer))
        loss.backward()

    klw = torch.tensor(1e-3, device="cuda", dtype=torch.float32)

    def step_injected():
        # the same gradients without the synthetic loss's own kernels (two products, two
        # reductions and their backward = ~170 us of torch elementwise work per step): the
        # upstream gradients c_re, c_im and the KL weight are handed to autograd directly
        layer.zero_grad(set_to_none=True)
        y = layer(x)
        kl = sum(penalties(layer))
        torch.autograd.backward([y.real, y.imag, kl], [c_re, c_im, klw.to(kl.dtype)])

    for fn, tag in ((step, "loss = <y, c> + 1e-3 KL"), (step_injected, "upstream gradients injected")):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        n = 10
        for _ in range(n):
            fn()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / n
        flops = 30.0 * B * D * D   # fwd 10 MNK + bwd 20 MNK
        print(json.dumps(dict(case=f"CplxLinearVD 4096 train step {dt_name}, {tag}", ms=round(ms, 3),
                              samples_per_s=round(B / ms * 1e3), tflops=round(flops / ms / 1e9, 1))))
