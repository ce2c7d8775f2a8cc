"""Real-plane conv 64->64 3x3 on 256x64x128x128 (+ the real Conv2dVD forward that composes two of them)
for same-box A/B runs of the kernel choice (CPLXK_CONV_REAL_PAIR, CPLXK_CONV_ROW); one JSON line per case."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cplxmodule_b200 import conv_ops                           # noqa: E402
from cplxmodule_b200.nn.relevance import Conv2dVD              # noqa: E402


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


torch.manual_seed(0)
knobs = {k: os.environ.get(k, "1") for k in ("CPLXK_CONV_REAL_PAIR", "CPLXK_CONV_OVERLAP", "CPLXK_CONV_ROW")}
with torch.no_grad():
    x = torch.randn(256, 64, 128, 128, device="cuda")
    w = torch.randn(64, 64, 3, 3, device="cuda") / 24
    w128 = torch.randn(128, 64, 3, 3, device="cuda") / 24
    m = Conv2dVD(64, 64, 3).cuda().train()
    for dt, tag in ((torch.float32, "fp32"), (torch.bfloat16, "bf16")):
        xd, wd, w2, md = x.to(dt), w.to(dt), w128.to(dt), m.to(dt)
        print(json.dumps(dict(knobs, case="real conv 64->64", dtype=tag,
                              ms=round(timeit(lambda: conv_ops.real_convnd(2, xd, wd)), 4))), flush=True)
        print(json.dumps(dict(knobs, case="real conv 64->128", dtype=tag,
                              ms=round(timeit(lambda: conv_ops.real_convnd(2, xd, w2)), 4))), flush=True)
        print(json.dumps(dict(knobs, case="Conv2dVD 64->64 (torch-exact noise)", dtype=tag,
                              ms=round(timeit(lambda: md(xd)), 4))), flush=True)
