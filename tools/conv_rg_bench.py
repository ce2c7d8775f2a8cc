"""Row f3 timings on config-4-sized work: real-plane convolution and grouped complex convolution
through the single-launch tcgen05 kernel, beside the exact-fp32 CUDA-core kernel ('simt') and
torch's own F.conv2d (cuDNN, what the reference calls) on the same box.  One JSON line per case."""
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cplxmodule_b200 import conv_ops, cplx, ops               # noqa: E402
from cplxmodule_b200.nn import CplxConv2d                      # noqa: E402
from cplxmodule_b200.nn.relevance import Conv2dVD              # noqa: E402

DEV = "cuda"


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


def main():
    torch.manual_seed(0)
    out = lambda **kw: print(json.dumps(kw), flush=True)
    with torch.no_grad():
        B, C, H, W, O = 256, 64, 128, 128, 64
        flops = 2.0 * B * O * (H - 2) * (W - 2) * C * 9
        x = torch.randn(B, C, H, W, device=DEV)
        w = torch.randn(O, C, 3, 3, device=DEV) / 24
        for dt, tag in ((torch.float32, "fp32"), (torch.bfloat16, "bf16")):
            xd, wd = x.to(dt), w.to(dt)
            ms = timeit(lambda: conv_ops.real_convnd(2, xd, wd))
            out(case="real conv 64->64 3x3 256x64x128x128", impl="tcgen05", dtype=tag, ms=round(ms, 4),
                tflops=round(flops / ms / 1e9, 1))
            for allow in (False, True):
                torch.backends.cudnn.allow_tf32 = allow
                ms = timeit(lambda: F.conv2d(xd, wd))
                out(case="real conv 64->64 3x3 256x64x128x128", impl=f"torch F.conv2d (cudnn, allow_tf32={allow})",
                    dtype=tag, ms=round(ms, 4), tflops=round(flops / ms / 1e9, 1))
        ops.set_math_mode("simt")
        ms = timeit(lambda: conv_ops.real_convnd(2, x, w), iters=2, warm=1)
        ops.set_math_mode("auto")
        out(case="real conv 64->64 3x3 256x64x128x128", impl="simt (exact fp32)", dtype="fp32", ms=round(ms, 4),
            tflops=round(flops / ms / 1e9, 1))
        m = Conv2dVD(C, O, 3).to(DEV).train()
        ms = timeit(lambda: m(x))
        out(case="Conv2dVD 64->64 3x3 256x64x128x128 (mean + variance conv + torch-exact noise)", impl="tcgen05",
            dtype="fp32", ms=round(ms, 4), tflops=round(2 * flops / ms / 1e9, 1))
        del x, w, m
        for groups in (2, 4):
            conv = CplxConv2d(64, 64, 3, groups=groups).to(DEV)
            z = cplx.randn(B, C, H, W, device=DEV)
            gflops = 4 * flops / groups
            for dt, tag in ((torch.float32, "fp32"), (torch.bfloat16, "bf16")):
                cd, zd = conv.to(dt), z.to(dt)
                ms = timeit(lambda: cd(zd))
                out(case=f"CplxConv2d 64->64 3x3 groups={groups} 256x64x128x128", impl="tcgen05, one launch", dtype=tag,
                    ms=round(ms, 4), tflops=round(gflops / ms / 1e9, 1))
            del conv, z


if __name__ == "__main__":
    main()
