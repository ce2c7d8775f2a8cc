"""Config 4 (CplxConv2d / CplxConv2dVD 64->64 3x3 on 256x64x128x128) in every plane layout, one
JSON line per case; run it with CPLXK_LIB=<other build> for a same-box A/B."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cplxmodule_b200 import cplx                               # noqa: E402
from cplxmodule_b200.nn import CplxConv2d                      # noqa: E402
from cplxmodule_b200.nn.relevance import CplxConv2dVD          # noqa: E402

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
    lib = os.environ.get("CPLXK_LIB", "default")
    with torch.no_grad():
        for cls in ((CplxConv2d,) if "--plain" in sys.argv else (CplxConv2d, CplxConv2dVD)):
            conv = cls(64, 64, 3).to(DEV).train()
            z = cplx.randn(256, 64, 128, 128, device=DEV)
            for dt, tag in ((torch.float32, "fp32"), (torch.bfloat16, "bf16")):
                if "--fp32-nchw" in sys.argv and tag != "fp32":
                    continue
                if "--bf16-nchw" in sys.argv and tag != "bf16":
                    continue
                convd, zd = conv.to(dt), z.to(dt)
                zcl = cplx.Cplx(zd.real.contiguous(memory_format=torch.channels_last),
                                zd.imag.contiguous(memory_format=torch.channels_last))
                for name, inp in (("nchw", zd), ("nhwc", zcl)):
                    if ("--fp32-nchw" in sys.argv or "--bf16-nchw" in sys.argv) and name != "nchw":
                        continue
                    ms = timeit(lambda: convd(inp))
                    print(json.dumps(dict(lib=os.path.basename(lib), amax_pass=os.environ.get("CPLXK_CONV_AMAX_PASS", "0"), row=os.environ.get("CPLXK_CONV_ROW", "1"), overlap=os.environ.get("CPLXK_CONV_OVERLAP", "8"),
                                          layer=cls.__name__, dtype=tag,
                                          layout=name, ms=round(ms, 4))), flush=True)
                del zcl, zd
            del conv, z


if __name__ == "__main__":
    main()
