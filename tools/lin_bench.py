"""Plain CplxLinear 4096^3 forward: fp32 planes (scaled-fp16 persistent kernel vs tf32) and bf16
planes (persistent double-buffered kernel vs one-tile-per-CTA kernel), interleaved on one box."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cplxmodule_b200 import ops                   # noqa: E402

DEV = "cuda"
M = N = K = int(os.environ.get("KB_SIZE", "4096"))


def timeit(fn, iters=50, warm=3):
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
    with torch.no_grad():
        for dt_name, dt, variants in (("f32", torch.float32, {"f16-persistent": "auto", "tf32": "tf32"}),
                                      ("bf16", torch.bfloat16, {"persistent": "auto"})):
            xr = (torch.randn(M, K, device=DEV) / 2 ** 0.5).to(dt)
            xi = (torch.randn(M, K, device=DEV) / 2 ** 0.5).to(dt)
            bound = 1 / (2 * K) ** 0.5
            w_re = torch.empty(N, K, device=DEV).uniform_(-bound, bound).to(dt)
            w_im = torch.empty(N, K, device=DEV).uniform_(-bound, bound).to(dt)
            b_re, b_im = torch.randn(N, device=DEV).to(dt), torch.randn(N, device=DEV).to(dt)
            fn = lambda: ops.cplx_linear(xr, xi, w_re, w_im, b_re, b_im)
            res = {v: [] for v in variants}
            for _ in range(4):
                for v, mode in variants.items():
                    ops.set_math_mode(mode)
                    res[v].append(timeit(fn))
                    ops.set_math_mode("auto")
            for v, ts in res.items():
                ms = sorted(ts)[len(ts) // 2]
                print(json.dumps(dict(dtype=dt_name, variant=v, ms_med=round(ms, 4), ms_min=round(min(ts), 4),
                                      tflops=round(8 * M * N * K / ms / 1e9, 1))), flush=True)


if __name__ == "__main__":
    main()
