"""Kernel-level timing sweep of the fused forward on one GPU (CUDA events, L2-exceeding
operands): swizzle/stage count, noise mode, dtype, plain vs variational, pre-pass vs in-kernel transform.  Writes gpurun_out/kbench.json."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cplxmodule_b200 as cb                      # noqa: E402
from cplxmodule_b200 import cplx, ops             # noqa: E402

DEV = "cuda"
M = N = K = int(os.environ.get("KB_SIZE", "4096"))


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
    rows = []
    for dt_name, dt in (("f32", torch.float32), ("bf16", torch.bfloat16)):
        x = cplx.Cplx((torch.randn(M, K, device=DEV) / 2 ** 0.5).to(dt),
                      (torch.randn(M, K, device=DEV) / 2 ** 0.5).to(dt))
        bound = 1 / (2 * K) ** 0.5
        w_re = torch.empty(N, K, device=DEV).uniform_(-bound, bound).to(dt)
        w_im = torch.empty(N, K, device=DEV).uniform_(-bound, bound).to(dt)
        b_re = torch.randn(N, device=DEV).to(dt)
        b_im = torch.randn(N, device=DEV).to(dt)
        ls2 = torch.full((N, K), -10.0, device=DEV).to(dt)
        eps = (torch.randn(M, N, device=DEV).to(dt), torch.randn(M, N, device=DEV).to(dt))
        for swz in ("64", "128"):
            os.environ["CPLXK_TC_SWIZZLE"] = swz
            cases = {
                "cplx_lin": lambda: ops.cplx_linear(x.real, x.imag, w_re, w_im, b_re, b_im),
                "cplx_vd_inject": lambda: ops.cplx_linear_vd(x.real, x.imag, w_re, w_im, b_re, b_im,
                                                             ls2, eps=eps),
                "cplx_vd_torch": lambda: ops.cplx_linear_vd(x.real, x.imag, w_re, w_im, b_re, b_im, ls2),
                "real_lin": lambda: ops.real_linear(x.real, w_re, b_re),
                "real_vd_torch": lambda: ops.real_linear_vd(x.real, w_re, b_re, ls2),
            }
            if os.environ.get("KB_QUICK"):
                cases = {k: v for k, v in cases.items() if k in ("cplx_lin", "cplx_vd_inject", "cplx_vd_torch")}
            for name, fn in list(cases.items()):
                if "vd" in name:
                    cb.set_operand_prepass(False)
                    ms = timeit(fn)
                    cb.set_operand_prepass(True)
                    rows.append(dict(dtype=dt_name, swz=swz, case=name + "_xform", ms=ms,
                                     tflops=(10 if "cplx" in name else 4) * M * N * K / ms / 1e9))
                    print(json.dumps(rows[-1]), flush=True)
            for name, fn in cases.items():
                cb.set_noise_mode("torch")
                ms = timeit(fn)
                nmma = {"cplx_lin": 8, "real_lin": 2, "real_vd_torch": 4}.get(name, 10)
                rows.append(dict(dtype=dt_name, swz=swz, case=name, ms=ms,
                                 tflops=nmma * M * N * K / ms / 1e9))
                print(json.dumps(rows[-1]), flush=True)
            cb.set_noise_mode("fast")
            ms = timeit(cases["cplx_vd_torch"])
            cb.set_noise_mode("torch")
            rows.append(dict(dtype=dt_name, swz=swz, case="cplx_vd_fast", ms=ms,
                             tflops=10 * M * N * K / ms / 1e9))
            print(json.dumps(rows[-1]), flush=True)
        os.environ.pop("CPLXK_TC_SWIZZLE", None)

    if os.environ.get("KB_QUICK"):
        json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "kbench_quick.json"), "w"), indent=1)
        return
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "kbench.json"), "w") as f:
        json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
