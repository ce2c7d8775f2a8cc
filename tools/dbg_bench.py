"""Bottleneck triage of the CTA-pair fused forward (CPLXK_DBG modes of fwd_tc2.cu): full kernel,
MMAs without operand loads, operand loads without MMAs, prologue + noise + epilogue only."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cplxmodule_b200 as cb                      # noqa: E402
from cplxmodule_b200 import ops                   # noqa: E402

DEV = "cuda"
M = N = K = int(os.environ.get("KB_SIZE", "4096"))


def timeit(fn, iters=20, warm=3):
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
    dts = [d for d in (("f32", torch.float32), ("bf16", torch.bfloat16))
           if d[0] in os.environ.get("DBG_DTYPES", "f32,bf16").split(",")]
    for dt_name, dt in dts:
        xr = (torch.randn(M, K, device=DEV) / 2 ** 0.5).to(dt)
        xi = (torch.randn(M, K, device=DEV) / 2 ** 0.5).to(dt)
        bound = 1 / (2 * K) ** 0.5
        w_re = torch.empty(N, K, device=DEV).uniform_(-bound, bound).to(dt)
        w_im = torch.empty(N, K, device=DEV).uniform_(-bound, bound).to(dt)
        b_re = torch.randn(N, device=DEV).to(dt)
        b_im = torch.randn(N, device=DEV).to(dt)
        ls2 = torch.full((N, K), -10.0, device=DEV).to(dt)
        eps = (torch.randn(M, N, device=DEV).to(dt), torch.randn(M, N, device=DEV).to(dt))
        variants = os.environ.get("DBG_VARIANTS", "default,nopersist,tf32").split(",")
        dbgs = os.environ.get("DBG_MODES", "0").split(",")
        rounds = int(os.environ.get("DBG_ROUNDS", "4"))

        def setenv(variant):
            os.environ.pop("CPLXK_PERSIST", None)
            os.environ.pop("CPLXK_F16", None)
            os.environ.pop("CPLXK_RASTER", None)
            if variant.startswith("raster"):
                os.environ["CPLXK_RASTER"] = variant[6:]
            if variant == "nopersist":
                os.environ["CPLXK_PERSIST"] = "0"
            elif variant == "tf32":
                os.environ["CPLXK_F16"] = "0"

        for noise in os.environ.get("DBG_NOISE", "inject,torch,fast").split(","):
            for dbg in dbgs:
                os.environ["CPLXK_DBG"] = dbg
                klr = (lambda: {"kind": 2}) if os.environ.get("DBG_KL") else (lambda: None)
                if noise == "inject":
                    fn = lambda: ops.cplx_linear_vd(xr, xi, w_re, w_im, b_re, b_im, ls2, eps=eps, kl_req=klr())
                else:
                    cb.set_noise_mode(noise)
                    fn = lambda: ops.cplx_linear_vd(xr, xi, w_re, w_im, b_re, b_im, ls2, kl_req=klr())
                res = {v: [] for v in variants if not (v == "tf32" and dt_name != "f32")}
                for _ in range(rounds):          # interleave the variants: clocks drift under the power cap
                    for v in res:
                        setenv(v)
                        res[v].append(timeit(fn, iters=50))
                for v, ts in res.items():
                    rows.append(dict(dtype=dt_name, variant=v, noise=noise, dbg=dbg,
                                     ms_min=round(min(ts), 4), ms_med=round(sorted(ts)[len(ts) // 2], 4),
                                     ms_mean=round(sum(ts) / len(ts), 4)))
                    print(json.dumps(rows[-1]), flush=True)
        os.environ.pop("CPLXK_PERSIST", None)
        os.environ.pop("CPLXK_F16", None)
        cb.set_noise_mode("torch")
    os.environ.pop("CPLXK_DBG", None)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "dbg_bench.json"), "w") as f:
        json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
