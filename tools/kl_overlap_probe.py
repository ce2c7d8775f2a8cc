"""Probe: does the stand-alone KL kernel co-run with the persistent GEMM when launched on a
low-priority side stream?  Prints device ms per step for: fused KL (pre-pass by-product), no KL,
no-KL forward + side-stream KL kernel."""
import json
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cplxmodule_b200 as cb                      # noqa: E402
from cplxmodule_b200 import cplx, ops             # noqa: E402
from cplxmodule_b200.nn import relevance          # noqa: E402


def timed(fn, n=30):
    fn(); fn(); fn()
    torch.cuda.synchronize()
    torch.cuda._sleep(100_000_000)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def main():
    B = D = 4096
    dev = torch.device("cuda")
    torch.manual_seed(0)
    layer = relevance.CplxLinearVD(D, D).to(dev).train()
    x = cplx.randn(B, D, device=dev)
    lo, hi = torch.cuda.Stream.priority_range()
    main_s = torch.cuda.Stream(dev, priority=hi)      # hi = numerically lowest = highest priority
    side = torch.cuda.Stream(dev, priority=lo)
    w = layer.weight
    out = {}
    with torch.no_grad(), torch.cuda.stream(main_s):
        def fused():
            layer(x); sum(relevance.penalties(layer))
        out["fused_kl_ms"] = timed(fused)
        cb.set_kl_fusion(False)
        out["no_kl_ms"] = timed(lambda: layer(x))

        def serial():
            layer(x); ops.kl(layer._kl_kind, w.real, w.imag, layer.log_sigma2, "sum")
        out["fwd_then_kl_same_stream_ms"] = timed(serial)

        def overlapped():
            ev = torch.cuda.Event(); ev.record(main_s)
            layer(x)
            side.wait_event(ev)
            with torch.cuda.stream(side):
                ops.kl(layer._kl_kind, w.real, w.imag, layer.log_sigma2, "sum")
                done = torch.cuda.Event(); done.record(side)
            main_s.wait_event(done)
        out["fwd_with_side_stream_kl_ms"] = timed(overlapped)
        cb.set_kl_fusion(True)
    print(json.dumps({k: round(v, 4) for k, v in out.items()}))


if __name__ == "__main__":
    main()
