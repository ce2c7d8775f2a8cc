"""Headline step in the SUSTAINED regime: the part is kept under load for `SETTLE` seconds first
(the 1 kW power cap pulls the SM clock from 1.97 to ~1.5 GHz within a few hundred ms), then K
steps are timed.  Knobs through the environment (NOISE, CPLXK_RASTER, CPLXK_LIB, ...).
    python tools/sustained_bench.py [B D [layer]]   ->  one JSON line"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cplxmodule_b200 as cb                      # noqa: E402
from cplxmodule_b200 import cplx                  # noqa: E402
from cplxmodule_b200.nn import relevance          # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    D = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    name = sys.argv[3] if len(sys.argv) > 3 else "CplxLinearVD"
    settle = float(os.environ.get("SETTLE", "1.5"))
    cb.set_noise_mode(os.environ.get("NOISE", "torch"))
    torch.manual_seed(0)
    layer = getattr(relevance, name)(D, D).cuda().train()
    x = cplx.randn(B, D, device="cuda")

    def step():
        layer(x)
        return sum(relevance.penalties(layer))

    def timed(n):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            step()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / n

    with torch.no_grad():
        for _ in range(5):
            step()
        burst = timed(50)
        t0 = time.perf_counter()
        while time.perf_counter() - t0 < settle:
            for _ in range(50):
                step()
            torch.cuda.synchronize()
        sustained = timed(300)
    print(json.dumps({"B": B, "D": D, "layer": name, "noise": cb.get_noise_mode(),
                      "raster": os.environ.get("CPLXK_RASTER", "6"), "lib": os.environ.get("CPLXK_LIB", "default"),
                      "burst_ms": round(burst, 4), "sustained_ms": round(sustained, 4),
                      "sustained_step_tflops": round(10.0 * B * D * D / sustained / 1e9, 1)}), flush=True)


if __name__ == "__main__":
    main()
