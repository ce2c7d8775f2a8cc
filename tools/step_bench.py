"""Same-box A/B of the headline step: device time of the forward call (pre-pass + GEMM) and of
the pre-pass alone, several interleaved rounds.  Process-level knobs (CPLXK_PDL, CPLXK_RASTER,
CPLXK_LIB=<other build>) are read once at load, so an A/B is two runs of this script.
    python tools/step_bench.py [B D [layer]]   ->  one JSON line
"""
import json
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cplxmodule_b200 as cb                      # noqa: E402
from cplxmodule_b200 import _native as nv         # noqa: E402
from cplxmodule_b200 import cplx                  # noqa: E402
from cplxmodule_b200.nn import relevance          # noqa: E402


def device_time(fn, n=30):
    fn(); fn(); fn()
    torch.cuda.synchronize()
    torch.cuda._sleep(100_000_000)
    evs = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    return statistics.median(a.elapsed_time(b) for a, b in evs)


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    D = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    name = sys.argv[3] if len(sys.argv) > 3 else "CplxLinearVD"
    dev = torch.device("cuda")
    cb.set_noise_mode(os.environ.get("NOISE", "torch"))
    if os.environ.get("NOKL") == "1":
        cb.set_kl_fusion(False)
    torch.manual_seed(0)
    layer = getattr(relevance, name)(D, D).to(dev).train()
    x = cplx.randn(B, D, device=dev)
    lib = nv.lib()
    ws_bytes = lib.cplxk_linear_vd_workspace_bytes(B, D, D, 0)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    kl_sum = torch.empty((), dtype=torch.float32, device=dev)
    kl_ws = nv.kl_workspace(dev)
    w = layer.weight

    def prep():
        nv.check(lib.cplxk_linear_vd_prepare(nv.ptr(x.real), nv.ptr(x.imag), nv.ptr(w.real), nv.ptr(w.imag),
                                             nv.ptr(layer.log_sigma2), B, D, D, 0, nv.ptr(ws), ws_bytes,
                                             (-1 if os.environ.get("NOKL") == "1" else layer._kl_kind), nv.ptr(kl_sum), nv.ptr(kl_ws), kl_ws.numel() * 8,
                                             nv.stream_ptr(dev)))

    fwd, pre = [], []
    with torch.no_grad():
        for _ in range(int(os.environ.get("ROUNDS", "5"))):
            fwd.append(device_time(lambda: layer(x)))
            pre.append(device_time(prep))
    f, p = statistics.median(fwd), statistics.median(pre)
    print(json.dumps({"B": B, "D": D, "layer": name, "lib": os.environ.get("CPLXK_LIB", "default"),
                      "pdl": os.environ.get("CPLXK_PDL", "1"), "nokl": os.environ.get("NOKL", "0"), "raster": os.environ.get("CPLXK_RASTER", "6"),
                      "fwd_ms": round(f, 4), "fwd_min": round(min(fwd), 4), "prep_ms": round(p, 4),
                      "gemm_ms": round(f - p, 4), "step_tflops": round(10.0 * B * D * D / f / 1e9, 1)}))


if __name__ == "__main__":
    main()
