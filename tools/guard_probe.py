"""Cost of the device-side guard of the fused KL hand-out: step time with the check on / off
(set_kl_fusion(True | "unchecked")), interleaved; run once more with CPLXK_PDL=0 to see the guard
serialised behind the GEMM instead of running under it."""
import sys, json, torch
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cplxmodule_b200 as cb
from cplxmodule_b200 import cplx
from cplxmodule_b200.nn import relevance
torch.manual_seed(0)
layer = relevance.CplxLinearVD(4096, 4096).cuda().train()
x = cplx.randn(4096, 4096, device="cuda")
def step():
    layer(x); return sum(relevance.penalties(layer))
def timed(n=60):
    for _ in range(5): step()
    torch.cuda.synchronize(); torch.cuda._sleep(100_000_000)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): step()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
with torch.no_grad():
    out = []
    for rep in range(3):
        for mode in (True, "unchecked"):
            cb.set_kl_fusion(mode)
            out.append((str(mode), round(timed(), 4)))
            import time; time.sleep(1.0)
print(json.dumps(out))
