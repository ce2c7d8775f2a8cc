"""Per-tile timeline of the headline GEMM kernel (cluster 0, leader CTA) from an instrumented side
build:  CPLXK_BUILD_TAG=trace CPLXK_BUILD_FLAGS=-DCPLXK_TRACE python -m cplxmodule_b200.build
        CPLXK_LIB=cplxmodule_b200/csrc/libcplxk_trace.so python tools/tc3_trace.py
Prints microseconds (SM clock / measured MHz) relative to the kernel's first stamp."""
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cplxmodule_b200 as cb                                          # noqa: E402
from cplxmodule_b200 import _native as nv                             # noqa: E402
from cplxmodule_b200 import cplx                                      # noqa: E402
from cplxmodule_b200.nn.relevance import CplxLinearVD, penalties      # noqa: E402

cb.set_noise_mode(os.environ.get("NOISE", "torch"))
torch.manual_seed(0)
layer = CplxLinearVD(4096, 4096).cuda().train()
x = cplx.randn(4096, 4096, device="cuda")
trace = torch.zeros(16 * 8, dtype=torch.int64, device="cuda")
lib = nv.lib()
lib.cplxk_debug_set_trace.argtypes = [ctypes.c_void_p]
with torch.no_grad():
    for _ in range(20):
        layer(x)
    torch.cuda.synchronize()
    lib.cplxk_debug_set_trace(ctypes.c_void_p(trace.data_ptr()))
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        layer(x)
    b.record()
    torch.cuda.synchronize()
    lib.cplxk_debug_set_trace(None)
t = trace.cpu().view(16, 8)
rows = [r for r in t.tolist() if r[0] or r[1]]
t0 = min(v for r in rows for v in r if v)
mhz = float(os.environ.get("SM_MHZ", "1500"))
names = ["mma_start", "mma_done", "noise_done", "accum_seen", "tmem_back", "stores_done", "loads_issued"]
for i, r in enumerate(rows):
    print(json.dumps({"tile": i, **{n: round((v - t0) / mhz, 2) if v else None for n, v in zip(names, r)}}))
print(json.dumps({"fwd_ms": a.elapsed_time(b) / 10, "assumed_sm_mhz": mhz}))
