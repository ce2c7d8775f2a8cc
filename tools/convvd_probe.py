"""Variational conv forward on config-4-sized work: composed (mean conv + variance conv + in-place
noise launch) against the fused single kernel, injected noise and the stand-alone draw, complex and
real planes (ops.set_conv_vd_mode).  One JSON line per case."""
import sys, json, torch
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cplxmodule_b200 import cplx, ops, conv_ops
from cplxmodule_b200.nn.relevance import CplxConv2dVD
def timeit(fn, iters=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters
torch.manual_seed(0)
with torch.no_grad():
    conv = CplxConv2dVD(64, 64, 3).cuda().train()
    z = cplx.randn(256, 64, 128, 128, device="cuda")
    eps = cplx.randn(256, 64, 126, 126, device="cuda")
    for dt, tag in ((torch.float32, "fp32"), (torch.bfloat16, "bf16")):
        c, zz, ee = conv.to(dt), z.to(dt), eps.to(dt)
        import cplxmodule_b200 as cb
        cb.set_conv_vd_mode("composed")
        composed = round(timeit(lambda: c(zz)), 3)
        cb.set_conv_vd_mode("fused")
        print(json.dumps({"dtype": tag, "composed_torch_ms": composed, "fused_torch_ms": round(timeit(lambda: c(zz)), 3),
                          "inject_ms": round(timeit(lambda: c(zz, eps=ee)), 3),
                          "draw_ms": round(timeit(lambda: conv_ops._draw_noise(True, (256, 64, 126, 126), zz.real.device, dt)), 3)}))
    from cplxmodule_b200.nn.relevance import Conv2dVD
    rconv = Conv2dVD(64, 64, 3).cuda().train()
    x = torch.randn(256, 64, 128, 128, device="cuda")
    for dt, tag in ((torch.float32, "fp32"), (torch.bfloat16, "bf16")):
        c, xx = rconv.to(dt), x.to(dt)
        cb.set_conv_vd_mode("composed")
        composed = round(timeit(lambda: c(xx)), 3)
        cb.set_conv_vd_mode("fused")
        print(json.dumps({"layer": "Conv2dVD (real)", "dtype": tag, "composed_torch_ms": composed,
                          "fused_torch_ms": round(timeit(lambda: c(xx)), 3)}))
