"""Time every BASELINE.json config on one B200 (CUDA events, warm) and print a JSON table:
algorithmic flops / bytes, achieved rate and the fraction of the measured roofline.
Config 5 is measured as its per-GPU shard (8192 rows of the 65536 batch, KL over 1/8 of
the weight rows): the full config needs 8 GPUs (bench.py --gpus 8 covers the scaling)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cplxmodule_b200 as cb                                               # noqa: E402
from cplxmodule_b200 import cplx                                           # noqa: E402
from cplxmodule_b200.nn import CplxConv2d, CplxLinear                      # noqa: E402
from cplxmodule_b200.nn.relevance import (CplxConv2dVD, CplxLinearARD, CplxLinearVD, LinearVD,
                                          penalties)                      # noqa: E402
from cplxmodule_b200 import ops, _native as nv                            # noqa: E402

DEV = "cuda"
PEAKS = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0, "bf16_tflops_sustained": 1400.0}


def timeit(fn, iters, warm=3):
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


def row(name, ms, flops, nbytes, note=""):
    tf = flops / ms / 1e9
    gbs = nbytes / ms / 1e6
    r = dict(config=name, ms=round(ms, 4), algorithmic_gflop=round(flops / 1e9, 2),
             algorithmic_mb=round(nbytes / 1e6, 1), tflops=round(tf, 1), gbs=round(gbs, 1),
             frac_tensor=round(tf / PEAKS["bf16_tflops_sustained"], 4),
             frac_tensor_burst=round(tf / PEAKS.get("bf16_tflops", PEAKS["bf16_tflops_sustained"]), 4),
             frac_hbm=round(gbs / PEAKS["hbm_gbs"], 4), note=note)
    print(json.dumps(r), flush=True)
    return r


def main():
    torch.manual_seed(0)
    cb.set_kl_fusion(False)      # forward and KL are timed as separate calls here
    out = []
    with torch.no_grad():
        # 1. real LinearVD 784 -> 256, batch 128, forward + penalties (launch-latency bound)
        m = LinearVD(784, 256).to(DEV).train()
        x = torch.randn(128, 784, device=DEV)
        ms = timeit(lambda: (m(x), sum(penalties(m))), 200, 20)
        out.append(row("1 LinearVD 784->256 B=128 fwd+KL", ms, 2 * 2 * 128 * 256 * 784,
                       4 * (128 * 784 + 3 * 256 * 784 + 128 * 256), "2 launches + prepass; latency bound"))
        # 2. CplxLinear 4096^2, bf16, B=4096
        lin = CplxLinear(4096, 4096).to(DEV).bfloat16()
        z = cplx.randn(4096, 4096, device=DEV).to(torch.bfloat16)
        ms = timeit(lambda: lin(z), 20)
        out.append(row("2 CplxLinear 4096->4096 bf16 B=4096", ms, 8 * 4096 ** 3,
                       2 * 6 * 4096 ** 2, "tcgen05 kind::f16, 4 MMAs/k-step"))
        del lin, z
        # 3. headline, fp32 and bf16
        for dt, es in ((torch.float32, 4), (torch.bfloat16, 2)):
            vd = CplxLinearVD(4096, 4096).to(DEV).train().to(dt)
            z = cplx.randn(4096, 4096, device=DEV).to(dt)
            ms_f = timeit(lambda: vd(z), 20)
            ms_k = timeit(lambda: sum(penalties(vd)), 50)
            out.append(row(f"3 CplxLinearVD 4096->4096 B=4096 fwd ({'fp32 planes / scaled fp16 operands' if es == 4 else 'bf16'})",
                           ms_f, 10 * 4096 ** 3, es * 7 * 4096 ** 2, "pre-pass + fused GEMM"))
            out.append(row(f"3 CplxLinearVD KL ({'fp32' if es == 4 else 'bf16'})", ms_k,
                           0, es * 3 * 4096 ** 2, "kl_kernel, HBM bound"))
            del vd, z
        # 4. CplxConv2d 64->64 3x3 on 256x64x128x128 (padding 0), and its VD variant
        for cls, name, nconv in ((CplxConv2d, "CplxConv2d", 4), (CplxConv2dVD, "CplxConv2dVD", 5)):
            conv = cls(64, 64, 3).to(DEV).train()
            z = cplx.randn(256, 64, 128, 128, device=DEV)
            flops = nconv * 2 * 256 * 64 * 126 * 126 * 64 * 9
            for dt, es, tag in ((torch.float32, 4, "fp32"), (torch.bfloat16, 2, "bf16")):
                convd, zd = conv.to(dt), z.to(dt)
                ms = timeit(lambda: convd(zd), 5, 2)
                nbytes = es * (2 * 256 * 64 * 128 * 128 + 2 * 256 * 64 * 126 * 126)
                out.append(row(f"4 {name} 64->64 3x3 128x128 B=256 {tag}", ms, flops, nbytes,
                               "NCHW: transposing pre-pass (fp32 planes: per-image scaled fp16 operands) + CTA-pair tcgen05 implicit GEMM"))
                zcl = cplx.Cplx(zd.real.contiguous(memory_format=torch.channels_last),
                                zd.imag.contiguous(memory_format=torch.channels_last))
                ms = timeit(lambda: convd(zcl), 5, 2)
                out.append(row(f"4 {name} 64->64 3x3 128x128 B=256 {tag} channels_last in/out", ms,
                               flops, nbytes, "CTA-pair implicit-GEMM kernel reads NHWC planes in place"))
                if cls is CplxConv2dVD:
                    cb.set_noise_mode("fast")
                    ms = timeit(lambda: convd(zcl), 5, 2)
                    cb.set_noise_mode("torch")
                    out.append(row(f"4 {name} 64->64 3x3 128x128 B=256 {tag} channels_last, fast noise",
                                   ms, flops, nbytes, "one Philox + Box-Muller per complex element"))
                del zcl
            ops.set_math_mode("simt")
            conv32, z32 = conv.float(), z
            ms = timeit(lambda: conv32(z32), 2, 1)
            ops.set_math_mode("auto")
            out.append(row(f"4 {name} fp32 exact (conv_simt_kernel)", ms, flops,
                           4 * (2 * 256 * 64 * 128 * 128 + 2 * 256 * 64 * 126 * 126), "CUDA-core path"))
            del conv, z, convd, zd
        # 5. CplxLinearARD 8192^2, per-GPU shard of the 8-GPU config: 8192 rows, 1/8 of the KL
        ard = CplxLinearARD(8192, 8192).to(DEV).train()
        z = cplx.randn(8192, 8192, device=DEV)
        ms_f = timeit(lambda: ard(z), 5, 2)
        w = ard.weight
        ms_k = timeit(lambda: ops.kl(nv.KL_CPLX_ARD, w.real[:1024], w.imag[:1024],
                                     ard.log_sigma2[:1024], "sum"), 50)
        out.append(row("5 CplxLinearARD 8192->8192, per-GPU shard B=8192 fwd fp32 planes / scaled fp16 operands", ms_f,
                       10 * 8192 ** 3, 4 * 7 * 8192 ** 2, "1/8 of the global batch 65536"))
        out.append(row("5 CplxLinearARD KL row shard (1024 of 8192 rows)", ms_k, 0,
                       4 * 3 * 1024 * 8192, "all-reduce of the scalar not included"))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "configs_r2.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
