"""First-contact diagnostics for the tcgen05 kernel on a real B200: every kernel variant
(real/complex x plain/VD x fp32/bf16 x swizzle 64/128) in its OWN subprocess under a timeout,
so a hang or a sticky CUDA error in one variant cannot take the others (or the box) down.
Writes gpurun_out/tc_diag.json.   usage: python tests/tools/tc_diag.py [--one cfg-json]"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def run_one(cfg):
    import torch
    from cplxmodule_b200 import ops
    from oracle import cplx_oracle as orc
    os.environ["CPLXK_TC_SWIZZLE"] = str(cfg["swz"])
    torch.manual_seed(0)
    M, N, K = cfg["M"], cfg["N"], cfg["K"]
    dt = torch.float32 if cfg["dtype"] == "f32" else torch.bfloat16
    dev = "cuda"
    mk = lambda *s: torch.randn(*s).to(dt)
    x_re, x_im = mk(M, K), mk(M, K)
    w_re, w_im = (torch.randn(N, K) / K ** 0.5).to(dt), (torch.randn(N, K) / K ** 0.5).to(dt)
    b_re, b_im = mk(N), mk(N)
    ls2 = torch.empty(N, K).uniform_(-8, 1).to(dt)
    e_re, e_im = mk(M, N), mk(M, N)
    d = lambda t: t.to(dev)
    c = lambda t: t.double()
    res = {}
    for mode in ("simt", "tensor"):
        ops.set_math_mode(mode)
        if cfg["cplx"] and cfg["vd"]:
            got = ops.cplx_linear_vd(d(x_re), d(x_im), d(w_re), d(w_im), d(b_re), d(b_im), d(ls2),
                                     eps=(d(e_re), d(e_im)))
            want = orc.cplx_linear_vd(c(x_re), c(x_im), c(w_re), c(w_im), c(b_re), c(b_im), c(ls2),
                                      c(e_re), c(e_im))
        elif cfg["cplx"]:
            got = ops.cplx_linear(d(x_re), d(x_im), d(w_re), d(w_im), d(b_re), d(b_im))
            want = orc.cplx_linear(c(x_re), c(x_im), c(w_re), c(w_im), c(b_re), c(b_im))
        elif cfg["vd"]:
            got = (ops.real_linear_vd(d(x_re), d(w_re), d(b_re), d(ls2), eps=d(e_re)),)
            want = (orc.real_linear_vd(c(x_re), c(w_re), c(b_re), c(ls2), c(e_re)),)
        else:
            got = (ops.real_linear(d(x_re), d(w_re), d(b_re)),)
            want = (torch.nn.functional.linear(c(x_re), c(w_re), c(b_re)),)
        torch.cuda.synchronize()
        errs = [float((g.double().cpu() - w).abs().max() / w.abs().max()) for g, w in zip(got, want)]
        res[mode] = errs
    print("RESULT " + json.dumps(res))


def main():
    if len(sys.argv) > 2 and sys.argv[1] == "--one":
        base = json.loads(sys.argv[2])
        for (M, N, K) in [(128, 128, 64), (256, 384, 512), (200, 136, 104)]:
            cfg = dict(base, M=M, N=N, K=K)
            print("SHAPE " + json.dumps([M, N, K]), flush=True)
            run_one(cfg)
            sys.stdout.flush()
        return
    out = []
    for dtype in ("f32", "bf16"):
        for swz in (128, 64):
            for cplx in (False, True):
                for vd in (False, True):
                    cfg = dict(dtype=dtype, swz=swz, cplx=cplx, vd=vd)
                    try:
                        p = subprocess.run([sys.executable, __file__, "--one", json.dumps(cfg)],
                                           capture_output=True, text=True, timeout=150)
                        stdout, cfg["rc"] = p.stdout, p.returncode
                        if p.returncode:
                            cfg["err"] = p.stderr[-800:]
                    except subprocess.TimeoutExpired as e:
                        stdout, cfg["rc"] = (e.stdout or b""), "timeout"
                        if isinstance(stdout, bytes):
                            stdout = stdout.decode(errors="replace")
                    cfg["lines"] = [l for l in stdout.splitlines() if l.startswith(("RESULT", "SHAPE"))]
                    out.append(cfg)
                    print(json.dumps(cfg), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "tc_diag.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
