"""Generate tests/golden/*.npz from the LIVE reference (run in the build container, where
/root/reference exists):   python oracle/make_golden.py

The reference tree is read-only and lacks the generated ``cplxmodule/__version__.py``
(cplxmodule/__init__.py:2 imports it; setup.py:7-12 writes it), so a stub module is
registered before importing -- no reference file is copied or modified.
"""
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("CPLX_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def import_reference():
    if "cplxmodule" in sys.modules:
        return sys.modules["cplxmodule"]
    if not os.path.isdir(os.path.join(REF, "cplxmodule")):
        raise ImportError(f"reference tree not found at {REF}")
    stub = types.ModuleType("cplxmodule.__version__")
    stub.__version__ = "2022.06"
    sys.modules["cplxmodule.__version__"] = stub
    sys.path.insert(0, REF)
    try:
        import cplxmodule  # noqa: F401
    finally:
        sys.path.remove(REF)
    return sys.modules["cplxmodule"]


def npy(d):
    return {k: (v.detach().numpy() if isinstance(v, torch.Tensor) else np.asarray(v))
            for k, v in d.items()}


def main():
    import_reference()
    from cplxmodule import cplx
    from cplxmodule.nn import CplxLinear, CplxConv2d
    from cplxmodule.nn.relevance import (LinearVD, LinearARD, CplxLinearVD, CplxLinearARD,
                                         CplxConv2dVD, penalties)
    os.makedirs(OUT, exist_ok=True)

    # -- CplxLinear 48 -> 40, ragged batch 37 (K, N not multiples of the tile)
    torch.manual_seed(101)
    lin = CplxLinear(48, 40)
    z = cplx.randn(37, 48)
    out = lin(z)
    np.savez(os.path.join(OUT, "cplx_linear.npz"), **npy(dict(
        x_re=z.real, x_im=z.imag, w_re=lin.weight.real, w_im=lin.weight.imag,
        b_re=lin.bias.real, b_im=lin.bias.imag, y_re=out.real, y_im=out.imag)))

    # -- CplxLinearVD 64 -> 72 training forward with captured noise, log_sigma2 ~ U(-12, 2)
    torch.manual_seed(202)
    vd = CplxLinearVD(64, 72).train()
    with torch.no_grad():
        vd.log_sigma2.uniform_(-12, 2)
    z = cplx.randn(50, 64)
    state = torch.get_rng_state()
    out = vd(z)
    torch.set_rng_state(state)
    eps = cplx.randn(50, 72)          # the very draw forward() consumed
    vd.eval()
    mu = vd(z)
    np.savez(os.path.join(OUT, "cplx_linear_vd.npz"), **npy(dict(
        x_re=z.real, x_im=z.imag, w_re=vd.weight.real, w_im=vd.weight.imag, b_re=vd.bias.real,
        b_im=vd.bias.imag, log_sigma2=vd.log_sigma2, eps_re=eps.real, eps_im=eps.imag,
        y_re=out.real, y_im=out.imag, mu_re=mu.real, mu_im=mu.imag,
        log_alpha=vd.log_alpha, penalty=vd.penalty,
        penalty_sum=sum(penalties(vd, reduction="sum")),
        penalty_mean=sum(penalties(vd, reduction="mean")))))

    # -- CplxLinearARD penalty on the same kind of parameters
    torch.manual_seed(303)
    ard = CplxLinearARD(33, 21)
    with torch.no_grad():
        ard.log_sigma2.uniform_(-12, 2)
    np.savez(os.path.join(OUT, "cplx_linear_ard.npz"), **npy(dict(
        w_re=ard.weight.real, w_im=ard.weight.imag, log_sigma2=ard.log_sigma2,
        log_alpha=ard.log_alpha, penalty=ard.penalty,
        penalty_sum=sum(penalties(ard, reduction="sum")),
        relevance=ard.relevance(threshold=3.0))))

    # -- real LinearVD / LinearARD (config 1 shape class, reduced): 96 -> 56, batch 24
    for name, cls, seed in (("linear_vd", LinearVD, 404), ("linear_ard", LinearARD, 505)):
        torch.manual_seed(seed)
        m = cls(96, 56).train()
        with torch.no_grad():
            m.log_sigma2.uniform_(-12, 2)
        x = torch.randn(24, 96)
        state = torch.get_rng_state()
        out = m(x)
        torch.set_rng_state(state)
        eps = torch.randn(24, 56)
        np.savez(os.path.join(OUT, f"{name}.npz"), **npy(dict(
            x=x, w=m.weight, b=m.bias, log_sigma2=m.log_sigma2, eps=eps, y=out,
            log_alpha=m.log_alpha, penalty=m.penalty,
            penalty_sum=sum(penalties(m, reduction="sum")))))

    # -- penalties on a log_alpha sweep [-40, 40] (forced through the parameters:
    #    w = 1 + 0j  =>  log_alpha == log_sigma2 up to the 1e-12 guard)
    la = torch.linspace(-40, 40, 161)
    one, zero = torch.ones_like(la), torch.zeros_like(la)
    sweep = {"log_sigma2": la}
    for name, cls in (("real_vd", LinearVD), ("real_ard", LinearARD)):
        m = cls(161, 1)
        with torch.no_grad():
            m.weight.copy_(one[None]); m.log_sigma2.copy_(la[None])
        sweep[name] = m.penalty[0]
    for name, cls in (("cplx_vd", CplxLinearVD), ("cplx_ard", CplxLinearARD)):
        m = cls(161, 1)
        with torch.no_grad():
            m.weight.real.copy_(one[None]); m.weight.imag.copy_(zero[None])
            m.log_sigma2.copy_(la[None])
        sweep[name] = m.penalty[0]
    np.savez(os.path.join(OUT, "penalty_sweep.npz"), **npy(sweep))

    # -- CplxConv2d 6 -> 5 ch, 3x3, on 2 x 6 x 11 x 13, default stride/padding + a strided case
    torch.manual_seed(606)
    conv = CplxConv2d(6, 5, 3)
    z = cplx.randn(2, 6, 11, 13)
    out = conv(z)
    conv2 = CplxConv2d(6, 5, (3, 2), stride=(2, 1), padding=(1, 2), dilation=(1, 2))
    out2 = conv2(z)
    np.savez(os.path.join(OUT, "cplx_conv2d.npz"), **npy(dict(
        x_re=z.real, x_im=z.imag, w_re=conv.weight.real, w_im=conv.weight.imag,
        b_re=conv.bias.real, b_im=conv.bias.imag, y_re=out.real, y_im=out.imag,
        w2_re=conv2.weight.real, w2_im=conv2.weight.imag, b2_re=conv2.bias.real,
        b2_im=conv2.bias.imag, y2_re=out2.real, y2_im=out2.imag)))

    # -- CplxConv2dVD training forward with captured noise
    torch.manual_seed(707)
    cvd = CplxConv2dVD(4, 6, 3, padding=1).train()
    with torch.no_grad():
        cvd.log_sigma2.uniform_(-12, 2)
    z = cplx.randn(2, 4, 9, 10)
    state = torch.get_rng_state()
    out = cvd(z)
    torch.set_rng_state(state)
    eps = cplx.randn(*out.shape)
    np.savez(os.path.join(OUT, "cplx_conv2d_vd.npz"), **npy(dict(
        x_re=z.real, x_im=z.imag, w_re=cvd.weight.real, w_im=cvd.weight.imag,
        b_re=cvd.bias.real, b_im=cvd.bias.imag, log_sigma2=cvd.log_sigma2, eps_re=eps.real,
        eps_im=eps.imag, y_re=out.real, y_im=out.imag,
        penalty_sum=sum(penalties(cvd, reduction="sum")))))
    real_conv_fixtures()
    extension_fixtures()
    bilinear_fixtures()
    print("golden fixtures written to", os.path.normpath(OUT))


def real_conv_fixtures():
    """Real-valued Conv2dVD / Conv1dARD training forwards with captured noise
    (nn/relevance/real/base.py:83-177, real/vd.py:103-125, real/ard.py:48-57)."""
    import_reference()
    from cplxmodule.nn.relevance import Conv1dARD, Conv2dVD, penalties
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(808)
    m = Conv2dVD(6, 8, (3, 2), stride=(1, 2), padding=(1, 0), dilation=(2, 1), groups=2).train()
    with torch.no_grad():
        m.log_sigma2.uniform_(-12, 2)
    x = torch.randn(3, 6, 10, 9)
    state = torch.get_rng_state()
    out = m(x)
    torch.set_rng_state(state)
    eps = torch.randn_like(out)
    m.eval()
    mu = m(x)
    np.savez(os.path.join(OUT, "conv2d_vd.npz"), **npy(dict(
        x=x, w=m.weight, b=m.bias, log_sigma2=m.log_sigma2, eps=eps, y=out, mu=mu,
        log_alpha=m.log_alpha, penalty=m.penalty,
        penalty_sum=sum(penalties(m, reduction="sum")))))
    torch.manual_seed(909)
    m = Conv1dARD(5, 7, 4, stride=2, padding=3).train()
    with torch.no_grad():
        m.log_sigma2.uniform_(-12, 2)
    x = torch.randn(4, 5, 23)
    state = torch.get_rng_state()
    out = m(x)
    torch.set_rng_state(state)
    eps = torch.randn_like(out)
    np.savez(os.path.join(OUT, "conv1d_ard.npz"), **npy(dict(
        x=x, w=m.weight, b=m.bias, log_sigma2=m.log_sigma2, eps=eps, y=out,
        log_alpha=m.log_alpha, penalty=m.penalty,
        penalty_sum=sum(penalties(m, reduction="sum")),
        relevance=m.relevance(threshold=3.0))))


def extension_fixtures():
    """CplxLinearVDApprox / CplxLinearVDScaleFree penalties (nn/relevance/extensions/complex.py)."""
    import_reference()
    from cplxmodule.nn.relevance.extensions import CplxLinearVDApprox, CplxLinearVDScaleFree
    from cplxmodule.nn.relevance import penalties
    os.makedirs(OUT, exist_ok=True)
    out = {}
    for name, cls, seed in (("approx", CplxLinearVDApprox, 1010), ("scalefree", CplxLinearVDScaleFree, 1111)):
        torch.manual_seed(seed)
        m = cls(29, 17)
        with torch.no_grad():
            m.log_sigma2.uniform_(-12, 2)
        out.update({f"{name}_w_re": m.weight.real, f"{name}_w_im": m.weight.imag,
                    f"{name}_log_sigma2": m.log_sigma2, f"{name}_penalty": m.penalty,
                    f"{name}_penalty_sum": sum(penalties(m, reduction="sum"))})
    np.savez(os.path.join(OUT, "ext_penalties.npz"), **npy(out))


def bilinear_fixtures():
    """CplxBilinearVD (conjugate and not) / BilinearARD training forwards with captured noise
    (cplx.py:1062-1090, nn/relevance/complex/base.py:59-84, real/base.py:52-80)."""
    import_reference()
    from cplxmodule import cplx
    from cplxmodule.nn.relevance import BilinearARD, CplxBilinearVD, penalties
    os.makedirs(OUT, exist_ok=True)
    out = {}
    for tag, conj, seed in (("conj", True, 1212), ("plain", False, 1313)):
        torch.manual_seed(seed)
        m = CplxBilinearVD(7, 10, 9, conjugate=conj).train()
        with torch.no_grad():
            m.log_sigma2.uniform_(-12, 2)
        z1, z2 = cplx.randn(11, 7), cplx.randn(11, 10)
        state = torch.get_rng_state()
        y = m(z1, z2)
        torch.set_rng_state(state)
        eps = cplx.randn(11, 9)
        m.eval()
        mu = m(z1, z2)
        out.update({f"{tag}_x1_re": z1.real, f"{tag}_x1_im": z1.imag, f"{tag}_x2_re": z2.real,
                    f"{tag}_x2_im": z2.imag, f"{tag}_w_re": m.weight.real, f"{tag}_w_im": m.weight.imag,
                    f"{tag}_b_re": m.bias.real, f"{tag}_b_im": m.bias.imag,
                    f"{tag}_log_sigma2": m.log_sigma2, f"{tag}_eps_re": eps.real,
                    f"{tag}_eps_im": eps.imag, f"{tag}_y_re": y.real, f"{tag}_y_im": y.imag,
                    f"{tag}_mu_re": mu.real, f"{tag}_mu_im": mu.imag,
                    f"{tag}_penalty_sum": sum(penalties(m, reduction="sum"))})
    torch.manual_seed(1414)
    m = BilinearARD(6, 5, 8).train()
    with torch.no_grad():
        m.log_sigma2.uniform_(-12, 2)
    x1, x2 = torch.randn(13, 6), torch.randn(13, 5)
    state = torch.get_rng_state()
    y = m(x1, x2)
    torch.set_rng_state(state)
    eps = torch.randn_like(y)
    m.eval()
    out.update({"real_x1": x1, "real_x2": x2, "real_w": m.weight, "real_b": m.bias,
                "real_log_sigma2": m.log_sigma2, "real_eps": eps, "real_y": y, "real_mu": m(x1, x2),
                "real_penalty_sum": sum(penalties(m, reduction="sum"))})
    np.savez(os.path.join(OUT, "bilinear.npz"), **npy(out))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "bilinear":
        bilinear_fixtures()
    elif len(sys.argv) > 1 and sys.argv[1] == "real_conv":
        real_conv_fixtures()
    elif len(sys.argv) > 1 and sys.argv[1] == "extensions":
        extension_fixtures()
    else:
        main()
