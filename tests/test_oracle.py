"""The oracle must reproduce the reference: (a) the committed golden fixtures generated
from the live reference, (b) the live reference itself when /root/reference is present,
(c) the float64 closed forms the reference's own tests use."""
import os

import numpy as np
import pytest
import scipy.signal
import scipy.special
import torch

from oracle import cplx_oracle as orc
from tests.conftest import load_golden

HAVE_REF = os.path.isdir("/root/reference/cplxmodule")


def test_golden_cplx_linear():
    g = load_golden("cplx_linear")
    re, im = orc.cplx_linear(g["x_re"], g["x_im"], g["w_re"], g["w_im"], g["b_re"], g["b_im"])
    assert torch.equal(re, g["y_re"]) and torch.equal(im, g["y_im"])


def test_golden_cplx_linear_vd():
    g = load_golden("cplx_linear_vd")
    re, im = orc.cplx_linear_vd(g["x_re"], g["x_im"], g["w_re"], g["w_im"], g["b_re"], g["b_im"],
                                g["log_sigma2"], g["eps_re"], g["eps_im"])
    assert torch.equal(re, g["y_re"]) and torch.equal(im, g["y_im"])
    mu_re, mu_im = orc.cplx_linear(g["x_re"], g["x_im"], g["w_re"], g["w_im"], g["b_re"], g["b_im"])
    assert torch.equal(mu_re, g["mu_re"]) and torch.equal(mu_im, g["mu_im"])
    la = orc.log_alpha_cplx(g["w_re"], g["w_im"], g["log_sigma2"])
    assert torch.equal(la, g["log_alpha"])
    assert torch.equal(orc.penalty_cplx_vd(la), g["penalty"])
    assert torch.equal(orc.layer_penalty("cplx_vd", g["w_re"], g["w_im"], g["log_sigma2"], "sum"),
                       g["penalty_sum"])
    assert torch.equal(orc.layer_penalty("cplx_vd", g["w_re"], g["w_im"], g["log_sigma2"], "mean"),
                       g["penalty_mean"])


def test_golden_cplx_ard():
    g = load_golden("cplx_linear_ard")
    la = orc.log_alpha_cplx(g["w_re"], g["w_im"], g["log_sigma2"])
    assert torch.equal(la, g["log_alpha"])
    assert torch.equal(orc.penalty_cplx_ard(la), g["penalty"])
    assert torch.equal((la <= 3.0).to(la), g["relevance"])


@pytest.mark.parametrize("name,kind", [("linear_vd", "real_vd"), ("linear_ard", "real_ard")])
def test_golden_real(name, kind):
    g = load_golden(name)
    y = orc.real_linear_vd(g["x"], g["w"], g["b"], g["log_sigma2"], g["eps"])
    assert torch.equal(y, g["y"])
    la = orc.log_alpha_real(g["w"], g["log_sigma2"])
    assert torch.equal(la, g["log_alpha"])
    assert torch.equal(orc.PENALTY[kind](la), g["penalty"])
    assert torch.equal(orc.layer_penalty(kind, g["w"], None, g["log_sigma2"], "sum"),
                       g["penalty_sum"])


def test_golden_real_conv_vd():
    """real-valued Conv2dVD (grouped, strided, dilated) and Conv1dARD fixtures of the live reference"""
    g = load_golden("conv2d_vd")
    geom = ((1, 2), (1, 0), (2, 1), 2)
    assert torch.equal(orc.real_conv2d_vd(g["x"], g["w"], g["b"], g["log_sigma2"], g["eps"], *geom), g["y"])
    assert torch.equal(orc.real_conv2d_vd(g["x"], g["w"], g["b"], g["log_sigma2"], None, *geom), g["mu"])
    la = orc.log_alpha_real(g["w"], g["log_sigma2"])
    assert torch.equal(la, g["log_alpha"]) and torch.equal(orc.penalty_real_vd(la), g["penalty"])
    assert torch.equal(orc.layer_penalty("real_vd", g["w"], None, g["log_sigma2"], "sum"), g["penalty_sum"])
    g = load_golden("conv1d_ard")
    assert torch.equal(orc.real_conv1d_vd(g["x"], g["w"], g["b"], g["log_sigma2"], g["eps"], 2, 3, 1, 1), g["y"])
    la = orc.log_alpha_real(g["w"], g["log_sigma2"])
    assert torch.equal(orc.penalty_real_ard(la), g["penalty"])
    assert torch.equal((la <= 3.0).to(la), g["relevance"])


def test_golden_extension_penalties():
    """nn/relevance/extensions/complex.py: VDApprox and VDScaleFree penalties of the live reference"""
    g = load_golden("ext_penalties")
    for name, kind in (("approx", "cplx_vd_approx"), ("scalefree", "cplx_vd_scalefree")):
        w_re, w_im, ls2 = g[f"{name}_w_re"], g[f"{name}_w_im"], g[f"{name}_log_sigma2"]
        assert torch.equal(orc.layer_penalty(kind, w_re, w_im, ls2, None), g[f"{name}_penalty"])
        assert torch.equal(orc.layer_penalty(kind, w_re, w_im, ls2, "sum"), g[f"{name}_penalty_sum"])


def test_golden_penalty_sweep():
    g = load_golden("penalty_sweep")
    la = g["log_sigma2"]
    one, zero = torch.ones_like(la), torch.zeros_like(la)
    for kind in ("real_vd", "real_ard"):
        assert torch.equal(orc.PENALTY[kind](orc.log_alpha_real(one, la)), g[kind])
    for kind in ("cplx_vd", "cplx_ard"):
        assert torch.equal(orc.PENALTY[kind](orc.log_alpha_cplx(one, zero, la)), g[kind])


def test_golden_conv2d():
    g = load_golden("cplx_conv2d")
    re, im = orc.cplx_conv2d(g["x_re"], g["x_im"], g["w_re"], g["w_im"], g["b_re"], g["b_im"])
    assert torch.equal(re, g["y_re"]) and torch.equal(im, g["y_im"])
    re, im = orc.cplx_conv2d(g["x_re"], g["x_im"], g["w2_re"], g["w2_im"], g["b2_re"], g["b2_im"],
                             stride=(2, 1), padding=(1, 2), dilation=(1, 2))
    assert torch.equal(re, g["y2_re"]) and torch.equal(im, g["y2_im"])
    g = load_golden("cplx_conv2d_vd")
    re, im = orc.cplx_conv2d_vd(g["x_re"], g["x_im"], g["w_re"], g["w_im"], g["b_re"], g["b_im"],
                                g["log_sigma2"], g["eps_re"], g["eps_im"], padding=1)
    assert torch.equal(re, g["y_re"]) and torch.equal(im, g["y_im"])


# ------------------------------------------------- closed forms used by the reference's tests
def test_linear_matches_numpy_dot():
    """tests/test_cplx.py:251-269: linear == np.dot(a, L.T) + b in float64."""
    rng = np.random.RandomState(7)
    a = rng.randn(5, 200) + 1j * rng.randn(5, 200)
    L = rng.randn(321, 200) + 1j * rng.randn(321, 200)
    b = rng.randn(321) + 1j * rng.randn(321)
    t = lambda v: torch.from_numpy(np.ascontiguousarray(v))
    re, im = orc.cplx_linear(t(a.real), t(a.imag), t(L.real), t(L.imag), t(b.real), t(b.imag))
    ref = np.dot(a, L.T) + b
    assert np.allclose(re.numpy() + 1j * im.numpy(), ref)


def test_conv_matches_scipy_correlate():
    """tests/test_cplx.py:272-301: conv == scipy.signal.correlate(x, w.conj(), 'valid')."""
    rng = np.random.RandomState(8)
    x = rng.randn(1, 1, 12, 9) + 1j * rng.randn(1, 1, 12, 9)
    w = rng.randn(1, 1, 3, 4) + 1j * rng.randn(1, 1, 3, 4)
    t = lambda v: torch.from_numpy(np.ascontiguousarray(v))
    re, im = orc.cplx_conv2d(t(x.real), t(x.imag), t(w.real), t(w.imag))
    ref = scipy.signal.correlate(x[0, 0], w[0, 0].conj(), mode="valid")
    assert np.allclose(re[0, 0].numpy() + 1j * im[0, 0].numpy(), ref)


def test_expi_matches_scipy():
    """tests/test_relevance.py:40-49."""
    x = torch.randn(200, dtype=torch.double)
    assert np.allclose(orc.expi(x).numpy(), scipy.special.expi(x.numpy()))


def test_exact64_penalty_agrees_with_reference_formula_in_f64():
    la = torch.linspace(-30, 12, 400, dtype=torch.double)
    exact = orc.penalty_cplx_vd_exact64(la)
    ref64 = orc.penalty_cplx_vd(la)
    assert torch.allclose(exact, ref64, rtol=1e-9, atol=1e-12)
    # and stays accurate where the reference's formula cancels catastrophically
    la = torch.tensor([20.0, 30.0, 40.0], dtype=torch.double)
    t = torch.exp(-la)
    assert torch.allclose(orc.penalty_cplx_vd_exact64(la), t - t * t / 4, rtol=1e-12)


# -------------------------------------------------------------- live reference (build box)
@pytest.mark.skipif(not HAVE_REF, reason="/root/reference not present on this machine")
def test_oracle_vs_live_reference():
    from oracle.make_golden import import_reference
    import_reference()
    from cplxmodule import cplx
    from cplxmodule.nn.relevance import CplxLinearVD, CplxLinearARD, LinearVD, LinearARD, penalties

    torch.manual_seed(1234)
    for cls, kind in ((CplxLinearVD, "cplx_vd"), (CplxLinearARD, "cplx_ard")):
        m = cls(70, 45).train()
        with torch.no_grad():
            m.log_sigma2.uniform_(-12, 2)
        z = cplx.randn(19, 70)
        gen_state = torch.get_rng_state()
        out = m(z)
        torch.set_rng_state(gen_state)
        er, ei = orc.cplx_randn(19, 45)
        re, im = orc.cplx_linear_vd(z.real, z.imag, m.weight.real, m.weight.imag, m.bias.real,
                                    m.bias.imag, m.log_sigma2, er, ei)
        assert torch.equal(re, out.real) and torch.equal(im, out.imag)
        assert torch.equal(orc.layer_penalty(kind, m.weight.real, m.weight.imag, m.log_sigma2),
                           sum(penalties(m)))
    for cls, kind in ((LinearVD, "real_vd"), (LinearARD, "real_ard")):
        m = cls(70, 45).train()
        with torch.no_grad():
            m.log_sigma2.uniform_(-12, 2)
        x = torch.randn(19, 70)
        gen_state = torch.get_rng_state()
        out = m(x)
        torch.set_rng_state(gen_state)
        eps = torch.randn(19, 45)
        assert torch.equal(orc.real_linear_vd(x, m.weight, m.bias, m.log_sigma2, eps), out)
        assert torch.equal(orc.layer_penalty(kind, m.weight, None, m.log_sigma2), sum(penalties(m)))


@pytest.mark.skipif(not HAVE_REF, reason="/root/reference not present on this machine")
def test_oracle_vs_live_reference_conv_vd_and_extensions():
    """real-valued conv VD forward (+ its gradients through the differentiable Ei restatement)
    and the extension penalties, bit-for-bit against the live reference"""
    from oracle.make_golden import import_reference
    import_reference()
    from cplxmodule.nn.relevance import Conv1dVD, Conv2dARD, penalties
    from cplxmodule.nn.relevance.extensions import CplxLinearVDApprox, CplxLinearVDScaleFree

    torch.manual_seed(4321)
    m = Conv2dARD(4, 6, 3, stride=2, padding=1, dilation=1, groups=2).train()
    with torch.no_grad():
        m.log_sigma2.uniform_(-12, 2)
    x = torch.randn(3, 4, 11, 9)
    state = torch.get_rng_state()
    out = m(x)
    torch.set_rng_state(state)
    eps = torch.randn_like(out)
    assert torch.equal(orc.real_conv2d_vd(x, m.weight, m.bias, m.log_sigma2, eps, 2, 1, 1, 2), out)
    assert torch.equal(orc.layer_penalty("real_ard", m.weight, None, m.log_sigma2), sum(penalties(m)))
    m = Conv1dVD(3, 5, 4, padding=2).train()
    x = torch.randn(2, 3, 17)
    state = torch.get_rng_state()
    out = m(x)
    torch.set_rng_state(state)
    eps = torch.randn_like(out)
    assert torch.equal(orc.real_conv1d_vd(x, m.weight, m.bias, m.log_sigma2, eps, 1, 2, 1, 1), out)

    for cls, kind in ((CplxLinearVDApprox, "cplx_vd_approx"), (CplxLinearVDScaleFree, "cplx_vd_scalefree")):
        m = cls(23, 11)
        with torch.no_grad():
            m.log_sigma2.uniform_(-12, 2)
        ref = sum(penalties(m))
        w_re, w_im, ls2 = (t.detach().clone().requires_grad_() for t in
                           (m.weight.real, m.weight.imag, m.log_sigma2))
        mine = orc.layer_penalty(kind, w_re, w_im, ls2)
        assert torch.equal(mine, ref)
        ref.backward()
        mine.backward()
        assert torch.equal(w_re.grad, m.weight.real.grad) and torch.equal(ls2.grad, m.log_sigma2.grad)


@pytest.mark.skipif(not HAVE_REF, reason="/root/reference not present on this machine")
def test_oracle_vs_live_reference_grouped_complex_conv():
    """groups > 1: cplx.conv2d -> convnd_naive (cplx.py:717-726) and the grouped CplxConv2dVD
    training forward (complex/base.py:120-135), bit-for-bit against the live reference"""
    from oracle.make_golden import import_reference
    import_reference()
    from cplxmodule import cplx
    from cplxmodule.nn.relevance import CplxConv2dVD

    torch.manual_seed(99)
    m = CplxConv2dVD(6, 4, (3, 2), stride=(1, 2), padding=(1, 0), dilation=(2, 1), groups=2).train()
    with torch.no_grad():
        m.log_sigma2.uniform_(-12, 2)
    z = cplx.randn(3, 6, 11, 9)
    w, b = m.weight, m.bias
    mu = m.eval()(z)
    re, im = orc.cplx_conv2d_grouped(z.real, z.imag, w.real, w.imag, b.real, b.imag, (1, 2), (1, 0), (2, 1), 2)
    assert torch.equal(re, mu.real) and torch.equal(im, mu.imag)
    m.train()
    state = torch.get_rng_state()
    out = m(z)
    torch.set_rng_state(state)
    er, ei = orc.cplx_randn(*out.shape)
    re, im = orc.cplx_conv2d_vd(z.real, z.imag, w.real, w.imag, b.real, b.imag, m.log_sigma2, er, ei,
                                (1, 2), (1, 0), (2, 1), 2)
    assert torch.equal(re, out.real) and torch.equal(im, out.imag)


def _bil(g, tag):
    names = ("x1_re", "x1_im", "x2_re", "x2_im", "w_re", "w_im", "b_re", "b_im")
    return [g[f"{tag}_{n}"] for n in names]


@pytest.mark.parametrize("tag,conj", [("conj", True), ("plain", False)])
def test_golden_cplx_bilinear(tag, conj):
    g = load_golden("bilinear")
    args = _bil(g, tag)
    mu = orc.cplx_bilinear(*args, conjugate=conj)
    assert torch.equal(mu[0], g[f"{tag}_mu_re"]) and torch.equal(mu[1], g[f"{tag}_mu_im"])
    y = orc.cplx_bilinear_vd(*args, g[f"{tag}_log_sigma2"], g[f"{tag}_eps_re"], g[f"{tag}_eps_im"], conj)
    assert torch.equal(y[0], g[f"{tag}_y_re"]) and torch.equal(y[1], g[f"{tag}_y_im"])
    O = args[4].shape[0]
    kl = orc.layer_penalty("cplx_vd", args[4].reshape(O, -1), args[5].reshape(O, -1),
                           g[f"{tag}_log_sigma2"].reshape(O, -1))
    assert torch.allclose(kl, g[f"{tag}_penalty_sum"], rtol=1e-6)
    # the identity the CUDA path relies on: bilinear == linear on the outer-product features
    x1 = torch.complex(args[0], args[1]).to(torch.complex128)
    x2 = torch.complex(args[2], args[3]).to(torch.complex128)
    z = ((x1.conj() if conj else x1)[:, :, None] * x2[:, None, :]).reshape(x1.shape[0], -1)
    lin = orc.cplx_linear(z.real, z.imag, args[4].double().reshape(O, -1), args[5].double().reshape(O, -1),
                          args[6].double(), args[7].double())
    assert torch.allclose(lin[0], mu[0].double(), atol=1e-5) and torch.allclose(lin[1], mu[1].double(), atol=1e-5)


def test_golden_real_bilinear():
    g = load_golden("bilinear")
    y = orc.real_bilinear_vd(g["real_x1"], g["real_x2"], g["real_w"], g["real_b"], g["real_log_sigma2"],
                             g["real_eps"])
    assert torch.equal(y, g["real_y"])
    assert torch.equal(orc.real_bilinear_vd(g["real_x1"], g["real_x2"], g["real_w"], g["real_b"], None, None),
                       g["real_mu"])
