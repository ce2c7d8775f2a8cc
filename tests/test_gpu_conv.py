"""Complex conv1d/conv2d (+ variational forward) on the GPU vs the oracle / golden fixtures."""
import pytest
import torch
import torch.nn.functional as F

import cplxmodule_b200 as cb
from cplxmodule_b200 import cplx
from cplxmodule_b200.nn import CplxConv1d, CplxConv2d
from cplxmodule_b200.nn.relevance import CplxConv2dVD, penalties
from oracle import cplx_oracle as orc
from tests.conftest import load_golden, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _mod(cls, g, suffix="", **kw):
    w = g[f"w{suffix}_re"]
    m = cls(w.shape[1], w.shape[0], tuple(w.shape[2:]), **kw)
    sd = {"weight.real": w, "weight.imag": g[f"w{suffix}_im"], "bias.real": g[f"b{suffix}_re"],
          "bias.imag": g[f"b{suffix}_im"]}
    if "log_sigma2" in g and suffix == "" and hasattr(m, "log_sigma2"):
        sd["log_sigma2"] = g["log_sigma2"]
    m.load_state_dict(sd)
    return m.to(DEV)


def test_golden_conv2d():
    g = load_golden("cplx_conv2d")
    z = cplx.Cplx(g["x_re"].to(DEV), g["x_im"].to(DEV))
    out = _mod(CplxConv2d, g)(z)
    assert out.shape == g["y_re"].shape
    assert rel_err(out.real, g["y_re"]) < 2e-5 and rel_err(out.imag, g["y_im"]) < 2e-5
    out = _mod(CplxConv2d, g, "2", stride=(2, 1), padding=(1, 2), dilation=(1, 2))(z)
    assert out.shape == g["y2_re"].shape
    assert rel_err(out.real, g["y2_re"]) < 2e-5 and rel_err(out.imag, g["y2_im"]) < 2e-5


def test_golden_conv2d_vd():
    g = load_golden("cplx_conv2d_vd")
    layer = _mod(CplxConv2dVD, g, padding=1).train()
    z = cplx.Cplx(g["x_re"].to(DEV), g["x_im"].to(DEV))
    with torch.no_grad():
        out = layer(z, eps=cplx.Cplx(g["eps_re"].to(DEV), g["eps_im"].to(DEV)))
        kl = sum(penalties(layer))
    assert rel_err(out.real, g["y_re"]) < 2e-5 and rel_err(out.imag, g["y_im"]) < 2e-5
    assert abs(kl.item() - g["penalty_sum"].item()) / g["penalty_sum"].item() < 1e-3


def test_conv2d_vd_fused_noise_matches_device_draw():
    torch.manual_seed(5)
    layer = CplxConv2dVD(3, 4, 3, padding=1).to(DEV).train()
    with torch.no_grad():
        layer.log_sigma2.uniform_(-6, 1)
    z = cplx.randn(2, 3, 9, 8, device=DEV)
    torch.manual_seed(42)
    with torch.no_grad():
        fused = layer(z)
    torch.manual_seed(42)
    eps = cplx.randn(2, 4, 9, 8, device=DEV)
    with torch.no_grad():
        inject = layer(z, eps=eps)
    assert torch.equal(fused.real, inject.real) and torch.equal(fused.imag, inject.imag)


def test_conv1d_groups_circular_vs_torch_formula():
    torch.manual_seed(6)
    # conv1d, stride / dilation / padding as in the reference's tests (tests/test_cplx.py:272-330)
    m = CplxConv1d(4, 6, 5, stride=2, padding=3, dilation=2).to(DEV)
    z = cplx.randn(3, 4, 40, device=DEV)
    out = m(z)
    c = lambda t: t.detach().cpu()
    ww = torch.cat([c(m.weight.real), c(m.weight.imag)], 0)
    wr = F.conv1d(c(z.real), ww, None, 2, 3, 2)
    wi = F.conv1d(c(z.imag), ww, None, 2, 3, 2)
    re = wr[:, :6] - wi[:, 6:] + c(m.bias.real)[None, :, None]
    im = wr[:, 6:] + wi[:, :6] + c(m.bias.imag)[None, :, None]
    assert out.shape == re.shape
    assert rel_err(out.real, re) < 2e-5 and rel_err(out.imag, im) < 2e-5
    # groups = 2 (convnd_naive, cplx.py:717-726)
    m = CplxConv2d(4, 6, 3, groups=2, bias=False).to(DEV)
    z = cplx.randn(2, 4, 7, 7, device=DEV)
    out = m(z)
    f = lambda a, w: F.conv2d(c(a), c(w), None, 1, 0, 1, 2)
    re = f(z.real, m.weight.real) - f(z.imag, m.weight.imag)
    im = f(z.real, m.weight.imag) + f(z.imag, m.weight.real)
    assert rel_err(out.real, re) < 2e-5 and rel_err(out.imag, im) < 2e-5
    # circular padding (cplx.py:701-714, 784-786)
    m = CplxConv2d(2, 3, 3, padding=2, padding_mode="circular").to(DEV)
    z = cplx.randn(1, 2, 6, 5, device=DEV)
    out = m(z)
    pad = lambda a: F.pad(c(a), (1, 1, 1, 1), mode="circular")
    want = orc.cplx_conv2d(pad(z.real), pad(z.imag), c(m.weight.real), c(m.weight.imag),
                           c(m.bias.real), c(m.bias.imag))
    assert rel_err(out.real, want[0]) < 2e-5 and rel_err(out.imag, want[1]) < 2e-5
    with pytest.raises(ValueError):
        cplx.conv2d(z, m.weight, None, padding_mode="reflect")
