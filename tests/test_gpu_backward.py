"""Gradients of the fused layers / KL on the GPU vs torch autograd over the float64 oracle."""
import pytest
import torch
import torch.nn.functional as F

import cplxmodule_b200 as cb
from cplxmodule_b200 import cplx, ops
from cplxmodule_b200.nn import CplxLinear
from cplxmodule_b200.nn.relevance import (CplxLinearARD, CplxLinearVD, LinearARD, LinearVD,
                                          penalties)
from oracle import cplx_oracle as orc
from tests.conftest import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = {"simt": 2e-4, "tensor": 3e-3}


@pytest.fixture(params=["simt", "tensor"])
def math(request):
    ops.set_math_mode(request.param)
    yield request.param
    ops.set_math_mode("auto")


def kl_with_grad64(kind, w_re, w_im, ls2):
    """sum(penalty) as a float64 torch graph; complex VD through its closed-form derivative
    d penalty / d log_alpha = exp(-1/alpha) - 1  (ExpiFunction.backward, complex/vd.py:38-41)."""
    la = orc.log_alpha_cplx(w_re, w_im, ls2) if kind.startswith("cplx") else orc.log_alpha_real(w_re, ls2)
    if kind != "cplx_vd":
        return orc.PENALTY[kind](la).sum()

    class _Ein(torch.autograd.Function):
        @staticmethod
        def forward(ctx, la_):
            ctx.save_for_backward(la_)
            return orc.penalty_cplx_vd_exact64(la_)

        @staticmethod
        def backward(ctx, g):
            (la_,) = ctx.saved_tensors
            return g * (torch.exp(-torch.exp(-la_)) - 1)

    return _Ein.apply(la).sum()


def make(M, N, K, seed):
    torch.manual_seed(seed)
    t = dict(x_re=torch.randn(M, K), x_im=torch.randn(M, K), w_re=torch.randn(N, K) / K ** 0.5,
             w_im=torch.randn(N, K) / K ** 0.5, b_re=torch.randn(N), b_im=torch.randn(N),
             ls2=torch.empty(N, K).uniform_(-6, 0), eps_re=torch.randn(M, N) / 2 ** 0.5,
             eps_im=torch.randn(M, N) / 2 ** 0.5, c_re=torch.randn(M, N), c_im=torch.randn(M, N))
    return t


@pytest.mark.parametrize("M,N,K", [(64, 48, 32), (200, 136, 264)])
@pytest.mark.parametrize("cls,kind", [(CplxLinearVD, "cplx_vd"), (CplxLinearARD, "cplx_ard")])
def test_cplx_vd_layer_gradients(M, N, K, cls, kind, math):
    t = make(M, N, K, M + N)
    klw = 0.37
    d = {k: v.double().requires_grad_(k not in ("eps_re", "eps_im", "c_re", "c_im")) for k, v in t.items()}
    y = orc.cplx_linear_vd(d["x_re"], d["x_im"], d["w_re"], d["w_im"], d["b_re"], d["b_im"], d["ls2"],
                           d["eps_re"], d["eps_im"])
    loss = (y[0] * d["c_re"]).sum() + (y[1] * d["c_im"]).sum() + klw * kl_with_grad64(
        kind, d["w_re"], d["w_im"], d["ls2"])
    loss.backward()

    layer = cls(K, N).to(DEV).train()
    with torch.no_grad():
        layer.weight.real.copy_(t["w_re"]); layer.weight.imag.copy_(t["w_im"])
        layer.bias.real.copy_(t["b_re"]); layer.bias.imag.copy_(t["b_im"])
        layer.log_sigma2.copy_(t["ls2"])
    x_re, x_im = t["x_re"].to(DEV).requires_grad_(), t["x_im"].to(DEV).requires_grad_()
    out = layer(cplx.Cplx(x_re, x_im), eps=cplx.Cplx(t["eps_re"].to(DEV), t["eps_im"].to(DEV)))
    loss_g = ((out.real * t["c_re"].to(DEV)).sum() + (out.imag * t["c_im"].to(DEV)).sum()
              + klw * sum(penalties(layer)))
    loss_g.backward()
    assert abs(loss_g.item() - loss.item()) / abs(loss.item()) < TOL[math]
    pairs = [(x_re.grad, d["x_re"].grad), (x_im.grad, d["x_im"].grad),
             (layer.weight.real.grad, d["w_re"].grad), (layer.weight.imag.grad, d["w_im"].grad),
             (layer.bias.real.grad, d["b_re"].grad), (layer.bias.imag.grad, d["b_im"].grad),
             (layer.log_sigma2.grad, d["ls2"].grad)]
    for got, want in pairs:
        assert got is not None and rel_err(got, want) < TOL[math]


@pytest.mark.parametrize("cls,kind", [(LinearVD, "real_vd"), (LinearARD, "real_ard")])
def test_real_vd_layer_gradients(cls, kind, math):
    M, N, K = 72, 40, 96
    t = make(M, N, K, 5)
    klw = 1.3
    d = {k: v.double().requires_grad_(k in ("x_re", "w_re", "b_re", "ls2")) for k, v in t.items()}
    eps = d["eps_re"] * 2 ** 0.5
    y = orc.real_linear_vd(d["x_re"], d["w_re"], d["b_re"], d["ls2"], eps)
    loss = (y * d["c_re"]).sum() + klw * kl_with_grad64(kind, d["w_re"], None, d["ls2"])
    loss.backward()
    layer = cls(K, N).to(DEV).train()
    with torch.no_grad():
        layer.weight.copy_(t["w_re"]); layer.bias.copy_(t["b_re"]); layer.log_sigma2.copy_(t["ls2"])
    x = t["x_re"].to(DEV).requires_grad_()
    out = layer(x, eps=(t["eps_re"] * 2 ** 0.5).to(DEV))
    loss_g = (out * t["c_re"].to(DEV)).sum() + klw * sum(penalties(layer))
    loss_g.backward()
    for got, want in [(x.grad, d["x_re"].grad), (layer.weight.grad, d["w_re"].grad),
                      (layer.bias.grad, d["b_re"].grad), (layer.log_sigma2.grad, d["ls2"].grad)]:
        assert rel_err(got, want) < TOL[math]


def test_plain_cplx_linear_gradients_with_batch_dims(math):
    torch.manual_seed(8)
    lin = CplxLinear(40, 24).to(DEV)
    x_re = torch.randn(3, 5, 40, device=DEV, requires_grad=True)
    x_im = torch.randn(3, 5, 40, device=DEV, requires_grad=True)
    c = torch.randn(3, 5, 24, device=DEV)
    out = lin(cplx.Cplx(x_re, x_im))
    ((out.real * c).sum() - (out.imag * c).sum()).backward()
    d = lambda v: v.detach().cpu().double().requires_grad_()
    xr, xi, wr, wi, br, bi = map(d, (x_re, x_im, lin.weight.real, lin.weight.imag, lin.bias.real,
                                     lin.bias.imag))
    o = orc.cplx_linear(xr, xi, wr, wi, br, bi)
    ((o[0] * c.cpu().double()).sum() - (o[1] * c.cpu().double()).sum()).backward()
    assert x_re.grad.shape == x_re.shape
    for got, want in [(x_re.grad, xr.grad), (x_im.grad, xi.grad), (lin.weight.real.grad, wr.grad),
                      (lin.weight.imag.grad, wi.grad), (lin.bias.real.grad, br.grad),
                      (lin.bias.imag.grad, bi.grad)]:
        assert rel_err(got, want) < TOL[math]


def test_fused_noise_backward_regenerates_the_forward_noise():
    """With in-kernel Philox noise the backward must see the SAME eps as the forward."""
    torch.manual_seed(3)
    M, N, K = 96, 80, 64
    layer = CplxLinearVD(K, N).to(DEV).train()
    with torch.no_grad():
        layer.log_sigma2.uniform_(-4, 0)
    x = cplx.randn(M, K, device=DEV)
    c = torch.randn(M, N, device=DEV)
    grads = []
    for inject in (False, True):
        layer.zero_grad()
        torch.manual_seed(99)
        if inject:
            eps = cplx.randn(M, N, device=DEV)
            out = layer(x, eps=eps)
        else:
            out = layer(x)
        ((out.real * c).sum() + (out.imag * c).sum()).backward()
        grads.append([layer.log_sigma2.grad.clone(), layer.weight.real.grad.clone()])
    assert torch.equal(grads[0][0], grads[1][0]) and torch.equal(grads[0][1], grads[1][1])


def test_penalty_elementwise_and_mean_gradients():
    torch.manual_seed(4)
    layer = CplxLinearVD(24, 16).to(DEV)
    with torch.no_grad():
        layer.log_sigma2.uniform_(-8, 2)
    coeff = torch.randn(16, 24, device=DEV)
    (layer.penalty * coeff).sum().backward()          # reduction=None path, per-element upstream grad
    w_re, w_im, ls2 = (v.detach().cpu().double().requires_grad_()
                       for v in (layer.weight.real, layer.weight.imag, layer.log_sigma2))
    la = orc.log_alpha_cplx(w_re, w_im, ls2)
    la.backward((torch.exp(-torch.exp(-la.detach())) - 1) * coeff.cpu().double())
    assert rel_err(layer.log_sigma2.grad, ls2.grad) < 2e-4
    assert rel_err(layer.weight.real.grad, w_re.grad) < 2e-4
    g_elem = layer.log_sigma2.grad.clone()
    layer.zero_grad()
    next(iter(penalties(layer, reduction="mean"))).backward()
    layer2_grad = layer.log_sigma2.grad.clone()
    layer.zero_grad()
    next(iter(penalties(layer, reduction="sum"))).backward()
    assert torch.allclose(layer2_grad * layer.log_sigma2.numel(), layer.log_sigma2.grad, rtol=1e-5)
    assert g_elem.shape == layer.log_sigma2.shape


def test_training_loop_drop_in():
    """The reference's training idiom (tests/test_relevance.py:62-73) runs end to end on the GPU
    and learns: loss = mse + C * sum(penalties(model))."""
    torch.manual_seed(0)
    model = torch.nn.Sequential(CplxLinearVD(32, 48), cb.nn.CplxToCplx[torch.nn.Tanh](),
                                CplxLinearVD(48, 8)).to(DEV)
    w_true = cplx.randn(8, 32, device=DEV)
    x = cplx.randn(512, 32, device=DEV)
    y = cplx.Cplx(x.real @ w_true.real.t() - x.imag @ w_true.imag.t(),
                  x.real @ w_true.imag.t() + x.imag @ w_true.real.t())
    opt = torch.optim.Adam(model.parameters(), lr=3e-3)
    losses = []
    for it in range(150):
        opt.zero_grad()
        out = model(x)
        mse = ((out.real - y.real) ** 2 + (out.imag - y.imag) ** 2).mean()
        loss = mse + 1e-4 * sum(penalties(model))
        loss.backward()
        opt.step()
        losses.append(mse.item())
    assert losses[-1] < 0.25 * losses[0]
    model.eval()
    with torch.no_grad():
        out = model(x)
    assert torch.isfinite(out.real).all()


# ------------------------------------------------------------------------------ convolutions
CONV_CASES = [
    # B, C, H, W, O, k, stride, padding, dilation
    (2, 4, 9, 10, 6, 3, 1, 1, 1),
    (3, 5, 12, 11, 4, (3, 2), (2, 1), (1, 0), (1, 2)),
    (2, 8, 16, 16, 8, 3, 2, 1, 1),
]


@pytest.mark.parametrize("B,C,H,W,O,k,stride,padding,dilation", CONV_CASES)
@pytest.mark.parametrize("vd", [False, True])
def test_cplx_conv2d_gradients(B, C, H, W, O, k, stride, padding, dilation, vd):
    """gradients of CplxConv2d / CplxConv2dVD (+ KL) w.r.t. input, kernel, bias and log_sigma2
    against float64 autograd over the oracle (cplx.py:729-742, complex/base.py:120-135)"""
    from cplxmodule_b200.nn import CplxConv2d
    from cplxmodule_b200.nn.relevance import CplxConv2dVD
    torch.manual_seed(B * 31 + C)
    cls = CplxConv2dVD if vd else CplxConv2d
    m = cls(C, O, k, stride=stride, padding=padding, dilation=dilation).to(DEV).train()
    if vd:
        with torch.no_grad():
            m.log_sigma2.uniform_(-6, 0)
    x_re = torch.randn(B, C, H, W, device=DEV, requires_grad=True)
    x_im = torch.randn(B, C, H, W, device=DEV, requires_grad=True)
    with torch.no_grad():
        shape = m.eval()(cplx.Cplx(x_re, x_im)).shape
        m.train()
    c_re, c_im = torch.randn(shape, device=DEV), torch.randn(shape, device=DEV)
    eps = cplx.Cplx(torch.randn(shape, device=DEV) / 2 ** 0.5, torch.randn(shape, device=DEV) / 2 ** 0.5)
    out = m(cplx.Cplx(x_re, x_im), eps=eps) if vd else m(cplx.Cplx(x_re, x_im))
    loss = (out.real * c_re).sum() + (out.imag * c_im).sum()
    if vd:
        loss = loss + 0.3 * sum(penalties(m))
    loss.backward()

    d = lambda t: t.detach().double().cpu().requires_grad_()
    xr, xi, wr, wi, br, bi = map(d, (x_re, x_im, m.weight.real, m.weight.imag, m.bias.real, m.bias.imag))
    geom = (m.stride, m.padding, m.dilation)
    if vd:
        l2 = d(m.log_sigma2)
        o = orc.cplx_conv2d_vd(xr, xi, wr, wi, br, bi, l2, eps.real.double().cpu(), eps.imag.double().cpu(), *geom)
        want = (o[0] * c_re.double().cpu()).sum() + (o[1] * c_im.double().cpu()).sum() \
            + 0.3 * kl_with_grad64("cplx_vd", wr, wi, l2)
    else:
        o = orc.cplx_conv2d(xr, xi, wr, wi, br, bi, *geom)
        want = (o[0] * c_re.double().cpu()).sum() + (o[1] * c_im.double().cpu()).sum()
    want.backward()
    assert abs(loss.item() - want.item()) <= 3e-3 * abs(want.item())
    pairs = [(x_re.grad, xr.grad), (x_im.grad, xi.grad), (m.weight.real.grad, wr.grad),
             (m.weight.imag.grad, wi.grad), (m.bias.real.grad, br.grad), (m.bias.imag.grad, bi.grad)]
    if vd:
        pairs.append((m.log_sigma2.grad, l2.grad))
    for got, ref in pairs:
        assert got is not None and got.shape == ref.shape and rel_err(got, ref) < 3e-3


@pytest.mark.parametrize("nd", [1, 2])
def test_real_conv_vd_gradients_and_fused_noise(nd):
    """real Conv1dVD / Conv2dVD (grouped): gradients vs float64 autograd over the oracle with
    injected noise; with in-kernel noise the backward regenerates the forward's draw"""
    from cplxmodule_b200.nn.relevance import Conv1dVD, Conv2dVD
    torch.manual_seed(40 + nd)
    if nd == 1:
        m = Conv1dVD(6, 4, 3, stride=2, padding=1, groups=2).to(DEV).train()
        x = torch.randn(3, 6, 19, device=DEV, requires_grad=True)
        ref_fn = orc.real_conv1d_vd
    else:
        m = Conv2dVD(4, 6, (2, 3), padding=(1, 1), dilation=(1, 2), groups=2).to(DEV).train()
        x = torch.randn(2, 4, 8, 9, device=DEV, requires_grad=True)
        ref_fn = orc.real_conv2d_vd
    with torch.no_grad():
        m.log_sigma2.uniform_(-5, 0)
        shape = m.eval()(x).shape
        m.train()
    c = torch.randn(shape, device=DEV)
    eps = torch.randn(shape, device=DEV)
    out = m(x, eps=eps)
    ((out * c).sum() + 0.7 * sum(penalties(m))).backward()
    d = lambda t: t.detach().double().cpu().requires_grad_()
    xr, w, b, l2 = map(d, (x, m.weight, m.bias, m.log_sigma2))
    o = ref_fn(xr, w, b, l2, eps.double().cpu(), m.stride, m.padding, m.dilation, m.groups)
    ((o * c.double().cpu()).sum() + 0.7 * kl_with_grad64("real_vd", w, None, l2)).backward()
    for got, ref in [(x.grad, xr.grad), (m.weight.grad, w.grad), (m.bias.grad, b.grad),
                     (m.log_sigma2.grad, l2.grad)]:
        assert rel_err(got, ref) < 3e-3      # real planes run on the tcgen05 kernel (tf32 operands)
    if m.groups == 1:
        return
    # ungrouped twin with the torch-exact in-kernel noise: same gradients as with that draw injected
    torch.manual_seed(5)
    m1 = (Conv1dVD(6, 4, 3, padding=1) if nd == 1 else Conv2dVD(4, 6, 3, padding=1)).to(DEV).train()
    xs = x.detach().clone().requires_grad_()
    grads = []
    for inject in (False, True):
        m1.zero_grad(); xs.grad = None
        torch.manual_seed(123)
        if inject:
            with torch.no_grad():
                e = torch.randn_like(m1.eval()(xs)); m1.train()
            y = m1(xs, eps=e)
        else:
            y = m1(xs)
        (y * y).sum().backward()
        grads.append([xs.grad.clone(), m1.weight.grad.clone(), m1.log_sigma2.grad.clone()])
    for a, b_ in zip(*grads):
        assert rel_err(a, b_) < 1e-5


def test_conv_vd_training_loop_reduces_loss():
    """drop-in use in a training loop (tests/test_relevance.py:62-73 shape): conv VD + linear VD"""
    from cplxmodule_b200.nn.relevance import CplxConv2dVD
    torch.manual_seed(0)
    conv = CplxConv2dVD(2, 4, 3, padding=1).to(DEV)
    head = CplxLinearVD(4 * 6 * 6, 3).to(DEV)
    params = list(conv.parameters()) + list(head.parameters())
    opt = torch.optim.Adam(params, lr=2e-2)
    z = cplx.randn(32, 2, 6, 6, device=DEV)
    target = torch.randn(32, 3, device=DEV)
    losses = []
    for _ in range(40):
        opt.zero_grad()
        h = conv(z)
        h = cplx.Cplx(h.real.reshape(32, -1), h.imag.reshape(32, -1))
        y = head(h)
        loss = ((y.real - target) ** 2).mean() + (y.imag ** 2).mean() \
            + 1e-4 * (sum(penalties(conv)) + sum(penalties(head)))
        loss.backward()
        opt.step()
        losses.append(float(loss))
    assert losses[-1] < 0.5 * losses[0]
