"""SURVEY section 8(f) rows finished in round 2, on the GPU through the C ABI:
f2  masked linear layers with the mask applied inside the operand pre-pass (nn/masked/complex.py:33-35),
    KL sum + relevance mask from one pass (relevance/base.py:192-216);
f3  grouped variational convolutions with the layer's single torch-exact noise draw
    (nn/relevance/complex/base.py:120-135 with groups > 1);
f4  bilinear layers (cplx.py:1062-1090, complex/base.py:59-84, real/base.py:52-80), `*VDBogus`
    names and the reference's class hierarchy (extensions/complex.py:47-198)."""
import pytest
import torch
import torch.nn.functional as F

import cplxmodule_b200 as cb
from cplxmodule_b200 import _native as nv
from cplxmodule_b200 import cplx, ops
from cplxmodule_b200.nn import masked
from cplxmodule_b200.nn import relevance as rel
from cplxmodule_b200.nn.modules import CplxBilinear
from cplxmodule_b200.nn.relevance import extensions as ext
from oracle import cplx_oracle as orc
from tests.conftest import load_golden, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"
c64 = lambda t: t.detach().cpu().double()


# ------------------------------------------------------------------------------ f2
@pytest.mark.parametrize("M,N,K", [(4, 8, 8), (130, 129, 36), (257, 200, 264), (513, 384, 1000),
                                   (300, 130, 4096 + 8)])
@pytest.mark.parametrize("cplx_", [True, False])
@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_masked_linear_fused_mask(M, N, K, cplx_, dt):
    """M > 128, K >= 64, K % 8 == 0 (fp32): mask multiplied inside the operand pre-pass; every
    other shape / dtype: one elementwise launch inside the same C call."""
    torch.manual_seed(M + N + K)
    mk = lambda *s: torch.randn(*s, device=DEV).to(dt)
    x_re, x_im = mk(M, K), mk(M, K)
    w_re, w_im = mk(N, K) / K ** 0.5, mk(N, K) / K ** 0.5
    b_re, b_im = mk(N), mk(N)
    mask = (torch.rand(N, K, device=DEV) < 0.3).to(dt)
    tol = 1e-3 if dt == torch.float32 else 1e-2
    c = lambda t: t.float().cpu().double()
    if cplx_:
        got = ops.cplx_linear_masked(x_re, x_im, w_re, w_im, mask, b_re, b_im)
        want = orc.cplx_linear(c(x_re), c(x_im), c(w_re) * c(mask), c(w_im) * c(mask), c(b_re), c(b_im))
        assert rel_err(got[0].float(), want[0]) < tol and rel_err(got[1].float(), want[1]) < tol
    else:
        got = ops.real_linear_masked(x_re, w_re, mask, b_re)
        want = F.linear(c(x_re), c(w_re) * c(mask), c(b_re))
        assert rel_err(got.float(), want) < tol


def test_masked_linear_fused_equals_materialised_and_backward():
    """same kernel fed with weight * mask gives the SAME bits as the fused mask; gradients against
    float64 autograd over the oracle (dW carries the mask, dx uses the masked weight)"""
    torch.manual_seed(7)
    M, N, K = 384, 256, 512
    layer = masked.CplxLinearMasked(K, N).to(DEV)
    layer.mask = (torch.rand(N, K, device=DEV) < 0.4).float()
    z = cplx.randn(M, K, device=DEV)
    out = layer(z)
    w = layer.weight
    ref = ops.cplx_linear(z.real, z.imag, (w.real * layer.mask).detach(), (w.imag * layer.mask).detach(),
                          layer.bias.real, layer.bias.imag)
    assert torch.equal(out.real, ref[0]) and torch.equal(out.imag, ref[1])
    zr, zi = z.real.clone().requires_grad_(), z.imag.clone().requires_grad_()
    o = layer(cplx.Cplx(zr, zi))
    g_re, g_im = torch.randn_like(o.real), torch.randn_like(o.imag)
    (o.real * g_re + o.imag * g_im).sum().backward()
    W_re, W_im = c64(w.real).requires_grad_(), c64(w.imag).requires_grad_()
    X_re, X_im = c64(z.real).requires_grad_(), c64(z.imag).requires_grad_()
    mk = c64(layer.mask)
    want = orc.cplx_linear(X_re, X_im, W_re * mk, W_im * mk, c64(layer.bias.real), c64(layer.bias.imag))
    (want[0] * c64(g_re) + want[1] * c64(g_im)).sum().backward()
    assert rel_err(zr.grad, X_re.grad) < 2e-3 and rel_err(zi.grad, X_im.grad) < 2e-3
    assert rel_err(w.real.grad, W_re.grad) < 2e-3 and rel_err(w.imag.grad, W_im.grad) < 2e-3
    assert float(w.real.grad[layer.mask == 0].abs().max()) == 0.0


def test_real_masked_layers_and_conv_masked():
    torch.manual_seed(8)
    lin = masked.LinearMasked(96, 40).to(DEV)
    lin.mask = (torch.rand(40, 96, device=DEV) < 0.5).float()
    x = torch.randn(200, 96, device=DEV)
    want = F.linear(c64(x), c64(lin.weight) * c64(lin.mask), c64(lin.bias))
    assert rel_err(lin(x), want) < 1e-3
    conv = masked.Conv2dMasked(6, 8, 3, padding=1).to(DEV)
    conv.mask = (torch.rand_like(conv.weight) < 0.5).float()
    xi = torch.randn(3, 6, 12, 10, device=DEV)
    want = F.conv2d(c64(xi), c64(conv.weight) * c64(conv.mask), c64(conv.bias), padding=1)
    assert rel_err(conv(xi), want) < 1e-3      # real planes run on the tcgen05 kernel (tf32 operands)
    bil = masked.CplxBilinearMasked(5, 6, 7).to(DEV)
    bil.mask = (torch.rand(7, 5, 6, device=DEV) < 0.5).float()
    z1, z2 = cplx.randn(9, 5, device=DEV), cplx.randn(9, 6, device=DEV)
    out = bil(z1, z2)
    w, b = bil.weight, bil.bias
    want = orc.cplx_bilinear(c64(z1.real), c64(z1.imag), c64(z2.real), c64(z2.imag),
                             c64(w.real) * c64(bil.mask), c64(w.imag) * c64(bil.mask), c64(b.real),
                             c64(b.imag))
    assert rel_err(out.real, want[0]) < 1e-3 and rel_err(out.imag, want[1]) < 1e-3


@pytest.mark.parametrize("kind,cplx_", [(nv.KL_CPLX_VD, True), (nv.KL_CPLX_ARD, True), (nv.KL_REAL_VD, False)])
def test_kl_and_mask_one_pass(kind, cplx_):
    torch.manual_seed(9)
    N, K = 301, 257
    w_re = torch.randn(N, K, device=DEV) * 0.05
    w_im = torch.randn(N, K, device=DEV) * 0.05 if cplx_ else None
    ls2 = torch.empty(N, K, device=DEV).uniform_(-12, 2)
    kl, mask = ops.kl_and_mask(kind, w_re, w_im, ls2, threshold=1.5)
    assert torch.equal(mask, ops.log_alpha(w_re, w_im, ls2, threshold=1.5))
    assert torch.equal(kl, ops.kl(kind, w_re, w_im, ls2, "sum"))
    assert 0.0 < float(mask.mean()) < 1.0


# ------------------------------------------------------------------------------ f4
@pytest.mark.parametrize("tag,conj", [("conj", True), ("plain", False)])
def test_golden_cplx_bilinear_vd(tag, conj):
    g = load_golden("bilinear")
    d = lambda n: g[f"{tag}_{n}"].to(DEV)
    m = rel.CplxBilinearVD(7, 10, 9, conjugate=conj)
    m.load_state_dict({"weight.real": g[f"{tag}_w_re"], "weight.imag": g[f"{tag}_w_im"],
                       "bias.real": g[f"{tag}_b_re"], "bias.imag": g[f"{tag}_b_im"],
                       "log_sigma2": g[f"{tag}_log_sigma2"]})
    m = m.to(DEV)
    z1, z2 = cplx.Cplx(d("x1_re"), d("x1_im")), cplx.Cplx(d("x2_re"), d("x2_im"))
    with torch.no_grad():
        mu = m.eval()(z1, z2)
        y = m.train()(z1, z2, eps=cplx.Cplx(d("eps_re"), d("eps_im")))
        kl = sum(rel.penalties(m))
    assert rel_err(mu.real, g[f"{tag}_mu_re"]) < 1e-3 and rel_err(mu.imag, g[f"{tag}_mu_im"]) < 1e-3
    assert rel_err(y.real, g[f"{tag}_y_re"]) < 1e-3 and rel_err(y.imag, g[f"{tag}_y_im"]) < 1e-3
    assert abs(kl.item() - g[f"{tag}_penalty_sum"].item()) < 1e-3 * abs(g[f"{tag}_penalty_sum"].item())


def test_golden_real_bilinear_ard():
    g = load_golden("bilinear")
    m = rel.BilinearARD(6, 5, 8)
    m.load_state_dict({"weight": g["real_w"], "bias": g["real_b"], "log_sigma2": g["real_log_sigma2"]})
    m = m.to(DEV)
    x1, x2 = g["real_x1"].to(DEV), g["real_x2"].to(DEV)
    with torch.no_grad():
        mu = m.eval()(x1, x2)
        y = m.train()(x1, x2, eps=g["real_eps"].to(DEV))
        kl = sum(rel.penalties(m))
    assert rel_err(mu, g["real_mu"]) < 1e-3 and rel_err(y, g["real_y"]) < 1e-3
    assert abs(kl.item() - g["real_penalty_sum"].item()) < 1e-3 * abs(g["real_penalty_sum"].item())


@pytest.mark.parametrize("conj", [True, False])
def test_bilinear_sizes_noise_and_gradients(conj):
    """in1 * in2 = 2304 features, batch 300 (the tcgen05 linear kernels); the fused noise equals
    the reference's draw on this device; gradients vs float64 autograd over the oracle"""
    torch.manual_seed(11)
    B, d1, d2, O = 300, 48, 48, 72
    m = rel.CplxBilinearVD(d1, d2, O, conjugate=conj).to(DEV).train()
    with torch.no_grad():
        m.log_sigma2.uniform_(-8, 0)
    z1, z2 = cplx.randn(B, d1, device=DEV), cplx.randn(B, d2, device=DEV)
    gen = torch.cuda.default_generators[torch.cuda.current_device()]
    torch.manual_seed(5)
    state = gen.get_state()
    with torch.no_grad():
        fused = m(z1, z2)
    gen.set_state(state)
    eps = cplx.randn(B, O, device=DEV)
    with torch.no_grad():
        inject = m(z1, z2, eps=eps)
    assert torch.equal(fused.real, inject.real) and torch.equal(fused.imag, inject.imag)
    w, b = m.weight, m.bias
    want = orc.cplx_bilinear_vd(c64(z1.real), c64(z1.imag), c64(z2.real), c64(z2.imag), c64(w.real),
                                c64(w.imag), c64(b.real), c64(b.imag), c64(m.log_sigma2), c64(eps.real),
                                c64(eps.imag), conj)
    assert rel_err(inject.real, want[0]) < 1e-3 and rel_err(inject.imag, want[1]) < 1e-3
    # gradients
    a_re, a_im = z1.real.clone().requires_grad_(), z1.imag.clone().requires_grad_()
    u_re, u_im = z2.real.clone().requires_grad_(), z2.imag.clone().requires_grad_()
    out = m(cplx.Cplx(a_re, a_im), cplx.Cplx(u_re, u_im), eps=eps)
    g_re, g_im = torch.randn_like(out.real), torch.randn_like(out.imag)
    (out.real * g_re + out.imag * g_im).sum().backward()
    leaves = [c64(t).requires_grad_() for t in (z1.real, z1.imag, z2.real, z2.imag, w.real, w.imag, m.log_sigma2)]
    ref = orc.cplx_bilinear_vd(*leaves[:6], c64(b.real), c64(b.imag), leaves[6], c64(eps.real), c64(eps.imag), conj)
    (ref[0] * c64(g_re) + ref[1] * c64(g_im)).sum().backward()
    for mine, theirs in zip((a_re, a_im, u_re, u_im, w.real, w.imag, m.log_sigma2), leaves):
        assert rel_err(mine.grad, theirs.grad) < 3e-3


def test_bilinear_broadcast_and_plain_module():
    torch.manual_seed(12)
    m = CplxBilinear(4, 6, 5).to(DEV)
    z1, z2 = cplx.randn(2, 3, 4, device=DEV), cplx.randn(2, 3, 6, device=DEV)
    out = m(z1, z2)
    assert out.shape == (2, 3, 5)
    w, b = m.weight, m.bias
    want = orc.cplx_bilinear(c64(z1.real), c64(z1.imag), c64(z2.real), c64(z2.imag), c64(w.real),
                             c64(w.imag), c64(b.real), c64(b.imag), True)
    assert rel_err(out.real, want[0]) < 1e-3 and rel_err(out.imag, want[1]) < 1e-3


def test_reference_class_hierarchy_and_bogus_names():
    """extensions/complex.py:47-198, complex/ard.py:42-74, real/ard.py:42-66"""
    assert issubclass(rel.CplxLinearARD, rel.CplxLinearVD) and issubclass(rel.LinearARD, rel.LinearVD)
    assert issubclass(rel.CplxConv2dARD, rel.CplxConv2dVD) and issubclass(rel.Conv1dARD, rel.Conv1dVD)
    for name in ("CplxLinearVDApprox", "CplxLinearVDScaleFree", "CplxLinearVDBogus"):
        assert issubclass(getattr(ext, name), rel.CplxLinearVD)
    assert issubclass(ext.CplxBilinearVDBogus, rel.CplxBilinearVD)
    assert issubclass(ext.CplxConv2dVDBogus, rel.CplxConv2dVD)
    torch.manual_seed(13)
    bogus, exact = ext.CplxLinearVDBogus(40, 24).to(DEV), rel.CplxLinearVD(40, 24).to(DEV)
    exact.load_state_dict(bogus.state_dict())
    assert torch.equal(sum(rel.penalties(bogus)), sum(rel.penalties(exact)))


def test_log_alpha_is_differentiable():
    """the reference's `.log_alpha` property carries a graph (complex/base.py:27-31)"""
    torch.manual_seed(14)
    m = rel.CplxLinearVD(33, 21).to(DEV)
    with torch.no_grad():
        m.log_sigma2.uniform_(-6, 2)
    la = m.log_alpha
    g = torch.randn_like(la)
    (la * g).sum().backward()
    w_re, w_im, ls2 = (c64(t).requires_grad_() for t in (m.weight.real, m.weight.imag, m.log_sigma2))
    (orc.log_alpha_cplx(w_re, w_im, ls2) * c64(g)).sum().backward()
    assert rel_err(m.weight.real.grad, w_re.grad) < 1e-4 and rel_err(m.log_sigma2.grad, ls2.grad) < 1e-6


def test_inplace_update_between_forward_and_backward_raises():
    """inputs are registered with autograd's version tracking (save_for_backward)"""
    torch.manual_seed(15)
    m = rel.CplxLinearVD(64, 32).to(DEV).train()
    z = cplx.randn(16, 64, device=DEV)
    out = m(z)
    with torch.no_grad():
        m.weight.real.add_(1.0)
    with pytest.raises(RuntimeError, match="modified by an inplace operation"):
        (out.real.sum() + out.imag.sum()).backward()


# ------------------------------------------------------------------------------ f3
@pytest.mark.parametrize("cplx_", [True, False])
def test_grouped_variational_conv_draws_the_layers_single_noise(cplx_):
    """groups > 1, noise drawn internally (default torch-exact mode): equals the same layer fed
    with the ONE draw the reference makes for the whole output; trainable."""
    torch.manual_seed(16)
    cls = rel.CplxConv2dVD if cplx_ else rel.Conv2dVD
    m = cls(8, 12, 3, padding=1, groups=2).to(DEV).train()
    with torch.no_grad():
        m.log_sigma2.uniform_(-6, 1)
    x = cplx.randn(3, 8, 14, 11, device=DEV) if cplx_ else torch.randn(3, 8, 14, 11, device=DEV)
    gen = torch.cuda.default_generators[torch.cuda.current_device()]
    torch.manual_seed(77)
    state = gen.get_state()
    with torch.no_grad():
        fused = m(x)
    off = gen.get_offset()
    gen.set_state(state)
    eps = cplx.randn(3, 12, 14, 11, device=DEV) if cplx_ else torch.randn(3, 12, 14, 11, device=DEV)
    assert gen.get_offset() == off
    with torch.no_grad():
        inject = m(x, eps=eps)
    if cplx_:
        assert torch.equal(fused.real, inject.real) and torch.equal(fused.imag, inject.imag)
    else:
        assert torch.equal(fused, inject)
        want = orc.real_conv2d_vd(c64(x), c64(m.weight), c64(m.bias), c64(m.log_sigma2), c64(eps), 1, 1, 1, 2)
        assert rel_err(inject, want) < 1e-3
    out = m(x)                                                  # and it trains
    loss = (out.real.square().mean() + out.imag.square().mean()) if cplx_ else out.square().mean()
    (loss + 1e-3 * sum(rel.penalties(m))).backward()
    assert m.log_sigma2.grad is not None and torch.isfinite(m.log_sigma2.grad).all()


def test_fast_noise_conv_layer_trains():
    """set_noise_mode('fast') + autograd on a conv VD layer no longer raises in backward"""
    torch.manual_seed(17)
    m = rel.CplxConv2dVD(8, 8, 3, padding=1).to(DEV).train()
    x = cplx.randn(2, 8, 12, 12, device=DEV)
    cb.set_noise_mode("fast")
    try:
        out = m(x)
        (out.real.square().mean() + out.imag.square().mean()).backward()
    finally:
        cb.set_noise_mode("torch")
    assert torch.isfinite(m.weight.real.grad).all() and torch.isfinite(m.log_sigma2.grad).all()
