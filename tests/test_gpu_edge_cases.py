"""Edge cases of the hot path as the reference (torch) handles them: empty batches, single rows /
columns, leading batch dimensions, reductions of `penalties`, eval-mode forwards."""
import pytest
import torch
import torch.nn.functional as F

import cplxmodule_b200 as cb
from cplxmodule_b200 import cplx
from cplxmodule_b200.nn import CplxConv1d, CplxConv2d, CplxLinear
from cplxmodule_b200.nn import relevance as rel
from oracle import cplx_oracle as orc
from tests.conftest import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"
c64 = lambda t: t.detach().double().cpu()


@pytest.mark.parametrize("cls", [CplxLinear, rel.CplxLinearVD, rel.CplxLinearARD])
def test_empty_batch_linear(cls):
    """F.linear / cplx.linear of a (0, K) input is a (0, N) output; the KL does not depend on it"""
    layer = cls(64, 24).to(DEV).train()
    z = cplx.Cplx(torch.empty(0, 64, device=DEV), torch.empty(0, 64, device=DEV))
    out = layer(z)
    assert out.real.shape == (0, 24) and out.imag.shape == (0, 24)
    z3 = cplx.Cplx(torch.empty(3, 0, 64, device=DEV), torch.empty(3, 0, 64, device=DEV))
    assert layer(z3).shape == (3, 0, 24)
    if cls is not CplxLinear:
        kl = sum(rel.penalties(layer))
        w = layer.weight
        want = orc.layer_penalty("cplx_vd" if cls is rel.CplxLinearVD else "cplx_ard", c64(w.real), c64(w.imag),
                                 c64(layer.log_sigma2), "sum")
        assert abs(kl.item() - want.item()) <= 1e-4 * abs(want.item())


def test_empty_batch_real_and_conv():
    lin = rel.LinearVD(40, 16).to(DEV).train()
    assert lin(torch.empty(0, 40, device=DEV)).shape == (0, 16)
    conv = rel.CplxConv2dVD(4, 6, 3, padding=1).to(DEV).train()
    z = cplx.Cplx(torch.empty(0, 4, 9, 9, device=DEV), torch.empty(0, 4, 9, 9, device=DEV))
    assert conv(z).shape == (0, 6, 9, 9)
    c1 = CplxConv1d(4, 6, 3).to(DEV)
    z1 = cplx.Cplx(torch.empty(0, 4, 20, device=DEV), torch.empty(0, 4, 20, device=DEV))
    assert c1(z1).shape == (0, 6, 18)
    rconv = rel.Conv2dVD(4, 6, 3).to(DEV).train()
    assert rconv(torch.empty(0, 4, 9, 9, device=DEV)).shape == (0, 6, 7, 7)


@pytest.mark.parametrize("M,N,K", [(1, 1, 8), (1, 200, 16), (130, 1, 64), (2, 3, 4), (1, 1, 1)])
def test_degenerate_shapes_vs_oracle(M, N, K):
    """single rows / columns / a 1 x 1 x 1 layer (K * 4 % 16 != 0 goes to the exact kernel)"""
    torch.manual_seed(M * 7 + N * 3 + K)
    layer = rel.CplxLinearVD(K, N).to(DEV).train()
    with torch.no_grad():
        layer.log_sigma2.uniform_(-8, 0)
    z, eps = cplx.randn(M, K, device=DEV), cplx.randn(M, N, device=DEV)
    with torch.no_grad():
        out = layer(z, eps=eps)
        fused = layer(z)
    w, b = layer.weight, layer.bias
    want = orc.cplx_linear_vd(c64(z.real), c64(z.imag), c64(w.real), c64(w.imag), c64(b.real), c64(b.imag),
                              c64(layer.log_sigma2), c64(eps.real), c64(eps.imag))
    assert rel_err(out.real, want[0]) < 1e-3 and rel_err(out.imag, want[1]) < 1e-3
    assert fused.shape == out.shape and torch.isfinite(fused.real).all()


def test_leading_batch_dims_and_eval_mode():
    torch.manual_seed(2)
    layer = rel.CplxLinearVD(48, 20).to(DEV)
    z = cplx.randn(3, 5, 7, 48, device=DEV)
    with torch.no_grad():
        mu = layer.eval()(z)
    w, b = layer.weight, layer.bias
    want = orc.cplx_linear(c64(z.real).reshape(-1, 48), c64(z.imag).reshape(-1, 48), c64(w.real), c64(w.imag),
                           c64(b.real), c64(b.imag))
    assert mu.shape == (3, 5, 7, 20)
    assert rel_err(mu.real.reshape(-1, 20), want[0]) < 1e-3 and rel_err(mu.imag.reshape(-1, 20), want[1]) < 1e-3
    with torch.no_grad():
        assert torch.equal(layer(z).real, mu.real)             # eval: no noise, deterministic
        out = layer.train()(z)
    assert out.shape == mu.shape and not torch.equal(out.real, mu.real)


@pytest.mark.parametrize("reduction", ["sum", "mean", None])
def test_penalties_reductions_match_reference_semantics(reduction):
    """named_penalties(reduction=...) (nn/relevance/base.py:88-141): 0-d sums / means, or the
    full per-weight tensor; a bad reduction raises ValueError before any module is visited"""
    torch.manual_seed(3)
    net = torch.nn.Sequential(rel.CplxLinearVD(32, 16), rel.CplxLinearARD(16, 8)).to(DEV).train()
    vals = list(rel.penalties(net, reduction=reduction))
    assert len(vals) == 2
    kinds = ("cplx_vd", "cplx_ard")
    for v, layer, kind in zip(vals, net, kinds):
        w = layer.weight
        want = orc.layer_penalty(kind, c64(w.real), c64(w.imag), c64(layer.log_sigma2), reduction)
        assert tuple(v.shape) == tuple(want.shape)
        assert rel_err(v, want) < 1e-4
    names = [n for n, _ in rel.named_penalties(net, reduction=reduction, prefix="net")]
    assert names == ["net.0", "net.1"]
    with pytest.raises(ValueError):
        list(rel.penalties(net, reduction="max"))


def test_shape_mismatch_and_cpu_inputs_raise():
    layer = rel.CplxLinearVD(32, 16).to(DEV).train()
    with pytest.raises(RuntimeError, match="size mismatch"):
        layer(cplx.randn(4, 31, device=DEV))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        layer(cplx.randn(4, 32))
    conv = CplxConv2d(4, 6, 3).to(DEV)
    with pytest.raises(RuntimeError):
        conv(cplx.randn(2, 5, 8, 8, device=DEV))
    with pytest.raises(RuntimeError):
        conv(cplx.randn(2, 4, 2, 2, device=DEV))               # kernel larger than the input
