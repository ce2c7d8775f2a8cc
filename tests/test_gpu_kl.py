"""KL path on the GPU vs the oracle (all calls go package -> ctypes -> C ABI)."""
import pytest
import torch
import torch.nn.functional as F

from cplxmodule_b200 import _native as nv
from cplxmodule_b200 import ops
from cplxmodule_b200.nn.relevance import (CplxLinearARD, CplxLinearVD, LinearARD, LinearVD,
                                          named_penalties, penalties)
from oracle import cplx_oracle as orc
from tests.conftest import load_golden, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"
KIND = {"real_vd": nv.KL_REAL_VD, "real_ard": nv.KL_REAL_ARD, "cplx_vd": nv.KL_CPLX_VD,
        "cplx_ard": nv.KL_CPLX_ARD}


def truth64(kind, w_re, w_im, ls2):
    """float64 closed forms (softplus / sigmoid / cancellation-free Ein)."""
    w_re, ls2 = w_re.double().cpu(), ls2.double().cpu()
    if kind.startswith("cplx"):
        la = orc.log_alpha_cplx(w_re, w_im.double().cpu(), ls2)
    else:
        la = orc.log_alpha_real(w_re, ls2)
    if kind == "cplx_vd":
        return orc.penalty_cplx_vd_exact64(la)
    return orc.PENALTY[kind](la)


@pytest.mark.parametrize("kind", list(KIND))
def test_penalty_sweep_elementwise(kind):
    """log_alpha from -40 to 40 through every branch of every penalty."""
    la = torch.linspace(-40, 40, 1601)
    one, zero = torch.ones_like(la), torch.zeros_like(la)
    w_im = zero.to(DEV) if kind.startswith("cplx") else None
    got = ops.kl_penalty(KIND[kind], one.to(DEV), w_im, la.to(DEV), None).cpu().double()
    want = truth64(kind, one, zero, la)
    err = ((got - want).abs() / want.abs().clamp_min(1e-30)).max().item()
    assert err < 2e-5, err


def test_golden_sweep_matches_reference_where_reference_is_accurate():
    g = load_golden("penalty_sweep")
    la = g["log_sigma2"]
    one, zero = torch.ones_like(la).to(DEV), torch.zeros_like(la).to(DEV)
    for kind in ("real_vd", "real_ard", "cplx_ard"):
        got = ops.kl_penalty(KIND[kind], one, zero if kind.startswith("cplx") else None,
                             la.to(DEV), None).cpu()
        assert torch.allclose(got, g[kind], rtol=2e-5, atol=1e-7), kind
    got = ops.kl_penalty(KIND["cplx_vd"], one, zero, la.to(DEV), None).cpu()
    ok = la <= 8  # beyond that the reference's fp32 gamma + n - Ei(..) loses all digits
    assert torch.allclose(got[ok], g["cplx_vd"][ok], rtol=1e-3, atol=2e-6)


@pytest.mark.parametrize("name,kind", [("cplx_linear_vd", "cplx_vd"), ("cplx_linear_ard", "cplx_ard"),
                                       ("linear_vd", "real_vd"), ("linear_ard", "real_ard")])
def test_golden_layers(name, kind):
    g = load_golden(name)
    cplx = kind.startswith("cplx")
    w_re = (g["w_re"] if cplx else g["w"]).to(DEV)
    w_im = g["w_im"].to(DEV) if cplx else None
    ls2 = g["log_sigma2"].to(DEV)
    la = ops.log_alpha(w_re, w_im, ls2).cpu()
    assert torch.allclose(la, g["log_alpha"], rtol=1e-5, atol=1e-5)
    pen = ops.kl_penalty(KIND[kind], w_re, w_im, ls2, None).cpu()
    want = truth64(kind, w_re, w_im if cplx else None, ls2)
    assert rel_err(pen, want) < 1e-5
    s = ops.kl_penalty(KIND[kind], w_re, w_im, ls2, "sum").item()
    assert abs(s - g["penalty_sum"].item()) / abs(g["penalty_sum"].item()) < 1e-3   # vs reference fp32
    assert abs(s - want.sum().item()) / want.sum().item() < 1e-5                    # vs float64
    m = ops.kl_penalty(KIND[kind], w_re, w_im, ls2, "mean").item()
    assert abs(m - want.mean().item()) / want.mean().item() < 1e-5
    if "relevance" in g:
        mask = ops.log_alpha(w_re, w_im, ls2, threshold=3.0).cpu()
        assert torch.equal(mask, g["relevance"])


@pytest.mark.parametrize("n", [0, 1, 3, 5, 255, 1023, 4099])
def test_ragged_and_unaligned(n):
    torch.manual_seed(n)
    base_r, base_i = torch.randn(n + 3, device=DEV), torch.randn(n + 3, device=DEV)
    base_l = torch.empty(n + 3, device=DEV).uniform_(-12, 2)
    for off in (0, 1):  # off=1: planes start 4 bytes off a 16-byte boundary -> scalar path
        w_re, w_im, ls2 = base_r[off:off + n], base_i[off:off + n], base_l[off:off + n]
        want = truth64("cplx_vd", w_re, w_im, ls2)
        got = ops.kl_penalty(nv.KL_CPLX_VD, w_re, w_im, ls2, "sum").item()
        assert abs(got - want.sum().item()) <= 1e-5 * max(want.sum().item(), 1e-30) + 1e-30
        if n:
            pen = ops.kl_penalty(nv.KL_CPLX_VD, w_re, w_im, ls2, None).cpu()
            assert rel_err(pen, want) < 1e-5


def test_zero_weights_and_extremes():
    w = torch.tensor([0.0, 1e-30, 1e-6, 1.0, 1e6, -3.0], device=DEV)
    ls2 = torch.tensor([-10.0, 0.0, 5.0, -30.0, 20.0, 3.0], device=DEV)
    for kind in KIND:
        w_im = torch.zeros_like(w) if kind.startswith("cplx") else None
        got = ops.kl_penalty(KIND[kind], w, w_im, ls2, None).cpu().double()
        want = truth64(kind, w, w_im, ls2)
        assert torch.isfinite(got).all()
        assert torch.allclose(got, want, rtol=3e-5, atol=1e-30), (kind, got, want)


def test_bf16_planes():
    torch.manual_seed(5)
    w_re = torch.randn(300, 64, device=DEV).bfloat16()
    w_im = torch.randn(300, 64, device=DEV).bfloat16()
    ls2 = torch.empty(300, 64, device=DEV).uniform_(-12, 2).bfloat16()
    want = truth64("cplx_vd", w_re.float(), w_im.float(), ls2.float())
    s = ops.kl_penalty(nv.KL_CPLX_VD, w_re, w_im, ls2, "sum").float().item()
    assert abs(s - want.sum().item()) / want.sum().item() < 1e-2
    pen = ops.kl_penalty(nv.KL_CPLX_VD, w_re, w_im, ls2, None)
    assert pen.dtype == torch.bfloat16 and rel_err(pen.float(), want) < 1e-2


def test_full_size_properties():
    """BASELINE size (4096 x 4096): additivity over row shards and mean == sum / n;
    a 64-row sample is checked elementwise against float64."""
    torch.manual_seed(11)
    N = K = 4096
    w_re = torch.empty(N, K, device=DEV).uniform_(-0.011, 0.011)
    w_im = torch.empty(N, K, device=DEV).uniform_(-0.011, 0.011)
    ls2 = torch.empty(N, K, device=DEV).uniform_(-12, 2)
    total = ops.kl_penalty(nv.KL_CPLX_VD, w_re, w_im, ls2, "sum").item()
    parts = sum(ops.kl_penalty(nv.KL_CPLX_VD, w_re[i:i + 512], w_im[i:i + 512], ls2[i:i + 512],
                               "sum").item() for i in range(0, N, 512))
    assert abs(total - parts) / total < 1e-6
    mean = ops.kl_penalty(nv.KL_CPLX_VD, w_re, w_im, ls2, "mean").item()
    assert abs(mean - total / (N * K)) / mean < 1e-6
    again = ops.kl_penalty(nv.KL_CPLX_VD, w_re, w_im, ls2, "sum").item()
    assert again == total  # deterministic reduction order
    rows = torch.randperm(N)[:64].to(DEV)
    pen = ops.kl_penalty(nv.KL_CPLX_VD, w_re[rows], w_im[rows], ls2[rows], None)
    assert rel_err(pen, truth64("cplx_vd", w_re[rows], w_im[rows], ls2[rows])) < 1e-5


def test_module_level_penalties():
    torch.manual_seed(3)
    net = torch.nn.ModuleDict({"a": CplxLinearVD(40, 24), "b": CplxLinearARD(24, 10),
                               "c": LinearVD(10, 7), "d": LinearARD(7, 3)}).to(DEV)
    for m in net.values():
        with torch.no_grad():
            m.log_sigma2.uniform_(-12, 2)
    got = {k: v.item() for k, v in named_penalties(net)}
    kinds = {"a": "cplx_vd", "b": "cplx_ard", "c": "real_vd", "d": "real_ard"}
    for k, m in net.items():
        w = m.weight
        cplx = kinds[k].startswith("cplx")
        want = truth64(kinds[k], w.real if cplx else w, w.imag if cplx else None, m.log_sigma2)
        assert abs(got[k] - want.sum().item()) / want.sum().item() < 1e-5
        assert rel_err(m.penalty, want) < 1e-5
        assert m.penalty.shape == m.log_sigma2.shape
        assert torch.equal(m.relevance(threshold=1.0).cpu(),
                           (m.log_alpha.cpu() <= 1.0).float())
    total = sum(penalties(net))
    assert total.shape == () and abs(total.item() - sum(got.values())) < 1e-3 * total.item()


# ------------------------------------------------ extension penalties (extensions/complex.py)
@pytest.mark.parametrize("name,kind,cls_name", [("approx", "cplx_vd_approx", "CplxLinearVDApprox"),
                                                ("scalefree", "cplx_vd_scalefree", "CplxLinearVDScaleFree")])
def test_extension_penalties(name, kind, cls_name):
    """golden values of the live reference, float64 oracle on a wide log_alpha range, gradients
    against float64 autograd over the oracle, and the pre-pass by-product of the forward"""
    import cplxmodule_b200 as cb
    from cplxmodule_b200 import cplx
    from cplxmodule_b200.nn.relevance import extensions, penalties
    from tests.conftest import load_golden
    g = load_golden("ext_penalties")
    w_re, w_im, ls2 = g[f"{name}_w_re"], g[f"{name}_w_im"], g[f"{name}_log_sigma2"]
    cls = getattr(extensions, cls_name)
    m = cls(w_re.shape[1], w_re.shape[0])
    m.load_state_dict({"weight.real": w_re, "weight.imag": w_im, "bias.real": torch.zeros(w_re.shape[0]),
                       "bias.imag": torch.zeros(w_re.shape[0]), "log_sigma2": ls2})
    m = m.to(DEV)
    want = orc.layer_penalty(kind, w_re.double(), w_im.double(), ls2.double(), None)
    got = m.penalty.double().cpu()
    assert float((got - want).abs().max()) <= 2e-5 * float(want.abs().max())
    ref32 = g[f"{name}_penalty"].double()
    assert float((got - ref32).abs().max()) <= 1e-3 * float(ref32.abs().max())
    s = float(sum(penalties(m)))
    assert abs(s - float(want.sum())) <= 1e-5 * abs(float(want.sum()))
    # gradients
    p64 = [t.double().clone().requires_grad_(True) for t in (w_re, w_im, ls2)]
    orc.layer_penalty(kind, *p64, "sum").backward()
    m.zero_grad()
    sum(penalties(m)).backward()
    for got_g, want_g in zip((m.weight.real.grad, m.weight.imag.grad, m.log_sigma2.grad), p64):
        assert float((got_g.double().cpu() - want_g.grad).abs().max()) <= 2e-4 * float(want_g.grad.abs().max())
    # fused into the forward's operand pre-pass
    torch.manual_seed(0)
    big = cls(256, 136).to(DEV).train()
    with torch.no_grad():
        big.log_sigma2.uniform_(-10, 3)
        big(cplx.randn(300, 256, device=DEV))
        assert big._kl_cache._entry is not None
        fused = float(sum(penalties(big)))
        alone = float(sum(penalties(big)))
    assert abs(fused - alone) <= 1e-6 * abs(alone)
