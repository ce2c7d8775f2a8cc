"""Linear / variational-linear forward on the GPU vs the oracle and the golden fixtures."""
import pytest
import torch

import cplxmodule_b200 as cb
from cplxmodule_b200 import _native as nv
from cplxmodule_b200 import cplx, ops
from cplxmodule_b200.nn import CplxLinear
from cplxmodule_b200.nn.relevance import CplxLinearARD, CplxLinearVD, LinearVD
from oracle import cplx_oracle as orc
from tests.conftest import load_golden, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = {"simt": 2e-5, "tensor": 1e-3}   # north_star: 1e-3 rel fp32(tf32 operands), 1e-2 bf16


@pytest.fixture(params=["simt", "tensor"])
def math(request):
    ops.set_math_mode(request.param)
    yield request.param
    ops.set_math_mode("auto")


def cuda(g, *names):
    return [g[n].to(DEV) for n in names]


def test_golden_cplx_linear(math):
    g = load_golden("cplx_linear")
    re, im = ops.cplx_linear(*cuda(g, "x_re", "x_im", "w_re", "w_im", "b_re", "b_im"))
    assert rel_err(re, g["y_re"]) < TOL[math] and rel_err(im, g["y_im"]) < TOL[math]


def test_golden_cplx_linear_vd(math):
    g = load_golden("cplx_linear_vd")
    args = cuda(g, "x_re", "x_im", "w_re", "w_im", "b_re", "b_im", "log_sigma2")
    re, im = ops.cplx_linear_vd(*args, eps=tuple(cuda(g, "eps_re", "eps_im")))
    assert rel_err(re, g["y_re"]) < TOL[math] and rel_err(im, g["y_im"]) < TOL[math]


def test_golden_real_linear_vd(math):
    g = load_golden("linear_vd")
    y = ops.real_linear_vd(*cuda(g, "x", "w", "b", "log_sigma2"), eps=g["eps"].to(DEV))
    assert rel_err(y, g["y"]) < TOL[math]


SHAPES = [(1, 4, 4), (5, 3, 8), (128, 256, 784), (130, 129, 36), (257, 200, 260), (64, 64, 4096),
          (300, 384, 1000)]


@pytest.mark.parametrize("M,N,K", SHAPES)
@pytest.mark.parametrize("cplx_", [True, False])
def test_shapes_vs_oracle(M, N, K, cplx_, math):
    """ragged M/N/K tails (TMA zero fill, partial tiles), real and complex, bias on."""
    torch.manual_seed(M * 7 + N * 3 + K)
    x_re, x_im = torch.randn(M, K), torch.randn(M, K)
    w_re, w_im = torch.randn(N, K) / K ** 0.5, torch.randn(N, K) / K ** 0.5
    b_re, b_im = torch.randn(N), torch.randn(N)
    ls2 = torch.empty(N, K).uniform_(-8, 1)
    eps_re, eps_im = torch.randn(M, N), torch.randn(M, N)
    d = lambda t: t.to(DEV)
    if cplx_:
        want = orc.cplx_linear_vd(x_re.double(), x_im.double(), w_re.double(), w_im.double(),
                                  b_re.double(), b_im.double(), ls2.double(), eps_re.double(),
                                  eps_im.double())
        got = ops.cplx_linear_vd(d(x_re), d(x_im), d(w_re), d(w_im), d(b_re), d(b_im), d(ls2),
                                 eps=(d(eps_re), d(eps_im)))
        assert rel_err(got[0], want[0]) < TOL[math] and rel_err(got[1], want[1]) < TOL[math]
        mu = ops.cplx_linear(d(x_re), d(x_im), d(w_re), d(w_im), None, None)
        want_mu = orc.cplx_linear(x_re.double(), x_im.double(), w_re.double(), w_im.double())
        assert rel_err(mu[0], want_mu[0]) < TOL[math] and rel_err(mu[1], want_mu[1]) < TOL[math]
    else:
        want = orc.real_linear_vd(x_re.double(), w_re.double(), b_re.double(), ls2.double(),
                                  eps_re.double())
        got = ops.real_linear_vd(d(x_re), d(w_re), d(b_re), d(ls2), eps=d(eps_re))
        assert rel_err(got, want) < TOL[math]
        mu = ops.real_linear(d(x_re), d(w_re), d(b_re))
        assert rel_err(mu, torch.nn.functional.linear(x_re.double(), w_re.double(),
                                                      b_re.double())) < TOL[math]


def test_unaligned_k_routes_to_simt_in_auto_mode():
    torch.manual_seed(1)
    M, N, K = 33, 17, 13            # K*4 % 16 != 0: the TMA path cannot take it
    x_re, x_im, w_re, w_im = (torch.randn(M, K), torch.randn(M, K), torch.randn(N, K),
                              torch.randn(N, K))
    got = ops.cplx_linear(*(t.to(DEV) for t in (x_re, x_im, w_re, w_im)))
    want = orc.cplx_linear(x_re, x_im, w_re, w_im)
    assert rel_err(got[0], want[0]) < 2e-5 and rel_err(got[1], want[1]) < 2e-5
    ops.set_math_mode("tensor")
    try:
        with pytest.raises(RuntimeError, match="alignment"):
            ops.cplx_linear(*(t.to(DEV) for t in (x_re, x_im, w_re, w_im)))
    finally:
        ops.set_math_mode("auto")


def test_leading_batch_dims_and_noncontiguous_inputs():
    torch.manual_seed(2)
    lin = CplxLinear(64, 32).to(DEV)
    z = cplx.Cplx(torch.randn(2, 3, 5, 64, device=DEV), torch.randn(2, 3, 5, 64, device=DEV))
    out = lin(z)
    assert out.shape == (2, 3, 5, 32)
    want = orc.cplx_linear(z.real.cpu(), z.imag.cpu(), lin.weight.real.cpu(), lin.weight.imag.cpu(),
                           lin.bias.real.cpu(), lin.bias.imag.cpu())
    assert rel_err(out.real, want[0]) < 1e-3
    inter = torch.randn(7, 128, device=DEV)
    zi = cplx.from_interleaved_real(inter, copy=False)   # stride-2 views
    out = lin(zi)
    want = orc.cplx_linear(inter[:, 0::2].cpu(), inter[:, 1::2].cpu(), lin.weight.real.cpu(),
                           lin.weight.imag.cpu(), lin.bias.real.cpu(), lin.bias.imag.cpu())
    assert rel_err(out.imag, want[1]) < 1e-3


def test_bf16_path():
    torch.manual_seed(4)
    M, N, K = 192, 160, 512
    bf = lambda t: t.to(DEV).bfloat16()
    x_re, x_im = bf(torch.randn(M, K)), bf(torch.randn(M, K))
    w_re, w_im = bf(torch.randn(N, K) / K ** 0.5), bf(torch.randn(N, K) / K ** 0.5)
    b_re, b_im = bf(torch.randn(N)), bf(torch.randn(N))
    ls2 = bf(torch.empty(N, K).uniform_(-8, 1))
    eps = (bf(torch.randn(M, N)), bf(torch.randn(M, N)))
    c = lambda t: t.float().cpu().double()
    want = orc.cplx_linear_vd(c(x_re), c(x_im), c(w_re), c(w_im), c(b_re), c(b_im), c(ls2),
                              c(eps[0]), c(eps[1]))
    for mode in ("tensor", "simt"):
        ops.set_math_mode(mode)
        try:
            got = ops.cplx_linear_vd(x_re, x_im, w_re, w_im, b_re, b_im, ls2, eps=eps)
        finally:
            ops.set_math_mode("auto")
        assert got[0].dtype == torch.bfloat16
        assert rel_err(got[0].float(), want[0]) < 1e-2 and rel_err(got[1].float(), want[1]) < 1e-2


# ----------------------------------------------------------------------- RNG parity
@pytest.mark.parametrize("n", [1, 7, 4096, 303104, 303104 * 4 + 5, 2 * 1000 * 777])
def test_philox_stream_is_torchs(n):
    """The epilogue's generator reproduces torch.randn on the same device bit for bit."""
    dev = torch.device("cuda", torch.cuda.current_device())
    torch.manual_seed(1234)
    torch.randn(3, device=DEV)                       # move the offset off zero
    gen, seed, offset, threads, inc = nv.philox_plan(dev, n)
    want = torch.randn(n, device=DEV)
    assert gen.get_offset() == offset + inc
    got = ops.randn_philox_torch(n, seed, offset, threads, 1.0, dev)
    assert torch.equal(got, want)


@pytest.mark.parametrize("cplx_", [True, False])
def test_fused_noise_equals_reference_draw_on_device(cplx_, math):
    """Same seed -> the fused kernel's output equals the kernel fed with the noise the
    reference would have drawn on this device (cplx.randn_like / torch.randn_like), and the
    generator ends at the same offset."""
    torch.manual_seed(77)
    M, N, K = 200, 136, 64
    layer = (CplxLinearVD(K, N) if cplx_ else LinearVD(K, N)).to(DEV).train()
    with torch.no_grad():
        layer.log_sigma2.uniform_(-6, 1)
    x = cplx.randn(M, K, device=DEV) if cplx_ else torch.randn(M, K, device=DEV)
    gen = torch.cuda.default_generators[torch.cuda.current_device()]
    torch.manual_seed(555)
    with torch.no_grad():
        fused = layer(x)
    off_fused = gen.get_offset()
    torch.manual_seed(555)
    eps = cplx.randn(M, N, device=DEV) if cplx_ else torch.randn(M, N, device=DEV)
    off_ref = gen.get_offset()
    with torch.no_grad():
        inject = layer(x, eps=eps)
    assert off_fused == off_ref
    if cplx_:
        assert torch.equal(fused.real, inject.real) and torch.equal(fused.imag, inject.imag)
    else:
        assert torch.equal(fused, inject)


def test_fast_noise_mode_statistics():
    torch.manual_seed(9)
    M, N, K = 512, 256, 64
    layer = CplxLinearVD(K, N, bias=False).to(DEV).train()
    with torch.no_grad():
        layer.weight.real.zero_(); layer.weight.imag.zero_(); layer.log_sigma2.zero_()
    x = cplx.Cplx(torch.ones(M, K, device=DEV), torch.zeros(M, K, device=DEV))  # s2 = K exactly
    cb.set_noise_mode("fast")
    try:
        with torch.no_grad():
            a, b = layer(x), layer(x)
    finally:
        cb.set_noise_mode("torch")
    assert not torch.equal(a.real, b.real)          # generator advanced
    for plane in (a.real, a.imag):
        z = plane / K ** 0.5
        assert abs(z.mean().item()) < 0.01 and abs(z.var().item() - 0.5) < 0.01
    assert abs((a.real * a.imag).mean().item()) / K < 0.01


def test_eval_mode_returns_mean_and_module_golden():
    g = load_golden("cplx_linear_vd")
    layer = CplxLinearVD(64, 72)
    layer.load_state_dict({"weight.real": g["w_re"], "weight.imag": g["w_im"], "bias.real": g["b_re"],
                           "bias.imag": g["b_im"], "log_sigma2": g["log_sigma2"]})
    layer = layer.to(DEV)
    z = cplx.Cplx(g["x_re"].to(DEV), g["x_im"].to(DEV))
    with torch.no_grad():
        mu = layer.eval()(z)
        y = layer.train()(z, eps=cplx.Cplx(g["eps_re"].to(DEV), g["eps_im"].to(DEV)))
    assert rel_err(mu.real, g["mu_re"]) < 1e-3 and rel_err(mu.imag, g["mu_im"]) < 1e-3
    assert rel_err(y.real, g["y_re"]) < 1e-3 and rel_err(y.imag, g["y_im"]) < 1e-3


def test_full_size_rows_sampled_against_oracle():
    """BASELINE headline size (B = d = 4096): 48 sampled rows of the fused tensor-core
    output against the float64 oracle; eval-mode linearity  f(a x) = a f(x)."""
    torch.manual_seed(21)
    B = D = 4096
    layer = CplxLinearVD(D, D).to(DEV).train()
    with torch.no_grad():
        layer.log_sigma2.uniform_(-12, -2)
    x = cplx.randn(B, D, device=DEV)
    eps = cplx.randn(B, D, device=DEV)
    with torch.no_grad():
        y = layer(x, eps=eps)
    rows = torch.randperm(B)[:48]
    c = lambda t: t.detach().cpu().double()
    w, b = layer.weight, layer.bias
    want = orc.cplx_linear_vd(c(x.real[rows]), c(x.imag[rows]), c(w.real), c(w.imag), c(b.real),
                              c(b.imag), c(layer.log_sigma2), c(eps.real[rows]), c(eps.imag[rows]))
    assert rel_err(y.real[rows], want[0]) < 1e-3 and rel_err(y.imag[rows], want[1]) < 1e-3
    lin = layer.eval()
    with torch.no_grad():
        f1 = lin(x)
        f2 = lin(cplx.Cplx(2 * x.real, 2 * x.imag))
    b2 = cplx.Cplx(b.real.detach(), b.imag.detach())
    assert rel_err(f2.real - b2.real, 2 * (f1.real - b2.real)) < 1e-5


@pytest.mark.parametrize("prepass", [True, False])
def test_prepass_and_inkernel_transform_agree(prepass):
    """Workspace variant (|x|^2, exp(log_sigma2) via an elementwise pre-pass + TMA) and the
    single-launch variant (made in the GEMM's smem pipeline) against the float64 oracle."""
    torch.manual_seed(31)
    M, N, K = 300, 260, 520
    x_re, x_im = torch.randn(M, K), torch.randn(M, K)
    w_re, w_im = torch.randn(N, K) / K ** 0.5, torch.randn(N, K) / K ** 0.5
    ls2 = torch.empty(N, K).uniform_(-8, 1)
    eps = (torch.randn(M, N), torch.randn(M, N))
    want = orc.cplx_linear_vd(x_re.double(), x_im.double(), w_re.double(), w_im.double(), None,
                              None, ls2.double(), eps[0].double(), eps[1].double())
    d = lambda t: t.to(DEV)
    cb.set_operand_prepass(prepass)
    try:
        got = ops.cplx_linear_vd(d(x_re), d(x_im), d(w_re), d(w_im), None, None, d(ls2),
                                 eps=(d(eps[0]), d(eps[1])))
        real = ops.real_linear_vd(d(x_re), d(w_re), None, d(ls2), eps=d(eps[0]))
    finally:
        cb.set_operand_prepass(True)
    assert rel_err(got[0], want[0]) < 1e-3 and rel_err(got[1], want[1]) < 1e-3
    assert rel_err(real, orc.real_linear_vd(x_re.double(), w_re.double(), None, ls2.double(),
                                            eps[0].double())) < 1e-3


def test_dense_to_vd_to_masked_pipeline_on_gpu():
    """compute_ard_masks -> binarize_masks -> deploy_masks; the masked layer's forward is the
    accelerated kernel on weight * mask (reference: tests/test_relevance.py:206-242)."""
    from cplxmodule_b200.nn.masked import CplxLinearMasked, binarize_masks, deploy_masks
    from cplxmodule_b200.nn.relevance import compute_ard_masks
    from cplxmodule_b200.nn.utils import sparsity
    torch.manual_seed(17)
    vd = CplxLinearVD(64, 48).to(DEV)
    with torch.no_grad():
        vd.log_sigma2.uniform_(-12, 4)
    masks = compute_ard_masks(vd, threshold=0.0)
    assert set(masks) == {"mask"} and 0 < masks["mask"].mean().item() < 1
    state, masks = binarize_masks(vd.state_dict(), masks)
    masked = CplxLinearMasked(64, 48).to(DEV)
    masked.load_state_dict(state, strict=False)
    deploy_masks(masked, state_dict=masks)
    z = cplx.randn(10, 64, device=DEV)
    out = masked(z)
    c = lambda t: t.detach().cpu()
    mk = c(masks["mask"])
    want = orc.cplx_linear(c(z.real), c(z.imag), c(vd.weight.real) * mk, c(vd.weight.imag) * mk,
                           c(vd.bias.real), c(vd.bias.imag))
    assert rel_err(out.real, want[0]) < 1e-3 and rel_err(out.imag, want[1]) < 1e-3
    frac = sparsity(masked)
    assert abs(frac - (1 - mk.mean().item()) * (2 * 64 * 48) / (2 * 64 * 48 + 2 * 48)) < 1e-6
    assert abs(sparsity(vd, threshold=0.0) - frac) < 1e-6
