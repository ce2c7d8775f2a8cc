"""real-plane convolutions and grouped convolutions on the tcgen05 implicit-GEMM
kernel (`conv_tc_kernel<T, VD, real>`), ONE launch per layer call whatever `groups` is.
Reference: F.conv{1,2}d as called by ConvNdGaussianMixin._forward_impl (nn/relevance/real/base.py:
149-163) and convnd_naive (cplxmodule/cplx.py:717-726); oracle = the same torch calls in float64."""
import pytest
import torch
import torch.nn.functional as F

import cplxmodule_b200 as cb
from cplxmodule_b200 import conv_ops, cplx, ops
from cplxmodule_b200.nn import CplxConv2d
from cplxmodule_b200.nn import relevance as rel
from oracle import cplx_oracle as orc
from tests.conftest import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"
c64 = lambda t: t.detach().double().cpu()


@pytest.fixture
def tensor_only(monkeypatch):
    """math mode 'tensor' (the C ABI refuses anything but the tcgen05 path) and no per-group loop"""
    def boom(*a, **k):
        raise AssertionError("grouped convolution fell back to one call per group")
    monkeypatch.setattr(conv_ops, "_conv2d_raw_per_group", boom)
    ops.set_math_mode("tensor")
    yield
    ops.set_math_mode("auto")


GEOMS = [
    # (B, C, H, W, O, k, stride, padding, dilation, groups)
    (2, 16, 20, 140, 200, 3, 1, 0, 1, 1),        # two 128-channel n-blocks, ragged w-tile
    (3, 5, 9, 11, 4, 3, 1, 1, 1, 1),             # channel padding, tiny O
    (2, 48, 17, 19, 72, (3, 5), 1, (1, 2), 1, 2),   # groups: Cg = 24 is NOT a whole k-block
    (2, 64, 12, 33, 64, 3, 2, 1, 1, 4),          # stride 2, Cg = Og = 16
    (1, 12, 14, 9, 18, 2, 1, 0, (2, 3), 3),      # dilation, Cg = 4 (padded to 8), Og = 6
    (2, 8, 1, 40, 8, (1, 7), (1, 3), (0, 3), 1, 8),  # depthwise conv1d-like
]


@pytest.mark.parametrize("geom", GEOMS)
@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_real_conv_on_tensor_cores(geom, dt, tensor_only):
    """plain real cross-correlation (the mean conv of every real layer, the variance conv of every
    conv backward)"""
    B, C, H, W, O, k, stride, padding, dilation, groups = geom
    torch.manual_seed(sum(geom[:5]))
    k2 = (k, k) if isinstance(k, int) else k
    x = torch.randn(B, C, H, W, device=DEV).to(dt)
    w = (torch.randn(O, C // groups, *k2, device=DEV) / (C // groups * k2[0] * k2[1]) ** 0.5).to(dt)
    b = torch.randn(O, device=DEV).to(dt)
    out = conv_ops.real_convnd(2, x, w, b, stride, padding, dilation, groups)
    want = F.conv2d(c64(x), c64(w), c64(b), stride, padding, dilation, groups)
    assert out.shape == want.shape and out.dtype == dt
    assert rel_err(out, want) < (1e-3 if dt == torch.float32 else 1e-2)


@pytest.mark.parametrize("geom", GEOMS)
def test_real_conv_vd_on_tensor_cores(geom, tensor_only):
    B, C, H, W, O, k, stride, padding, dilation, groups = geom
    torch.manual_seed(sum(geom[:5]) + 1)
    m = rel.Conv2dVD(C, O, k, stride=stride, padding=padding, dilation=dilation, groups=groups).to(DEV).train()
    with torch.no_grad():
        m.log_sigma2.uniform_(-10, 1)
    x = torch.randn(B, C, H, W, device=DEV)
    with torch.no_grad():
        mu = m.eval()(x)
        eps = torch.randn_like(mu)
        out = m.train()(x, eps=eps)
        torch.manual_seed(77)
        fused = m(x)                                  # in-kernel Philox, torch layout, ONE launch
        torch.manual_seed(77)
        e2 = torch.randn_like(fused)
        inject = m(x, eps=e2)
    want = orc.real_conv2d_vd(c64(x), c64(m.weight), c64(m.bias), c64(m.log_sigma2), c64(eps), m.stride,
                              m.padding, m.dilation, groups)
    assert rel_err(mu, orc.real_conv2d_vd(c64(x), c64(m.weight), c64(m.bias), None, None, m.stride,
                                          m.padding, m.dilation, groups)) < 1e-3
    assert rel_err(out, want) < 1e-3
    # the fused draw is the stream torch.randn_like(out) produces from the same generator state
    assert torch.equal(fused, inject)


@pytest.mark.parametrize("geom", GEOMS[2:])
@pytest.mark.parametrize("vd", [False, True])
def test_grouped_complex_conv_one_launch(geom, vd, tensor_only):
    B, C, H, W, O, k, stride, padding, dilation, groups = geom
    torch.manual_seed(sum(geom[:5]) + 2)
    cls = rel.CplxConv2dVD if vd else CplxConv2d
    m = cls(C, O, k, stride=stride, padding=padding, dilation=dilation, groups=groups).to(DEV).train()
    z = cplx.randn(B, C, H, W, device=DEV)
    args = [c64(z.real), c64(z.imag), c64(m.weight.real), c64(m.weight.imag), c64(m.bias.real), c64(m.bias.imag)]
    if vd:
        with torch.no_grad():
            m.log_sigma2.uniform_(-8, 0)
            torch.manual_seed(5)
            fused = m(z)
            torch.manual_seed(5)
            eps = cplx.randn(*fused.shape, device=DEV)       # the reference's ONE randn(2, ...) / sqrt(2)
            out = m(z, eps=eps)
        want = orc.cplx_conv2d_vd(*args, c64(m.log_sigma2), c64(eps.real), c64(eps.imag), m.stride,
                                  m.padding, m.dilation, groups)
        assert torch.equal(fused.real, out.real) and torch.equal(fused.imag, out.imag)
    else:
        with torch.no_grad():
            out = m(z)
        want = orc.cplx_conv2d_grouped(*args, m.stride, m.padding, m.dilation, groups)
    assert out.shape == want[0].shape
    assert rel_err(out.real, want[0]) < 1e-3 and rel_err(out.imag, want[1]) < 1e-3


def test_group_boundary_channels_do_not_leak(tensor_only):
    """Cg = 24 (fp32 k-block = 32 channels): the k-block of group 0 also LOADS the first 8 channels
    of group 1; their weight rows are zero in the prepared planes.  Make those channels huge."""
    torch.manual_seed(3)
    x = torch.randn(2, 48, 10, 12, device=DEV)
    x[:, 24:32] *= 1e6
    w = torch.randn(16, 24, 3, 3, device=DEV) / 15
    out = conv_ops.real_convnd(2, x, w, None, 1, 1, 1, 2)
    want = F.conv2d(c64(x), c64(w), None, 1, 1, 1, 2)
    assert rel_err(out[:, :8], want[:, :8]) < 1e-3        # group 0 never sees the 1e6 channels
    assert rel_err(out[:, 8:], want[:, 8:]) < 1e-3


@pytest.mark.parametrize("cplx_", [False, True])
def test_grouped_conv_vd_gradients(cplx_):
    """grouped variational conv, forward in one launch, backward = grouped dgrad in one launch +
    one wgrad GEMM per group: gradients vs float64 autograd over the oracle"""
    torch.manual_seed(11)
    G = 2
    if cplx_:
        m = rel.CplxConv2dVD(8, 6, 3, padding=1, groups=G).to(DEV).train()
        x = cplx.randn(2, 8, 7, 9, device=DEV)
        x = cplx.Cplx(x.real.requires_grad_(), x.imag.requires_grad_())
        eps = cplx.randn(2, 6, 7, 9, device=DEV)
        params = [m.weight.real, m.weight.imag, m.bias.real, m.bias.imag, m.log_sigma2]
        ins = [x.real, x.imag]
    else:
        m = rel.Conv2dVD(8, 6, 3, padding=1, groups=G).to(DEV).train()
        x = torch.randn(2, 8, 7, 9, device=DEV, requires_grad=True)
        eps = torch.randn(2, 6, 7, 9, device=DEV)
        params = [m.weight, m.bias, m.log_sigma2]
        ins = [x]
    with torch.no_grad():
        m.log_sigma2.uniform_(-6, -1)
    out = m(x, eps=eps)
    gy = [torch.randn_like(t) for t in ((out.real, out.imag) if cplx_ else (out,))]
    loss = sum((a * b).sum() for a, b in zip((out.real, out.imag) if cplx_ else (out,), gy))
    got = torch.autograd.grad(loss, ins + params)

    d = lambda t: t.detach().double().cpu().requires_grad_()
    ins64, par64 = [d(t) for t in ins], [d(t) for t in params]
    if cplx_:
        o = orc.cplx_conv2d_vd(ins64[0], ins64[1], par64[0], par64[1], par64[2], par64[3], par64[4],
                               c64(eps.real), c64(eps.imag), 1, 1, 1, G)
    else:
        o = (orc.real_conv2d_vd(ins64[0], par64[0], par64[1], par64[2], c64(eps), 1, 1, 1, G),)
    loss64 = sum((a * c64(b)).sum() for a, b in zip(o, gy))
    want = torch.autograd.grad(loss64, ins64 + par64)
    for a, b in zip(got, want):
        assert a.shape == b.shape and rel_err(a, b) < 2e-3


def test_simt_mode_still_loops_over_groups():
    """exact-fp32 CUDA-core kernel: one call per group, same numbers to 2e-5"""
    torch.manual_seed(4)
    ops.set_math_mode("simt")
    try:
        m = rel.Conv2dVD(6, 8, 3, padding=1, groups=2).to(DEV).train()
        x = torch.randn(2, 6, 8, 8, device=DEV)
        eps = torch.randn(2, 8, 8, 8, device=DEV)
        with torch.no_grad():
            out = m(x, eps=eps)
        want = orc.real_conv2d_vd(c64(x), c64(m.weight), c64(m.bias), c64(m.log_sigma2), c64(eps), 1, 1, 1, 2)
        assert rel_err(out, want) < 2e-5
    finally:
        ops.set_math_mode("auto")


def test_grouped_conv1d_and_circular_padding(tensor_only):
    """conv1d is the H == 1 case of the same launch; circular padding pads first (cplx.py:701-714,
    784-786), then the grouped kernel runs on the padded planes"""
    torch.manual_seed(9)
    from cplxmodule_b200.nn import CplxConv1d
    m = CplxConv1d(12, 8, 5, stride=2, padding=3, dilation=2, groups=4).to(DEV)
    z = cplx.randn(3, 12, 50, device=DEV)
    with torch.no_grad():
        out = m(z)
    conv = lambda a, w: F.conv1d(c64(a), c64(w), None, 2, 3, 2, 4)
    re = conv(z.real, m.weight.real) - conv(z.imag, m.weight.imag) + c64(m.bias.real)[None, :, None]
    im = conv(z.real, m.weight.imag) + conv(z.imag, m.weight.real) + c64(m.bias.imag)[None, :, None]
    assert out.shape == re.shape
    assert rel_err(out.real, re) < 1e-3 and rel_err(out.imag, im) < 1e-3
    m2 = CplxConv2d(8, 8, 3, padding=2, padding_mode="circular", groups=2).to(DEV)
    z2 = cplx.randn(2, 8, 9, 10, device=DEV)
    with torch.no_grad():
        out2 = m2(z2)
    pad = lambda a: F.pad(c64(a), (1, 1, 1, 1), mode="circular")
    want = orc.cplx_conv2d_grouped(pad(z2.real), pad(z2.imag), c64(m2.weight.real), c64(m2.weight.imag),
                                   c64(m2.bias.real), c64(m2.bias.imag), 1, 0, 1, 2)
    assert rel_err(out2.real, want[0]) < 1e-3 and rel_err(out2.imag, want[1]) < 1e-3


def test_real_masked_conv_with_groups(tensor_only):
    """Conv2dMasked (nn/masked/real.py) with groups: weight * mask on the grouped tcgen05 path"""
    from cplxmodule_b200.nn import masked
    torch.manual_seed(10)
    conv = masked.Conv2dMasked(8, 12, 3, padding=1, groups=2).to(DEV)
    conv.mask = (torch.rand_like(conv.weight) < 0.6).float()
    x = torch.randn(2, 8, 11, 13, device=DEV)
    with torch.no_grad():
        out = conv(x)
    want = F.conv2d(c64(x), c64(conv.weight) * c64(conv.mask), c64(conv.bias), 1, 1, 1, 2)
    assert rel_err(out, want) < 1e-3


@pytest.mark.parametrize("mode", ["composed", "fused"])
@pytest.mark.parametrize("cplx_", [True, False])
@pytest.mark.parametrize("dt", [torch.float32, torch.bfloat16])
def test_conv_vd_modes_agree_with_oracle_and_each_other(mode, cplx_, dt):
    """ops.set_conv_vd_mode: mean conv + variance conv + in-place noise launch ('composed') and the
    single fused kernel draw the SAME torch-exact noise and meet the same tolerance"""
    torch.manual_seed(21)
    tol = 1e-3 if dt == torch.float32 else 1e-2
    cls = rel.CplxConv2dVD if cplx_ else rel.Conv2dVD
    m = cls(16, 24, 3, padding=1, stride=(1, 2)).to(DEV).train()
    with torch.no_grad():
        m.log_sigma2.uniform_(-8, 0)
    m = m.to(dt)
    x = (cplx.randn(3, 16, 20, 37, device=DEV) if cplx_ else torch.randn(3, 16, 20, 37, device=DEV)).to(dt)
    planes = (lambda t: (t.real, t.imag)) if cplx_ else (lambda t: (t,))
    cb.set_conv_vd_mode(mode)
    try:
        with torch.no_grad():
            torch.manual_seed(5)
            fused = m(x)
            torch.manual_seed(5)
            eps = (cplx.randn(*fused.shape, device=DEV) if cplx_ else torch.randn(*fused.shape, device=DEV))
            out = m(x, eps=eps.to(dt))
    finally:
        cb.set_conv_vd_mode("auto")
    f32 = lambda t: t.float()
    if dt == torch.float32:                      # same stream: bit-equal to the injected draw
        for a, b in zip(planes(fused), planes(out)):
            assert torch.equal(a, b)
    else:                                        # the injected draw is rounded to bf16 first
        for a, b in zip(planes(fused), planes(out)):
            assert rel_err(f32(a), f32(b)) < 2e-2
    w, b = m.weight, m.bias
    if cplx_:
        want = orc.cplx_conv2d_vd(c64(x.real), c64(x.imag), c64(w.real), c64(w.imag), c64(b.real), c64(b.imag),
                                  c64(m.log_sigma2), c64(eps.real.to(dt)), c64(eps.imag.to(dt)), (1, 2), 1, 1)
    else:
        want = (orc.real_conv2d_vd(c64(x), c64(w), c64(b), c64(m.log_sigma2), c64(eps.to(dt)), (1, 2), 1, 1, 1),)
    for a, r in zip(planes(out), want):
        assert rel_err(f32(a), r) < tol


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_variance_conv_squares_the_input_in_the_prepass(dtype):
    """s2 = conv(|x|^2, exp(log_sigma2)) of the complex variational layers in ONE call: the real-plane
    kernel's transposing pre-pass forms x_re^2 + x_im^2 (cplxk_conv2d_fwd_g with x_im but real weights
    and output) -- against float64 F.conv2d, against the two-step form, and the fall-back (None) for a
    width the 16-byte-load transposers cannot take.  Reference: complex/base.py:100-117."""
    from cplxmodule_b200 import conv_ops
    torch.manual_seed(3)
    B, C, H, W, O = 3, 32, 11, 136, 40
    xr, xi = torch.randn(B, C, H, W, device=DEV).to(dtype), torch.randn(B, C, H, W, device=DEV).to(dtype)
    E = torch.rand(O, C, 3, 3, device=DEV).to(dtype)
    geom = ((1, 1), (1, 1), (1, 1))
    s2 = conv_ops._variance_conv2d(xr, xi, E, geom, 1)
    assert s2 is not None and s2.shape == (B, O, H, W)
    c = lambda t: t.detach().double().cpu()
    want = F.conv2d(c(xr) ** 2 + c(xi) ** 2, c(E), padding=1)
    tol = 1e-3 if dtype == torch.float32 else 1e-2
    assert rel_err(s2.float(), want) < tol
    q = (xr.float() ** 2 + xi.float() ** 2).to(dtype)
    two_step, _, _ = conv_ops._conv2d_raw(q, None, E, None, None, None, None, None, None, 0, geom, 1)
    assert rel_err(s2.float(), two_step.float().cpu()) < tol
    # groups, stride, a narrow image (several rows per tile)
    xr2, xi2 = xr[:, :, :, :40].contiguous(), xi[:, :, :, :40].contiguous()
    Eg = torch.rand(O, C // 2, 3, 3, device=DEV).to(dtype)
    geom2 = ((2, 1), (0, 1), (1, 1))
    s2g = conv_ops._variance_conv2d(xr2, xi2, Eg, geom2, 2)
    want = F.conv2d(c(xr2) ** 2 + c(xi2) ** 2, c(Eg), stride=(2, 1), padding=(0, 1), groups=2)
    assert s2g is not None and rel_err(s2g.float(), want) < tol
    # W % 4 != 0: not available, the caller forms |x|^2 itself
    assert conv_ops._variance_conv2d(xr[..., :37].contiguous(), xi[..., :37].contiguous(), E, geom, 1) is None


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("O,W", [(8, 140), (40, 140), (64, 200), (64, 30), (24, 30), (72, 140)])
def test_real_conv_pair_kernel_channel_blocks(O, W, dtype, tensor_only):
    """Ungrouped real planes on the CTA-pair kernel (128-channel n-blocks: layers with fewer output
    channels multiply zero-padded weight rows and store only their own channels; 72 -> one block, the
    second half mostly padding): wide images (one output row per tile: row mode) and narrow ones
    (several rows per tile), bias, vs float64 F.conv2d."""
    torch.manual_seed(O + W)
    B, C, H = 3, 24, 9
    x = torch.randn(B, C, H, W, device=DEV).to(dtype)
    w = (torch.randn(O, C, 3, 3, device=DEV) / 8).to(dtype)
    b = torch.randn(O, device=DEV).to(dtype)
    y = conv_ops.real_convnd(2, x, w, b, padding=1)
    want = F.conv2d(c64(x), c64(w), c64(b), padding=1)
    assert y.shape == want.shape
    assert rel_err(y.float(), want) < (1e-3 if dtype == torch.float32 else 1e-2)
