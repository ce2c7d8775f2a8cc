"""Complex conv1d/conv2d (+ variational forward) on the GPU vs the oracle / golden fixtures."""
import pytest
import torch
import torch.nn.functional as F

import cplxmodule_b200 as cb
from cplxmodule_b200 import cplx, ops
from cplxmodule_b200.nn import CplxConv1d, CplxConv2d
from cplxmodule_b200.nn.relevance import CplxConv2dVD, penalties
from oracle import cplx_oracle as orc
from tests.conftest import load_golden, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = {"simt": 2e-5, "tensor": 1e-3}


@pytest.fixture(params=["simt", "tensor"])
def math(request):
    from cplxmodule_b200 import ops
    ops.set_math_mode(request.param)
    yield request.param
    ops.set_math_mode("auto")


def _mod(cls, g, suffix="", **kw):
    w = g[f"w{suffix}_re"]
    m = cls(w.shape[1], w.shape[0], tuple(w.shape[2:]), **kw)
    sd = {"weight.real": w, "weight.imag": g[f"w{suffix}_im"], "bias.real": g[f"b{suffix}_re"],
          "bias.imag": g[f"b{suffix}_im"]}
    if "log_sigma2" in g and suffix == "" and hasattr(m, "log_sigma2"):
        sd["log_sigma2"] = g["log_sigma2"]
    m.load_state_dict(sd)
    return m.to(DEV)


def test_golden_conv2d(math):
    g = load_golden("cplx_conv2d")
    z = cplx.Cplx(g["x_re"].to(DEV), g["x_im"].to(DEV))
    out = _mod(CplxConv2d, g)(z)
    assert out.shape == g["y_re"].shape
    assert rel_err(out.real, g["y_re"]) < TOL[math] and rel_err(out.imag, g["y_im"]) < TOL[math]
    out = _mod(CplxConv2d, g, "2", stride=(2, 1), padding=(1, 2), dilation=(1, 2))(z)
    assert out.shape == g["y2_re"].shape
    assert rel_err(out.real, g["y2_re"]) < TOL[math] and rel_err(out.imag, g["y2_im"]) < TOL[math]


def test_golden_conv2d_vd(math):
    g = load_golden("cplx_conv2d_vd")
    layer = _mod(CplxConv2dVD, g, padding=1).train()
    z = cplx.Cplx(g["x_re"].to(DEV), g["x_im"].to(DEV))
    with torch.no_grad():
        out = layer(z, eps=cplx.Cplx(g["eps_re"].to(DEV), g["eps_im"].to(DEV)))
        kl = sum(penalties(layer))
    assert rel_err(out.real, g["y_re"]) < TOL[math] and rel_err(out.imag, g["y_im"]) < TOL[math]
    assert abs(kl.item() - g["penalty_sum"].item()) / g["penalty_sum"].item() < 1e-3


@pytest.mark.parametrize("shape", [
    # (B, C, H, W, O, k, stride, padding, dilation)
    (3, 64, 20, 140, 64, 3, 1, 0, 1),       # Wt = 128 with a ragged second w-tile
    (2, 40, 17, 19, 72, (3, 5), 1, (1, 2), 1),   # Cp/Op padding, Wt = 32, two n-blocks
    (2, 16, 30, 33, 8, 3, 2, 1, 1),         # stride 2 through the tensor map's element strides
    (1, 24, 12, 9, 5, 2, 1, 0, (2, 3)),     # dilation, Wt = 16
    (2, 8, 6, 40, 4, (1, 7), (1, 3), (0, 3), 1),  # conv1d-like row kernel, stride 3
])
@pytest.mark.parametrize("vd", [False, True])
def test_tensor_core_conv_shapes(shape, vd):
    """tcgen05 implicit GEMM (channels-last pre-pass + 4-d TMA boxes) vs the float64 oracle."""
    from cplxmodule_b200 import ops
    B, C, H, W, O, k, stride, padding, dilation = shape
    torch.manual_seed(hash(shape) % 1000)
    cls = CplxConv2dVD if vd else CplxConv2d
    m = cls(C, O, k, stride=stride, padding=padding, dilation=dilation).to(DEV).train()
    z = cplx.randn(B, C, H, W, device=DEV)
    c = lambda t: t.detach().cpu().double()
    args = [c(z.real), c(z.imag), c(m.weight.real), c(m.weight.imag), c(m.bias.real), c(m.bias.imag)]
    if vd:
        with torch.no_grad():
            m.log_sigma2.uniform_(-8, 0)
        want_mu = orc.cplx_conv2d(*args, m.stride, m.padding, m.dilation)
        eps = cplx.randn(*want_mu[0].shape, device=DEV)
        want = orc.cplx_conv2d_vd(*args, c(m.log_sigma2), c(eps.real), c(eps.imag), m.stride,
                                  m.padding, m.dilation)
    else:
        eps = None
        want = orc.cplx_conv2d(*args, m.stride, m.padding, m.dilation)
    ops.set_math_mode("tensor")
    try:
        with torch.no_grad():
            out = m(z, eps=eps) if vd else m(z)
    finally:
        ops.set_math_mode("auto")
    assert out.shape == want[0].shape
    assert rel_err(out.real, want[0]) < 1e-3 and rel_err(out.imag, want[1]) < 1e-3


def test_full_size_conv_config4_one_image():
    """BASELINE config 4 (256 x 64 x 128 x 128, 3x3, 64 -> 64): images 0 and 255 of the
    tensor-core result against the oracle on CPU; bf16 planes within 1e-2."""
    torch.manual_seed(12)
    m = CplxConv2d(64, 64, 3).to(DEV)
    z = cplx.randn(256, 64, 128, 128, device=DEV)
    with torch.no_grad():
        out = m(z)
    assert out.shape == (256, 64, 126, 126)
    c = lambda t: t.detach().cpu()
    for b in (0, 255):
        want = orc.cplx_conv2d(c(z.real[b:b + 1]), c(z.imag[b:b + 1]), c(m.weight.real),
                               c(m.weight.imag), c(m.bias.real), c(m.bias.imag))
        assert rel_err(out.real[b:b + 1], want[0]) < 1e-3
        assert rel_err(out.imag[b:b + 1], want[1]) < 1e-3
    m16, z16 = m.bfloat16(), z.to(torch.bfloat16)
    with torch.no_grad():
        out16 = m16(z16)
    want = orc.cplx_conv2d(c(z16.real[:1]).float(), c(z16.imag[:1]).float(), c(m16.weight.real).float(),
                           c(m16.weight.imag).float(), c(m16.bias.real).float(), c(m16.bias.imag).float())
    assert rel_err(out16.real[:1].float(), want[0]) < 1e-2


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("vd", [False, True])
def test_channels_last_conv_matches_nchw_and_oracle(vd, dtype):
    """torch.channels_last activations are read in place by the implicit-GEMM kernel (no
    transposing pre-pass); the output keeps the memory format and equals the NCHW result."""
    torch.manual_seed(31)
    B, C, H, W, O = 3, 32, 18, 37, 24
    cls = CplxConv2dVD if vd else CplxConv2d
    m = cls(C, O, 3, padding=1, stride=(1, 2)).to(DEV).train()
    if vd:
        with torch.no_grad():
            m.log_sigma2.uniform_(-8, 0)
    m = m.to(dtype)
    z = cplx.randn(B, C, H, W, device=DEV).to(dtype)
    zcl = cplx.Cplx(z.real.contiguous(memory_format=torch.channels_last),
                    z.imag.contiguous(memory_format=torch.channels_last))
    eps = cplx.randn(B, O, 18, 19, device=DEV).to(dtype) if vd else None
    with torch.no_grad():
        ref = m(z, eps=eps) if vd else m(z)
        out = m(zcl, eps=eps) if vd else m(zcl)
    assert out.shape == ref.shape == (B, O, 18, 19)
    assert out.real.is_contiguous(memory_format=torch.channels_last)
    assert out.imag.is_contiguous(memory_format=torch.channels_last)
    tol = 1e-3 if dtype == torch.float32 else 1e-2
    assert rel_err(out.real.float(), ref.real.float().cpu()) < tol
    assert rel_err(out.imag.float(), ref.imag.float().cpu()) < tol
    c = lambda t: t.detach().cpu().double()
    args = [c(z.real), c(z.imag), c(m.weight.real), c(m.weight.imag), c(m.bias.real), c(m.bias.imag)]
    if vd:
        want = orc.cplx_conv2d_vd(*args, c(m.log_sigma2), c(eps.real), c(eps.imag), m.stride,
                                  m.padding, m.dilation)
    else:
        want = orc.cplx_conv2d(*args, m.stride, m.padding, m.dilation)
    assert rel_err(out.real.float(), want[0]) < tol and rel_err(out.imag.float(), want[1]) < tol


def test_channels_last_conv_fused_noise_and_layout_fallback():
    """Fused torch-exact noise is indexed in logical NCHW order (cplx.randn draws a contiguous
    tensor, cplx.py:544-550) whatever the activation layout; channel counts the in-place path
    cannot take (C % 8 != 0) go through the NCHW kernels."""
    torch.manual_seed(6)
    layer = CplxConv2dVD(16, 8, 3, padding=1).to(DEV).train()
    with torch.no_grad():
        layer.log_sigma2.uniform_(-6, 1)
    z = cplx.randn(2, 16, 9, 12, device=DEV)
    zcl = cplx.Cplx(z.real.contiguous(memory_format=torch.channels_last),
                    z.imag.contiguous(memory_format=torch.channels_last))
    torch.manual_seed(42)
    with torch.no_grad():
        fused = layer(zcl)
    torch.manual_seed(42)
    eps = cplx.randn(2, 8, 9, 12, device=DEV)
    with torch.no_grad():
        inject = layer(zcl, eps=eps)
        nchw = layer(z, eps=eps)
    assert fused.real.is_contiguous(memory_format=torch.channels_last)
    assert torch.equal(fused.real, inject.real) and torch.equal(fused.imag, inject.imag)
    assert rel_err(inject.real, nchw.real.cpu()) < 1e-3
    m = CplxConv2d(6, 4, 3).to(DEV)
    z6 = cplx.randn(2, 6, 8, 8, device=DEV)
    z6cl = cplx.Cplx(z6.real.contiguous(memory_format=torch.channels_last),
                     z6.imag.contiguous(memory_format=torch.channels_last))
    with torch.no_grad():
        assert torch.equal(m(z6cl).real, m(z6).real)


def test_conv_fast_noise_same_stream_on_both_kernels_and_unit_variance():
    """'fast' noise for complex conv: one Philox call + one Box-Muller per complex output element,
    indexed by the logical NCHW element -> the CUDA-core and tensor-core kernels draw the same
    stream, and the standardised draw is CN(0, 1)."""
    from cplxmodule_b200 import ops
    torch.manual_seed(3)
    layer = CplxConv2dVD(16, 24, 3, padding=1, bias=False).to(DEV).train()
    with torch.no_grad():
        layer.weight.real.zero_(); layer.weight.imag.zero_(); layer.log_sigma2.zero_()
    z = cplx.Cplx(torch.ones(4, 16, 20, 33, device=DEV), torch.zeros(4, 16, 20, 33, device=DEV))
    cb.set_noise_mode("fast")
    try:
        outs = {}
        for mode in ("simt", "tensor"):
            ops.set_math_mode(mode)
            torch.manual_seed(77)
            with torch.no_grad():
                outs[mode] = layer(z)
        ops.set_math_mode("auto")
        with torch.no_grad():
            again = layer(z)
    finally:
        ops.set_math_mode("auto")
        cb.set_noise_mode("torch")
    a, b = outs["simt"], outs["tensor"]
    assert rel_err(b.real, a.real.cpu()) < 1e-3 and rel_err(b.imag, a.imag.cpu()) < 1e-3
    assert not torch.equal(again.real, b.real)      # generator advanced
    # interior pixels see all 9 taps x 16 channels of |x|^2 = 1, sigma2 = 1  ->  s2 = 144
    zr, zi = b.real[:, :, 1:-1, 1:-1] / 12.0, b.imag[:, :, 1:-1, 1:-1] / 12.0
    for plane in (zr, zi):
        assert abs(plane.mean().item()) < 0.01 and abs(plane.var().item() - 0.5) < 0.01
    assert abs((zr * zi).mean().item()) < 0.01


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
    assert rel_err(out.real, re) < 1e-3 and rel_err(out.imag, im) < 1e-3
    # groups = 2 (convnd_naive, cplx.py:717-726)
    m = CplxConv2d(4, 6, 3, groups=2, bias=False).to(DEV)
    z = cplx.randn(2, 4, 7, 7, device=DEV)
    out = m(z)
    f = lambda a, w: F.conv2d(c(a), c(w), None, 1, 0, 1, 2)
    re = f(z.real, m.weight.real) - f(z.imag, m.weight.imag)
    im = f(z.real, m.weight.imag) + f(z.imag, m.weight.real)
    assert rel_err(out.real, re) < 1e-3 and rel_err(out.imag, im) < 1e-3
    # circular padding (cplx.py:701-714, 784-786)
    m = CplxConv2d(2, 3, 3, padding=2, padding_mode="circular").to(DEV)
    z = cplx.randn(1, 2, 6, 5, device=DEV)
    out = m(z)
    pad = lambda a: F.pad(c(a), (1, 1, 1, 1), mode="circular")
    want = orc.cplx_conv2d(pad(z.real), pad(z.imag), c(m.weight.real), c(m.weight.imag),
                           c(m.bias.real), c(m.bias.imag))
    assert rel_err(out.real, want[0]) < 1e-3 and rel_err(out.imag, want[1]) < 1e-3
    with pytest.raises(ValueError):
        cplx.conv2d(z, m.weight, None, padding_mode="reflect")


# ------------------------------------------------------------------ real-valued VD / ARD conv layers
def test_golden_real_conv2d_vd_and_conv1d_ard(math):
    """nn/relevance/real/base.py:149-163 through the CUDA conv kernels (exact-fp32 CUDA-core kernel
    and the tcgen05 real-plane kernel): grouped / strided / dilated Conv2dVD and a Conv1dARD,
    training forward with the reference's captured noise, eval forward, log_alpha, penalties and
    the relevance mask"""
    from cplxmodule_b200.nn.relevance import Conv1dARD, Conv2dVD
    g = load_golden("conv2d_vd")
    m = Conv2dVD(6, 8, (3, 2), stride=(1, 2), padding=(1, 0), dilation=(2, 1), groups=2)
    m.load_state_dict({"weight": g["w"], "bias": g["b"], "log_sigma2": g["log_sigma2"]})
    m = m.to(DEV).train()
    x = g["x"].to(DEV)
    with torch.no_grad():
        out = m(x, eps=g["eps"].to(DEV))
        kl = sum(penalties(m))
        la = m.log_alpha
        assert rel_err(out, g["y"]) < TOL[math]
        assert rel_err(m.eval()(x), g["mu"]) < TOL[math]
    assert rel_err(la, g["log_alpha"]) < 1e-5
    assert abs(kl.item() - g["penalty_sum"].item()) / g["penalty_sum"].item() < 1e-4
    assert rel_err(m.penalty, g["penalty"]) < 1e-4

    g = load_golden("conv1d_ard")
    m = Conv1dARD(5, 7, 4, stride=2, padding=3)
    m.load_state_dict({"weight": g["w"], "bias": g["b"], "log_sigma2": g["log_sigma2"]})
    m = m.to(DEV).train()
    with torch.no_grad():
        out = m(g["x"].to(DEV), eps=g["eps"].to(DEV))
        assert out.shape == g["y"].shape and rel_err(out, g["y"]) < TOL[math]
        kl = sum(penalties(m))
        assert torch.equal(m.relevance(threshold=3.0).cpu(), g["relevance"])
    assert abs(kl.item() - g["penalty_sum"].item()) / g["penalty_sum"].item() < 1e-4


@pytest.mark.parametrize("B,C,H,W,O,k,stride,padding,dilation", [
    (2, 3, 9, 11, 4, 3, 1, 0, 1), (1, 8, 17, 5, 5, (1, 3), (2, 1), (0, 2), 1), (3, 4, 12, 12, 6, 2, 2, 1, 2)])
def test_real_conv2d_vd_shapes_vs_oracle(B, C, H, W, O, k, stride, padding, dilation, math):
    from cplxmodule_b200.nn.relevance import Conv2dVD
    torch.manual_seed(B * 100 + C)
    m = Conv2dVD(C, O, k, stride=stride, padding=padding, dilation=dilation).to(DEV).train()
    with torch.no_grad():
        m.log_sigma2.uniform_(-10, 1)
    x = torch.randn(B, C, H, W, device=DEV)
    c = lambda t: t.detach().double().cpu()
    with torch.no_grad():
        mu = m.eval()(x)
        eps = torch.randn_like(mu)
        out = m.train()(x, eps=eps)
        fused = m(x)                                    # in-kernel Philox, torch layout
    want = orc.real_conv2d_vd(c(x), c(m.weight), c(m.bias), c(m.log_sigma2), c(eps), m.stride,
                              m.padding, m.dilation, 1)
    assert rel_err(out, want) < TOL[math]
    assert fused.shape == out.shape and torch.isfinite(fused).all()
    # the fused draw is the stream torch.randn_like(out) would produce from the same generator state
    torch.manual_seed(77)
    a = m(x)
    torch.manual_seed(77)
    e2 = torch.randn_like(a)
    assert rel_err(a, m(x, eps=e2)) < 1e-6


def test_real_conv_vd_rejects_nonzero_padding_mode_and_cpu():
    from cplxmodule_b200.nn.relevance import Conv2dVD
    with pytest.raises(ValueError, match="zeros"):
        Conv2dVD(3, 3, 3, padding=1, padding_mode="circular")
    m = Conv2dVD(3, 3, 3)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.randn(1, 3, 5, 5))


def test_fp32_nchw_conv_on_scaled_fp16_operands():
    """fp32 NCHW planes run as per-image / per-output-channel scaled fp16 (conv_tc_pair_kernel
    <float, half operands>): images and output channels of very different magnitude, zero image,
    vs the float64 oracle, and against the tf32 path on the same inputs"""
    import os
    torch.manual_seed(21)
    B, C, H, W, O = 9, 24, 20, 36, 40
    m = CplxConv2d(C, O, 3, padding=1, bias=False).to(DEV)   # no bias: errors are judged per (image, channel) scale
    z_re, z_im = torch.randn(B, C, H, W, device=DEV), torch.randn(B, C, H, W, device=DEV)
    # 1.0, 7e2, 0.06: copied unscaled by the optimistic pre-pass (amax in [2^-2, 2^15)); 1e-6, 3e-3, 1e4,
    # 1.6e4 (overflows fp16 unscaled), 0.04 (amax just below 2^-2): re-converted by the fix-up launch
    img_scale = torch.tensor([1e-6, 1.0, 1e4, 0.0, 3e-3, 7e2, 1.6e4, 0.06, 0.04], device=DEV).view(B, 1, 1, 1)
    z_re, z_im = z_re * img_scale, z_im * img_scale
    with torch.no_grad():
        ch_scale = torch.logspace(-4, 3, O, device=DEV).view(O, 1, 1, 1)
        m.weight.real.mul_(ch_scale); m.weight.imag.mul_(ch_scale)
        out = m(cplx.Cplx(z_re, z_im))
        ops.set_math_mode("tf32")
        try:
            out32 = m(cplx.Cplx(z_re, z_im))
        finally:
            ops.set_math_mode("auto")
    c = lambda t: t.detach().double().cpu()
    want = orc.cplx_conv2d(c(z_re), c(z_im), c(m.weight.real), c(m.weight.imag), None, None, 1, 1, 1)
    for got, got32, ref in ((out.real, out32.real, want[0]), (out.imag, out32.imag, want[1])):
        # relative to the scale of each (image, output channel) pair
        scale = (img_scale.double().cpu().clamp_min(1e-30) * ch_scale.double().cpu().view(1, O, 1, 1))
        err = ((c(got) - ref).abs() / scale).amax(dim=(2, 3))
        mag = (ref.abs() / scale).amax(dim=(2, 3)).clamp_min(1e-30)
        keep = img_scale.view(B).cpu() > 0
        assert float((err / mag)[keep].max()) < 2e-3
        assert torch.isfinite(got).all()
        assert rel_err(got, ref) < TOL["tensor"] and rel_err(got32, ref) < TOL["tensor"]


@pytest.mark.parametrize("shape", [
    # (B, C, H, W, O, k, stride, padding, dilation): output rows wider than 64 pixels, unit stride along W
    (2, 64, 9, 200, 64, 3, 1, 1, 1),                      # padding: the row box starts at iw = -1; two w-tiles
    (2, 24, 9, 150, 72, (3, 2), (2, 1), (1, 0), (1, 3)),  # kw = 2, dilation 3 (halo 3), vertical stride 2, two n-blocks
    (1, 16, 5, 300, 8, (2, 3), 1, (0, 4), (1, 4)),        # kw = 3, dilation 4 (halo 8 = the maximum), three w-tiles
    (3, 32, 4, 131, 16, (1, 3), 1, 0, (1, 2)),            # row kernel, halo 4, odd tile out (3 * 4 * 2 tiles ... pairs)
])
@pytest.mark.parametrize("mode", ["f16", "tf32", "bf16", "f32_cl", "bf16_cl"])
def test_row_mode_conv_shapes(shape, mode):
    """Row mode of the CTA-pair kernel (ConvPairCfg<.., kRow>): one load of 128 + (kw-1)*dw pixels
    serves every tap of a kernel row through shifted shared-memory descriptors -- every operand
    format that reaches the kernel (scaled fp16 copies of fp32 NCHW planes, tf32, bf16, channels-last
    in place) against the float64 oracle."""
    B, C, H, W, O, k, stride, padding, dilation = shape
    torch.manual_seed(7 + len(mode))
    m = CplxConv2d(C, O, k, stride=stride, padding=padding, dilation=dilation).to(DEV)
    z = cplx.randn(B, C, H, W, device=DEV)
    bf16 = mode.startswith("bf16")
    if bf16:
        m, z = m.bfloat16(), z.to(torch.bfloat16)
    c = lambda t: t.detach().cpu().double()
    want = orc.cplx_conv2d(c(z.real), c(z.imag), c(m.weight.real), c(m.weight.imag), c(m.bias.real),
                           c(m.bias.imag), m.stride, m.padding, m.dilation)
    if mode.endswith("_cl"):
        z = cplx.Cplx(z.real.contiguous(memory_format=torch.channels_last),
                      z.imag.contiguous(memory_format=torch.channels_last))
    ops.set_math_mode("tf32" if mode == "tf32" else "tensor")
    try:
        with torch.no_grad():
            out = m(z)
    finally:
        ops.set_math_mode("auto")
    tol = 1e-2 if bf16 else 1e-3
    assert out.shape == want[0].shape
    assert rel_err(out.real.float(), want[0]) < tol and rel_err(out.imag.float(), want[1]) < tol


def test_fp32_nchw_conv_chunked_prepass_overlap():
    """Large fp32 NCHW batches run in chunks of images (pre-pass of chunk c + 1 on the caller's stream
    under the GEMM of chunk c on a high-priority side stream, launch_conv_f16): uneven chunks
    (21 images -> 6 + 6 + 6 + 3), images that take the fix-up conversion in either chunk, a second call
    right behind the first on the same workspace, and a call on a non-default stream."""
    torch.manual_seed(5)
    B, C, H, W, O = 21, 16, 40, 130, 16
    m = CplxConv2d(C, O, 3, padding=1).to(DEV)
    z = cplx.randn(B, C, H, W, device=DEV)
    scale = torch.ones(B, device=DEV)
    scale[3], scale[12], scale[20] = 1e-5, 3e4, 0.0
    z = cplx.Cplx(z.real * scale.view(B, 1, 1, 1), z.imag * scale.view(B, 1, 1, 1))
    z2 = cplx.Cplx(z.imag.clone(), z.real.clone())
    c = lambda t: t.detach().cpu().double()
    with torch.no_grad():
        out = m(z)
        out2 = m(z2)               # same workspace, enqueued while the first call's chunks are in flight
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            out3 = m(z)
        torch.cuda.current_stream().wait_stream(s)
    w = [c(m.weight.real), c(m.weight.imag), c(m.bias.real), c(m.bias.imag)]
    want = orc.cplx_conv2d(c(z.real), c(z.imag), *w, 1, 1, 1)
    want2 = orc.cplx_conv2d(c(z2.real), c(z2.imag), *w, 1, 1, 1)
    for b in range(B):      # per image: the scales differ by nine orders of magnitude
        if float(scale[b]) == 0.0:
            continue
        for got, ref in ((out.real, want[0]), (out.imag, want[1]), (out2.real, want2[0]), (out2.imag, want2[1])):
            bias_mag = float(m.bias.real.abs().max())
            err = float((c(got[b]) - ref[b]).abs().max()) / max(float(ref[b].abs().max()), bias_mag)
            assert err < 1e-3, (b, err)
    assert torch.equal(out3.real, out.real) and torch.equal(out3.imag, out.imag)
