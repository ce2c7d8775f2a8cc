"""Host-side simulation of the operand arithmetic the tensor-core kernels use (no GPU): the error
bounds DESIGN.md states for the per-image fp16 copies of the fp32 NCHW convolution input
(`f16_image_scale_exp`, csrc/conv_tc.cu) hold against the float64 oracle."""
import numpy as np
import pytest
import torch

from oracle import cplx_oracle as orc


def _f16_scale_exp(amax):
    """csrc/conv_tc.cu: f16_scale_exp -- the power of two that puts amax into [2^13, 2^14)"""
    if amax == 0 or not np.isfinite(amax):
        return 0
    return min(13 - int(np.floor(np.log2(amax))), 126)


def _image_scale_exp(amax):
    """csrc/conv_tc.cu: f16_image_scale_exp -- images with amax in [2^-2, 2^15) stay unscaled"""
    if amax == 0:
        return 0
    ex = int(np.floor(np.log2(amax)))
    return 0 if -2 <= ex <= 14 else _f16_scale_exp(amax)


def _to_f16(x, s):
    """round-to-nearest fp16 copy of x * 2^s, returned as float64 with the scale undone"""
    return (x * 2.0 ** s).astype(np.float16).astype(np.float64) * 2.0 ** -s


@pytest.mark.parametrize("img_scale", [0.3, 1.0, 37.0, 4000.0, 2.0 ** 14.9 / 5])
def test_unscaled_fp16_copy_is_as_accurate_as_the_scaled_one(img_scale):
    """An image whose largest magnitude lies in [2^-2, 2^15) is copied to fp16 UNSCALED by the
    optimistic pre-pass.  Against the float64 convolution the result is as accurate as with the
    amax-derived scale (every element above 2^-14 keeps its 11-bit significand; the absolute error of
    the ones below is <= 2^-25 <= 2^-23 of the image maximum) and within the 1e-3 bound."""
    rng = np.random.default_rng(7)
    C, H, W, O = 16, 12, 14, 8
    # heavy-tailed magnitudes: many elements far below the maximum (incl. fp16 subnormals unscaled)
    mag = np.exp(rng.normal(0.0, 4.0, size=(2, 1, C, H, W)))
    x = rng.standard_normal((2, 1, C, H, W)) * mag
    x *= img_scale / np.abs(x).max()
    w = rng.standard_normal((2, O, C, 3, 3)) / 12
    amax = float(np.abs(x).max())
    assert _image_scale_exp(amax) == 0
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    want = orc.cplx_conv2d(t(x[0]), t(x[1]), t(w[0]), t(w[1]), None, None)
    errs = []
    for s in (0, _f16_scale_exp(amax)):
        got = orc.cplx_conv2d(t(_to_f16(x[0], s)), t(_to_f16(x[1], s)), t(w[0]), t(w[1]), None, None)
        errs.append(max(float((got[k] - want[k]).abs().max() / want[k].abs().max()) for k in (0, 1)))
    unscaled, scaled = errs
    assert unscaled < 1e-3 and scaled < 1e-3
    assert unscaled <= 1.05 * scaled + 2.0 ** -22


@pytest.mark.parametrize("amax,expect", [(0.0, 0), (0.2499, 16), (0.25, 0), (1.0, 0), (32767.9, 0),
                                          (32768.0, -2), (65504.0, -2), (1e-6, 33), (3e4 * 4, -3)])
def test_image_scale_exponent_table(amax, expect):
    """the decision the conversion kernels and the GEMM epilogue share (both call
    f16_image_scale_exp on the same amax bits): unscaled inside [2^-2, 2^15), else amax -> [2^13, 2^14)"""
    assert _image_scale_exp(amax) == expect
    if amax > 0:
        scaled = amax * 2.0 ** expect
        assert scaled < 65504.0 and (expect == 0 or 2.0 ** 13 <= scaled < 2.0 ** 14)
