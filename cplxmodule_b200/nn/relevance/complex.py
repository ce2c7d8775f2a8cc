"""Complex-valued variational dropout / ARD layers.

Reference: ``cplxmodule/nn/relevance/complex/{base,vd,ard}.py``.  State-dict keys:
``weight.real, weight.imag, bias.real, bias.imag, log_sigma2``.
The exact complex-VD KL ``gamma - log_alpha - Ei(-1/alpha)`` is evaluated on the device
(the reference round-trips through host ``scipy.special.expi``, ``complex/vd.py:31-36``).
"""
import torch

from ... import _native as nv
from ... import cplx, ops
from ..modules.conv import CplxConv1d, CplxConv2d
from ..modules.linear import CplxBilinear, CplxLinear
from .base import BaseARD


class _CplxGaussianMixin:
    _kl_kind = None
    __sparsity_ignore__ = ("log_sigma2",)

    def reset_variational_parameters(self):
        self.log_sigma2.data.fill_(-10.0)
        self.__dict__.pop("_kl_cache", None)   # `.data` writes do not bump the version counter

    @property
    def log_alpha(self):
        w = self.weight
        return ops.log_alpha(w.real, w.imag, self.log_sigma2)

    @property
    def penalty(self):
        w = self.weight
        return ops.kl(self._kl_kind, w.real, w.imag, self.log_sigma2, None)

    def _penalty_reduced(self, reduction):
        w = self.weight
        pre = None
        cache = self.__dict__.get("_kl_cache")
        if cache is not None and reduction in ("sum", "mean"):
            pre = cache.take((w.real, w.imag, self.log_sigma2))
        return ops.kl(self._kl_kind, w.real, w.imag, self.log_sigma2, reduction, precomputed=pre)

    def relevance(self, *, threshold, **kwargs):
        w = self.weight
        with torch.no_grad():
            return ops.log_alpha(w.real, w.imag, self.log_sigma2, threshold=threshold)

    def sparsity(self, *, threshold, **kwargs):
        w = self.weight
        n_dropped = float(w.real.numel()) - float(self.relevance(threshold=threshold).sum().item())
        return [(id(w.real), n_dropped), (id(w.imag), n_dropped)]


class CplxLinearGaussian(_CplxGaussianMixin, CplxLinear):
    """Complex linear layer with the fused local-reparameterisation forward."""

    def __init__(self, in_features, out_features, bias=True):
        super().__init__(in_features, out_features, bias=bias)
        self.log_sigma2 = torch.nn.Parameter(torch.empty(out_features, in_features))
        self.reset_variational_parameters()

    def forward(self, input, eps=None):
        if not self.training:
            return super().forward(input)
        w, b = self.weight, self.bias
        b_re, b_im = (None, None) if b is None else (b.real, b.imag)
        eps = None if eps is None else (eps.real, eps.imag)
        # the operand pre-pass reads every weight: it hands back the KL sum for penalties()
        kl_req = ops.kl_request(self._kl_kind, self.log_sigma2.shape[0])
        re, im = ops.cplx_linear_vd(input.real, input.imag, w.real, w.imag, b_re, b_im,
                                    self.log_sigma2, eps=eps, kl_req=kl_req)
        cache = self.__dict__.get("_kl_cache")
        if cache is None:
            cache = self.__dict__["_kl_cache"] = ops.FusedKLCache()
        cache.put((w.real, w.imag, self.log_sigma2), kl_req)
        return cplx.Cplx(re, im)


class CplxBilinearGaussian(_CplxGaussianMixin, CplxBilinear):
    """Complex bilinear layer with the fused local-reparameterisation forward
    (``nn/relevance/complex/base.py:59-84``): ``s2 = bilinear(|x1|^2, |x2|^2, exp(log_sigma2))`` is
    the variance GEMM of the linear kernel on the outer-product features."""

    def __init__(self, in1_features, in2_features, out_features, bias=True, conjugate=True):
        super().__init__(in1_features, in2_features, out_features, bias=bias, conjugate=conjugate)
        self.log_sigma2 = torch.nn.Parameter(torch.empty(*self.weight.shape))
        self.reset_variational_parameters()

    def forward(self, input1, input2, eps=None):
        if not self.training:
            return super().forward(input1, input2)
        w, b = self.weight, self.bias
        b_re, b_im = (None, None) if b is None else (b.real, b.imag)
        eps = None if eps is None else (eps.real, eps.imag)
        re, im = ops.cplx_bilinear(input1.real, input1.imag, input2.real, input2.imag, w.real, w.imag,
                                   b_re, b_im, self.conjugate, log_sigma2=self.log_sigma2, eps=eps)
        return cplx.Cplx(re, im)


# class hierarchy as in the reference (complex/vd.py:102-126, complex/ard.py:42-74): the ARD
# layers ARE VD layers with another penalty, so `isinstance(layer, CplxLinearVD)` behaves alike
class CplxLinearVD(CplxLinearGaussian, BaseARD):
    """Complex variational dropout with the exact KL divergence."""
    _kl_kind = nv.KL_CPLX_VD


class CplxLinearARD(CplxLinearVD):
    """Complex ARD: ``softplus(-log_alpha)``."""
    _kl_kind = nv.KL_CPLX_ARD


class CplxBilinearVD(CplxBilinearGaussian, BaseARD):
    _kl_kind = nv.KL_CPLX_VD


class CplxBilinearARD(CplxBilinearVD):
    _kl_kind = nv.KL_CPLX_ARD


class _CplxConvGaussianMixin(_CplxGaussianMixin):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1,
                 groups=1, bias=True, padding_mode="zeros"):
        super().__init__(in_channels, out_channels, kernel_size, stride=stride, padding=padding,
                         dilation=dilation, groups=groups, bias=bias, padding_mode=padding_mode)
        if self.padding_mode != "zeros":
            raise ValueError(f"Only `zeros` padding mode is supported. Got `{self.padding_mode}`.")
        self.log_sigma2 = torch.nn.Parameter(torch.empty(*self.weight.shape))
        self.reset_variational_parameters()

    def forward(self, input, eps=None):
        if not self.training:
            return super().forward(input)
        from ... import conv_ops
        return conv_ops.cplx_convnd(len(self.kernel_size), input, self.weight, self.bias,
                                    self.stride, self.padding, self.dilation, self.groups,
                                    self.padding_mode, log_sigma2=self.log_sigma2, eps=eps)


class CplxConv1dVD(_CplxConvGaussianMixin, CplxConv1d, BaseARD):
    _kl_kind = nv.KL_CPLX_VD


class CplxConv2dVD(_CplxConvGaussianMixin, CplxConv2d, BaseARD):
    _kl_kind = nv.KL_CPLX_VD


class CplxConv1dARD(CplxConv1dVD):
    _kl_kind = nv.KL_CPLX_ARD


class CplxConv2dARD(CplxConv2dVD):
    _kl_kind = nv.KL_CPLX_ARD


# the reference's `*Gaussian` names (complex/base.py:138-149)
class CplxConv1dGaussian(_CplxConvGaussianMixin, CplxConv1d):
    pass


class CplxConv2dGaussian(_CplxConvGaussianMixin, CplxConv2d):
    pass
