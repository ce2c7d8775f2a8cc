"""Real-valued variational dropout / ARD linear and convolutional layers.

Reference: ``cplxmodule/nn/relevance/real/{base,vd,ard}.py``.  Parameters and
state-dict keys are those of ``torch.nn.Linear`` / ``torch.nn.Conv{1,2}d`` plus
``log_sigma2`` (init -10).
"""
import torch

from ... import _native as nv
from ... import ops
from .base import BaseARD


class _RealGaussianLinear(torch.nn.Linear):
    """Local-reparameterisation forward: training draws ``mu + eps * sqrt(max(s2, 1e-8))``
    with ``s2 = x^2 . exp(log_sigma2)^T`` inside one fused kernel; eval returns ``mu``."""

    _kl_kind = None

    def __init__(self, in_features, out_features, bias=True):
        super().__init__(in_features, out_features, bias=bias)
        self.log_sigma2 = torch.nn.Parameter(torch.empty(*self.weight.shape))
        self.reset_variational_parameters()

    def reset_variational_parameters(self):
        self.log_sigma2.data.fill_(-10.0)
        self.__dict__.pop("_kl_cache", None)   # `.data` writes do not bump the version counter

    def forward(self, input, eps=None):
        if not self.training:
            return ops.real_linear(input, self.weight, self.bias)
        kl_req = ops.kl_request(self._kl_kind, self.log_sigma2.shape[0])
        out = ops.real_linear_vd(input, self.weight, self.bias, self.log_sigma2, eps=eps,
                                 kl_req=kl_req)
        cache = self.__dict__.get("_kl_cache")
        if cache is None:
            cache = self.__dict__["_kl_cache"] = ops.FusedKLCache()
        cache.put((self.weight, self.log_sigma2), kl_req)
        return out

    @property
    def log_alpha(self):
        return ops.log_alpha(self.weight, None, self.log_sigma2)

    @property
    def penalty(self):
        return ops.kl(self._kl_kind, self.weight, None, self.log_sigma2, None)

    def _penalty_reduced(self, reduction):
        pre = None
        cache = self.__dict__.get("_kl_cache")
        if cache is not None and reduction in ("sum", "mean"):
            pre = cache.take((self.weight, self.log_sigma2))
        return ops.kl(self._kl_kind, self.weight, None, self.log_sigma2, reduction, precomputed=pre)

    def relevance(self, *, threshold, **kwargs):
        with torch.no_grad():
            return ops.log_alpha(self.weight, None, self.log_sigma2, threshold=threshold)

    __sparsity_ignore__ = ("log_sigma2",)

    def sparsity(self, *, threshold, **kwargs):
        n_relevant = float(self.relevance(threshold=threshold).sum().item())
        return [(id(self.weight), self.weight.numel() - n_relevant)]


LinearGaussian = _RealGaussianLinear      # the reference's name (real/base.py:29)


# class hierarchy as in the reference (real/vd.py:79-127, real/ard.py:42-66): ARD layers are
# VD layers with another penalty
class LinearVD(_RealGaussianLinear, BaseARD):
    """Variational dropout, softplus-sigmoid KL approximation of arXiv:1701.05369."""
    _kl_kind = nv.KL_REAL_VD


class LinearARD(LinearVD):
    """Automatic relevance determination: ``0.5 * softplus(-log_alpha)``."""
    _kl_kind = nv.KL_REAL_ARD


class BilinearGaussian(torch.nn.Bilinear):
    """``torch.nn.Bilinear`` with the local-reparameterisation forward (real/base.py:52-80):
    an outer-product kernel, then the fused linear kernel on ``[.., in1 * in2]`` features with
    the weight viewed as ``[out, in1 * in2]``."""

    _kl_kind = None
    __sparsity_ignore__ = ("log_sigma2",)

    def __init__(self, in1_features, in2_features, out_features, bias=True):
        super().__init__(in1_features, in2_features, out_features, bias=bias)
        self.log_sigma2 = torch.nn.Parameter(torch.empty(*self.weight.shape))
        self.reset_variational_parameters()

    def reset_variational_parameters(self):
        self.log_sigma2.data.fill_(-10.0)

    def forward(self, input1, input2, eps=None):
        ls2 = self.log_sigma2 if self.training else None
        return ops.real_bilinear(input1, input2, self.weight, self.bias, log_sigma2=ls2,
                                 eps=eps if self.training else None)

    @property
    def log_alpha(self):
        return ops.log_alpha(self.weight, None, self.log_sigma2)

    @property
    def penalty(self):
        return ops.kl(self._kl_kind, self.weight, None, self.log_sigma2, None)

    def _penalty_reduced(self, reduction):
        return ops.kl(self._kl_kind, self.weight, None, self.log_sigma2, reduction)

    def relevance(self, *, threshold, **kwargs):
        with torch.no_grad():
            return ops.log_alpha(self.weight, None, self.log_sigma2, threshold=threshold)

    def sparsity(self, *, threshold, **kwargs):
        n_relevant = float(self.relevance(threshold=threshold).sum().item())
        return [(id(self.weight), self.weight.numel() - n_relevant)]


class BilinearVD(BilinearGaussian, BaseARD):
    """Bilinear layer with variational dropout (real/vd.py:90-98)."""
    _kl_kind = nv.KL_REAL_VD


class BilinearARD(BilinearVD):
    """Bilinear layer with automatic relevance determination (real/ard.py:66-69)."""
    _kl_kind = nv.KL_REAL_ARD


class _RealGaussianConvNd:
    """``ConvNdGaussianMixin`` (nn/relevance/real/base.py:83-163) on the CUDA conv kernels: the mean
    conv, the variance conv ``x^2 * exp(log_sigma2)``, the noise and ``mu + eps sqrt(max(s2, 1e-8))``
    are one C-ABI call.  Zero padding only, as in the reference (:109-112)."""

    _kl_kind = None
    _nd = None
    __sparsity_ignore__ = ("log_sigma2",)

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1,
                 groups=1, bias=True, padding_mode="zeros"):
        super().__init__(in_channels, out_channels, kernel_size, stride=stride, padding=padding,
                         dilation=dilation, groups=groups, bias=bias, padding_mode=padding_mode)
        if self.padding_mode != "zeros":
            raise ValueError(f"Only `zeros` padding mode is supported. Got `{self.padding_mode}`.")
        self.log_sigma2 = torch.nn.Parameter(torch.empty(*self.weight.shape))
        self.reset_variational_parameters()

    def reset_variational_parameters(self):
        self.log_sigma2.data.fill_(-10.0)
        self.__dict__.pop("_kl_cache", None)   # `.data` writes do not bump the version counter

    def forward(self, input, eps=None):
        from ... import conv_ops
        if isinstance(self.padding, str):
            raise ValueError("string padding modes are not supported by the CUDA conv path")
        ls2 = self.log_sigma2 if self.training else None
        return conv_ops.real_convnd(self._nd, input, self.weight, self.bias, self.stride,
                                    self.padding, self.dilation, self.groups, log_sigma2=ls2,
                                    eps=eps if self.training else None)

    @property
    def log_alpha(self):
        return ops.log_alpha(self.weight, None, self.log_sigma2)

    @property
    def penalty(self):
        return ops.kl(self._kl_kind, self.weight, None, self.log_sigma2, None)

    def _penalty_reduced(self, reduction):
        return ops.kl(self._kl_kind, self.weight, None, self.log_sigma2, reduction)

    def relevance(self, *, threshold, **kwargs):
        with torch.no_grad():
            return ops.log_alpha(self.weight, None, self.log_sigma2, threshold=threshold)

    def sparsity(self, *, threshold, **kwargs):
        n_relevant = float(self.relevance(threshold=threshold).sum().item())
        return [(id(self.weight), self.weight.numel() - n_relevant)]


class Conv1dGaussian(_RealGaussianConvNd, torch.nn.Conv1d):
    """real/base.py:166-177"""
    _nd = 1


class Conv2dGaussian(_RealGaussianConvNd, torch.nn.Conv2d):
    """real/base.py:180-191"""
    _nd = 2


class Conv1dVD(Conv1dGaussian, BaseARD):
    """1D convolution with variational dropout (nn/relevance/real/vd.py:103-113)."""
    _kl_kind = nv.KL_REAL_VD


class Conv2dVD(Conv2dGaussian, BaseARD):
    """2D convolution with variational dropout (nn/relevance/real/vd.py:115-125)."""
    _kl_kind = nv.KL_REAL_VD


class Conv1dARD(Conv1dVD):
    """1D convolution with automatic relevance determination (nn/relevance/real/ard.py:48-51)."""
    _kl_kind = nv.KL_REAL_ARD


class Conv2dARD(Conv2dVD):
    """2D convolution with automatic relevance determination (nn/relevance/real/ard.py:54-57)."""
    _kl_kind = nv.KL_REAL_ARD
