"""Alternative penalties of the complex variational-dropout layers.

Reference: ``cplxmodule/nn/relevance/extensions/complex.py``: ``*VDApprox`` (softplus-sigmoid
fit, :77-100) and ``*VDScaleFree`` (exact KL against the scale-free prior, :18-44).  Same
forward as the ``*VD`` layers; the penalty is one more ``kind`` of the KL kernels (so it is
also fused into the forward's operand pre-pass).  ``*VDBogus`` (:143-163) only exists in the
reference to dodge its host-side ``Ei``; with the device-side ``Ei`` it has no purpose here.
"""
from ... import _native as nv
from .base import BaseARD
from .complex import CplxLinearGaussian, _CplxConvGaussianMixin
from ..modules.conv import CplxConv1d, CplxConv2d


class CplxLinearVDApprox(CplxLinearGaussian, BaseARD):
    _kl_kind = nv.KL_CPLX_VD_APPROX


class CplxLinearVDScaleFree(CplxLinearGaussian, BaseARD):
    _kl_kind = nv.KL_CPLX_VD_SCALEFREE


class CplxConv1dVDApprox(_CplxConvGaussianMixin, CplxConv1d, BaseARD):
    _kl_kind = nv.KL_CPLX_VD_APPROX


class CplxConv2dVDApprox(_CplxConvGaussianMixin, CplxConv2d, BaseARD):
    _kl_kind = nv.KL_CPLX_VD_APPROX


class CplxConv1dVDScaleFree(_CplxConvGaussianMixin, CplxConv1d, BaseARD):
    _kl_kind = nv.KL_CPLX_VD_SCALEFREE


class CplxConv2dVDScaleFree(_CplxConvGaussianMixin, CplxConv2d, BaseARD):
    _kl_kind = nv.KL_CPLX_VD_SCALEFREE
