"""Alternative penalties of the complex variational-dropout layers.

Reference: ``cplxmodule/nn/relevance/extensions/complex.py``: ``*VDApprox`` (softplus-sigmoid
fit, :77-100) and ``*VDScaleFree`` (exact KL against the scale-free prior, :18-44).  Same
forward as the ``*VD`` layers -- of which they are SUBCLASSES, as in the reference (:47-141) --
with one more ``kind`` of the KL kernels (so the penalty is also fused into the forward's
operand pre-pass).

``*VDBogus`` (:143-198) exists in the reference only to dodge its host-side ``Ei``: its penalty
has the CORRECT gradient but a deliberately bogus forward value (``Ei`` replaced by zeros).
With the device-side ``Ei`` there is nothing to dodge: the names are kept, for checkpoints and
``isinstance`` checks, as plain subclasses of the exact ``*VD`` layers -- same gradient as the
reference's, and the true penalty value instead of the bogus one.
"""
from ... import _native as nv
from .complex import CplxBilinearVD, CplxConv1dVD, CplxConv2dVD, CplxLinearVD


class CplxLinearVDApprox(CplxLinearVD):
    _kl_kind = nv.KL_CPLX_VD_APPROX


class CplxBilinearVDApprox(CplxBilinearVD):
    _kl_kind = nv.KL_CPLX_VD_APPROX


class CplxConv1dVDApprox(CplxConv1dVD):
    _kl_kind = nv.KL_CPLX_VD_APPROX


class CplxConv2dVDApprox(CplxConv2dVD):
    _kl_kind = nv.KL_CPLX_VD_APPROX


class CplxLinearVDScaleFree(CplxLinearVD):
    _kl_kind = nv.KL_CPLX_VD_SCALEFREE


class CplxBilinearVDScaleFree(CplxBilinearVD):
    _kl_kind = nv.KL_CPLX_VD_SCALEFREE


class CplxConv1dVDScaleFree(CplxConv1dVD):
    _kl_kind = nv.KL_CPLX_VD_SCALEFREE


class CplxConv2dVDScaleFree(CplxConv2dVD):
    _kl_kind = nv.KL_CPLX_VD_SCALEFREE


class CplxLinearVDBogus(CplxLinearVD):
    pass


class CplxBilinearVDBogus(CplxBilinearVD):
    pass


class CplxConv1dVDBogus(CplxConv1dVD):
    pass


class CplxConv2dVDBogus(CplxConv2dVD):
    pass
