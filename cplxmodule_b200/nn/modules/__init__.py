from .base import CplxParameter, CplxToCplx, is_from_cplx, is_to_cplx, is_cplx_to_cplx
from .linear import CplxLinear, CplxBilinear
from .conv import CplxConv1d, CplxConv2d
from .casting import (InterleavedRealToCplx, RealToCplx, ConcatenatedRealToCplx,
                      CplxToInterleavedReal, CplxToReal, CplxToConcatenatedReal, AsTypeCplx,
                      CplxReal, CplxImag)
