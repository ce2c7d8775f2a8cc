"""Real <-> complex layout casts needed to put the hot-path layers into a model
(pure torch views; reference: ``cplxmodule/nn/modules/casting.py``)."""
import torch

from ... import cplx
from .base import BaseCplxToReal, BaseRealToCplx


class InterleavedRealToCplx(BaseRealToCplx):
    def __init__(self, copy=False, dim=-1):
        super().__init__()
        self.copy, self.dim = copy, dim

    def forward(self, input):
        return cplx.from_interleaved_real(input, self.copy, self.dim)


RealToCplx = InterleavedRealToCplx


class ConcatenatedRealToCplx(BaseRealToCplx):
    def __init__(self, copy=False, dim=-1):
        super().__init__()
        self.copy, self.dim = copy, dim

    def forward(self, input):
        return cplx.from_concatenated_real(input, self.copy, self.dim)


class CplxToInterleavedReal(BaseCplxToReal):
    def __init__(self, dim=-1):
        super().__init__()
        self.dim = dim

    def forward(self, input):
        return cplx.to_interleaved_real(input, True, self.dim)


CplxToReal = CplxToInterleavedReal


class CplxToConcatenatedReal(BaseCplxToReal):
    def __init__(self, dim=-1):
        super().__init__()
        self.dim = dim

    def forward(self, input):
        return cplx.to_concatenated_real(input, None, self.dim)


class AsTypeCplx(BaseRealToCplx):
    def forward(self, input):
        return cplx.Cplx(input)


class CplxReal(BaseCplxToReal):
    def forward(self, input):
        return input.real


class CplxImag(BaseCplxToReal):
    def forward(self, input):
        return input.imag
