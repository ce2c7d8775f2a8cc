"""``CplxLinear`` / ``CplxBilinear``: same constructors, parameters and default init as the
reference (``cplxmodule/nn/modules/linear.py:24-117``); forward is one tcgen05 kernel (bilinear:
an elementwise outer-product kernel + the same tcgen05 kernel)."""
import math

from ... import cplx
from .. import init
from .base import CplxParameter, CplxToCplx


class CplxLinear(CplxToCplx):
    r"""Complex linear map :math:`z \mapsto W z + b`, :math:`W \in \mathbb{C}^{out \times in}`."""

    def __init__(self, in_features, out_features, bias=True):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.weight = CplxParameter(cplx.Cplx.empty(out_features, in_features))
        if bias:
            self.bias = CplxParameter(cplx.Cplx.empty(out_features))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        init.cplx_kaiming_uniform_(self.weight, a=math.sqrt(5))
        if self.bias is not None:
            fan_in, _ = init.get_fans(self.weight)
            bound = 1 / math.sqrt(fan_in)
            init.cplx_uniform_independent_(self.bias, -bound, bound)

    def forward(self, input):
        return cplx.linear(input, self.weight, self.bias)

    def extra_repr(self):
        return (f"in_features={self.in_features}, out_features={self.out_features}, "
                f"bias={self.bias is not None}")


class CplxBilinear(CplxToCplx):
    r"""Complex bilinear map :math:`(u, v) \mapsto (u^{H|\top} A_j v + b_j)_j`
    (``cplxmodule/nn/modules/linear.py:67-117``)."""

    def __init__(self, in1_features, in2_features, out_features, bias=True, conjugate=True):
        super().__init__()
        self.in1_features, self.in2_features = in1_features, in2_features
        self.out_features = out_features
        self.weight = CplxParameter(cplx.Cplx.empty(out_features, in1_features, in2_features))
        if bias:
            self.bias = CplxParameter(cplx.Cplx.empty(out_features))
        else:
            self.register_parameter("bias", None)
        self.conjugate = conjugate
        self.reset_parameters()

    def reset_parameters(self):
        init.cplx_kaiming_uniform_(self.weight, a=math.sqrt(5))
        if self.bias is not None:
            fan_in, _ = init.get_fans(self.weight)
            bound = 1 / math.sqrt(fan_in)
            init.cplx_uniform_independent_(self.bias, -bound, bound)

    def forward(self, input1, input2):
        return cplx.bilinear(input1, input2, self.weight, self.bias, self.conjugate)

    def extra_repr(self):
        return (f"in1_features={self.in1_features}, in2_features={self.in2_features}, "
                f"out_features={self.out_features}, bias={self.bias is not None}, "
                f"conjugate={self.conjugate}")
