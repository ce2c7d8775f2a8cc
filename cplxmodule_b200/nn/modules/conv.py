"""``CplxConv1d`` / ``CplxConv2d`` with the reference's constructor signature and
state-dict keys (``cplxmodule/nn/modules/conv.py:11-196``)."""
import math

from torch.nn.modules.utils import _pair, _single

from ... import cplx
from .. import init
from .base import CplxParameter, CplxToCplx


class CplxConvNd(CplxToCplx):
    _ntuple = None
    _functional = None

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1,
                 groups=1, bias=True, padding_mode="zeros"):
        super().__init__()
        if in_channels % groups != 0:
            raise ValueError("in_channels must be divisible by groups")
        if out_channels % groups != 0:
            raise ValueError("out_channels must be divisible by groups")
        nt = type(self)._ntuple
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size, self.stride = nt(kernel_size), nt(stride)
        self.padding, self.dilation = nt(padding), nt(dilation)
        self.groups, self.padding_mode = groups, padding_mode
        self.weight = CplxParameter(
            cplx.Cplx.empty(out_channels, in_channels // groups, *self.kernel_size))
        if bias:
            self.bias = CplxParameter(cplx.Cplx.empty(out_channels))
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
        return type(self)._functional(input, self.weight, self.bias, self.stride, self.padding,
                                      self.dilation, self.groups, self.padding_mode)

    def extra_repr(self):
        s = (f"{self.in_channels}, {self.out_channels}, kernel_size={self.kernel_size}, "
             f"stride={self.stride}")
        if any(p != 0 for p in self.padding):
            s += f", padding={self.padding}"
        if any(d != 1 for d in self.dilation):
            s += f", dilation={self.dilation}"
        if self.groups != 1:
            s += f", groups={self.groups}"
        if self.bias is None:
            s += ", bias=False"
        if self.padding_mode != "zeros":
            s += f", padding_mode={self.padding_mode}"
        return s


class CplxConv1d(CplxConvNd):
    r"""Complex 1D convolution :math:`F \colon \mathbb{C}^{B \times c_{in} \times L} \to
    \mathbb{C}^{B \times c_{out} \times L'}`."""
    _ntuple = staticmethod(_single)
    _functional = staticmethod(cplx.conv1d)


class CplxConv2d(CplxConvNd):
    r"""Complex 2D convolution on ``B x c_in x H x W``."""
    _ntuple = staticmethod(_pair)
    _functional = staticmethod(cplx.conv2d)
