"""Split real/imaginary complex tensor (`Cplx`) and the functional layer.

API mirror of the reference's ``cplxmodule/cplx.py`` for the hot path: the
container semantics follow ``cplx.py:10-376`` and the functional entry points
``linear`` (``:634-698``), ``conv1d/conv2d`` (``:770-838``), ``randn`` /
``randn_like`` (``:544-562``) keep their signatures, but ``linear`` and
``conv*`` execute as hand-written sm_100a kernels (``ops`` -> C ABI) instead
of four ``F.linear`` / two ``F.conv`` calls.  Arithmetic on the container
itself stays plain elementwise torch ops on whatever device the planes live on
(outside the accelerated path, exactly like the reference).
"""
import copy
import math

import torch

from . import ops as _ops


def _is_cplx_like(v):
    return isinstance(v, (Cplx, complex))


class Cplx:
    """Pair of same-shape real tensors ``(real, imag)``; never copies its inputs."""

    __slots__ = ("_re", "_im")

    def __new__(cls, real, imag=None):
        if isinstance(real, cls):
            return real
        if isinstance(real, complex):
            real, imag = torch.tensor(real.real), torch.tensor(real.imag)
        elif isinstance(real, float):
            if imag is None:
                imag = 0.0
            elif not isinstance(imag, float):
                raise TypeError("Imaginary part must be float.")
            real, imag = torch.tensor(real), torch.tensor(imag)
        elif not isinstance(real, torch.Tensor):
            raise TypeError("Real part must be torch.Tensor.")
        if imag is None:
            imag = torch.zeros_like(real)
        elif not isinstance(imag, torch.Tensor):
            raise TypeError("Imaginary part must be torch.Tensor.")
        if real.shape != imag.shape:
            raise ValueError("Real and imaginary parts have mistmatching shape.")
        obj = object.__new__(cls)
        object.__setattr__(obj, "_re", real)
        object.__setattr__(obj, "_im", imag)
        return obj

    def __setattr__(self, name, value):
        raise AttributeError("Cplx is immutable: build a new Cplx(real, imag) instead.")

    # ---- components
    @property
    def real(self):
        return self._re

    @property
    def imag(self):
        return self._im

    @property
    def conj(self):
        return Cplx(self._re, -self._im)

    def conjugate(self):
        return self.conj

    @property
    def angle(self):
        return torch.atan2(self._im, self._re)

    def __abs__(self):
        return torch.sqrt(self._re * self._re + self._im * self._im)

    def apply(self, f, *args, **kwargs):
        """Apply ``f`` to both planes independently."""
        return Cplx(f(self._re, *args, **kwargs), f(self._im, *args, **kwargs))

    # ---- copies
    def __copy__(self):
        return Cplx(self._re, self._im)

    def __deepcopy__(self, memo):
        return Cplx(copy.deepcopy(self._re, memo), copy.deepcopy(self._im, memo))

    def clone(self):
        return self.apply(torch.clone)

    def detach(self):
        return Cplx(self._re.detach(), self._im.detach())

    def requires_grad_(self, requires_grad=True):
        return Cplx(self._re.requires_grad_(requires_grad), self._im.requires_grad_(requires_grad))

    @property
    def grad(self):
        g_re, g_im = self._re.grad, self._im.grad
        return None if g_re is None or g_im is None else Cplx(g_re, g_im)

    # ---- indexing / iteration
    def __getitem__(self, key):
        return Cplx(self._re[key], self._im[key])

    def __setitem__(self, key, value):
        if _is_cplx_like(value):
            self._re[key], self._im[key] = value.real, value.imag
        else:
            self._re[key], self._im[key] = value, value

    def __iter__(self):
        return (Cplx(r, i) for r, i in zip(self._re, self._im))

    def __reversed__(self):
        return Cplx(reversed(self._re), reversed(self._im))

    def __len__(self):
        return self.shape[0]

    # ---- arithmetic (elementwise torch ops, not the accelerated path)
    def __pos__(self):
        return self

    def __neg__(self):
        return Cplx(-self._re, -self._im)

    def __add__(self, other):
        if _is_cplx_like(other):
            return Cplx(self._re + other.real, self._im + other.imag)
        return Cplx(self._re + other, self._im)

    __radd__ = __add__
    __iadd__ = __add__

    def __sub__(self, other):
        if _is_cplx_like(other):
            return Cplx(self._re - other.real, self._im - other.imag)
        return Cplx(self._re - other, self._im)

    def __rsub__(self, other):
        return (-self) + other

    __isub__ = __sub__

    def __mul__(self, other):
        if not _is_cplx_like(other):
            return Cplx(self._re * other, self._im * other)
        a, b, c, d = self._re, self._im, other.real, other.imag
        return Cplx(a * c - b * d, b * c + a * d)

    __rmul__ = __mul__
    __imul__ = __mul__

    def __truediv__(self, other):
        if not _is_cplx_like(other):
            return Cplx(self._re / other, self._im / other)
        norm2 = other.real * other.real + other.imag * other.imag
        return self * (Cplx(other.real, -other.imag) / norm2)

    def __rtruediv__(self, other):
        norm2 = self._re * self._re + self._im * self._im
        return (self.conj / norm2) * other

    __itruediv__ = __truediv__

    def __matmul__(self, other):
        if not isinstance(other, Cplx):
            return Cplx(torch.matmul(self._re, other), torch.matmul(self._im, other))
        re = torch.matmul(self._re, other._re) - torch.matmul(self._im, other._im)
        im = torch.matmul(self._im, other._re) + torch.matmul(self._re, other._im)
        return Cplx(re, im)

    def __rmatmul__(self, other):
        return Cplx(torch.matmul(other, self._re), torch.matmul(other, self._im))

    __imatmul__ = __matmul__

    # ---- shape
    @property
    def shape(self):
        return self._re.shape

    def size(self, *dim):
        return self._re.size(*dim)

    def dim(self):
        return self._re.dim()

    def t(self):
        return Cplx(self._re.t(), self._im.t())

    def h(self):
        return self.conj.t()

    def flatten(self, start_dim=0, end_dim=-1):
        return self.apply(torch.flatten, start_dim, end_dim)

    @staticmethod
    def _shape_arg(shape):
        return shape[0] if shape and isinstance(shape[0], (tuple, torch.Size)) else shape

    def view(self, *shape):
        shape = self._shape_arg(shape)
        return Cplx(self._re.view(*shape), self._im.view(*shape))

    def view_as(self, other):
        return self.view(*other.shape)

    def reshape(self, *shape):
        shape = self._shape_arg(shape)
        return Cplx(self._re.reshape(*shape), self._im.reshape(*shape))

    def squeeze(self, dim=None):
        if dim is None:
            return Cplx(self._re.squeeze(), self._im.squeeze())
        return Cplx(self._re.squeeze(dim=dim), self._im.squeeze(dim=dim))

    def unsqueeze(self, dim):
        return Cplx(self._re.unsqueeze(dim=dim), self._im.unsqueeze(dim=dim))

    def permute(self, *dims):
        return Cplx(self._re.permute(*dims), self._im.permute(*dims))

    def transpose(self, dim0, dim1):
        return Cplx(self._re.transpose(dim0, dim1), self._im.transpose(dim0, dim1))

    # ---- placement / conversion
    def cuda(self, device=None, non_blocking=False):
        return Cplx(self._re.cuda(device=device, non_blocking=non_blocking),
                    self._im.cuda(device=device, non_blocking=non_blocking))

    def cpu(self):
        return Cplx(self._re.cpu(), self._im.cpu())

    def to(self, *args, **kwargs):
        return Cplx(self._re.to(*args, **kwargs), self._im.to(*args, **kwargs))

    @property
    def device(self):
        return self._re.device

    @property
    def dtype(self):
        return self._re.dtype

    def is_complex(self):
        return True

    def item(self):
        return complex(float(self._re), float(self._im))

    @classmethod
    def from_numpy(cls, array):
        return cls(torch.from_numpy(array.real.copy()), torch.from_numpy(array.imag.copy()))

    def numpy(self):
        return self._re.numpy() + 1j * self._im.numpy()

    def __repr__(self):
        return f"{type(self).__name__}(\n  real={self._re},\n  imag={self._im}\n)"

    # ---- factories
    @classmethod
    def empty(cls, *sizes, dtype=None, device=None, requires_grad=False):
        re = torch.empty(*sizes, dtype=dtype, device=device, requires_grad=requires_grad)
        return cls(re, torch.empty_like(re, requires_grad=requires_grad))

    @classmethod
    def zeros(cls, *sizes, dtype=None, device=None, requires_grad=False):
        re = torch.zeros(*sizes, dtype=dtype, device=device, requires_grad=requires_grad)
        return cls(re, torch.zeros_like(re, requires_grad=requires_grad))

    @classmethod
    def ones(cls, *sizes, dtype=None, device=None, requires_grad=False):
        re = torch.ones(*sizes, dtype=dtype, device=device, requires_grad=requires_grad)
        return cls(re, torch.zeros_like(re, requires_grad=requires_grad))


# --------------------------------------------------------------- structural helpers
def _pairwise(fn, tensors, *args, **kwargs):
    tensors = [Cplx(z) for z in tensors]
    return Cplx(fn([z.real for z in tensors], *args, **kwargs),
                fn([z.imag for z in tensors], *args, **kwargs))


def cat(tensors, dim):
    return _pairwise(torch.cat, tensors, dim=dim)


def stack(tensors, dim):
    return _pairwise(torch.stack, tensors, dim=dim)


def _multi(fn, input, *args, **kwargs):
    return tuple(Cplx(r, i) for r, i in zip(fn(input.real, *args, **kwargs),
                                            fn(input.imag, *args, **kwargs)))


def split(input, split_size_or_sections, dim=0):
    return _multi(torch.split, input, split_size_or_sections, dim)


def chunk(input, chunks, dim=0):
    return _multi(torch.chunk, input, chunks, dim)


def unbind(input, dim=0):
    return _multi(torch.unbind, input, dim)


def from_interleaved_real(input, copy=True, dim=-1):
    """``[..., 2*D]`` with (re, im) interleaved along ``dim`` -> Cplx ``[..., D]``."""
    dim = dim if dim >= 0 else input.dim() + dim
    if input.shape[dim] % 2:
        raise ValueError("the interleaved dimension must have even size")
    index = [slice(None)] * input.dim()
    index[dim] = slice(0, None, 2)
    re = input[tuple(index)]
    index[dim] = slice(1, None, 2)
    im = input[tuple(index)]
    out = Cplx(re, im)
    return out.clone() if copy else out


from_real = from_interleaved_real


def from_concatenated_real(input, copy=True, dim=-1):
    out = Cplx(*torch.chunk(input, 2, dim=dim))
    return out.clone() if copy else out


def to_interleaved_real(input, flatten=True, dim=-1):
    dim = dim if dim >= 0 else input.dim() + dim
    out = torch.stack([input.real, input.imag], dim=dim + 1)
    return out.flatten(dim, dim + 1) if flatten else out


to_real = to_interleaved_real


def to_concatenated_real(input, flatten=None, dim=-1):
    assert flatten is None
    return torch.cat([input.real, input.imag], dim=dim)


# ------------------------------------------------------------------------- noise
def randn(*size, dtype=None, device=None, requires_grad=False):
    """Standard circular complex Gaussian: ONE ``torch.randn(2, *size) / sqrt(2)`` call,
    plane 0 -> real, plane 1 -> imag (this fixes the RNG consumption order the fused
    kernel's in-epilogue Philox reproduces)."""
    planes = torch.randn(2, *size, dtype=dtype, device=device) / math.sqrt(2)
    z = Cplx(planes[0], planes[1])
    return z.requires_grad_(True) if requires_grad else z


def randn_like(input, dtype=None, device=None, requires_grad=False):
    return randn(*input.size(), dtype=input.dtype if dtype is None else dtype,
                 device=input.device if device is None else device, requires_grad=requires_grad)


# ----------------------------------------------------------------- accelerated ops
def linear(input, weight, bias=None):
    """Complex affine map ``y = x W^T + b`` as ONE tcgen05 kernel (4 MMAs per k-step into
    two TMEM accumulators, bias in the epilogue)."""
    b_re, b_im = (None, None) if bias is None else (bias.real, bias.imag)
    re, im = _ops.cplx_linear(input.real, input.imag, weight.real, weight.imag, b_re, b_im)
    return Cplx(re, im)


# the reference exposes three arithmetic variants of the same map; here they are one kernel
linear_naive = linear
linear_cat = linear
linear_3m = linear


def bilinear(input1, input2, weight, bias=None, conjugate=True):
    """Complex bilinear map ``y_j = x1^{H|T} A_j x2 + b_j`` (reference: ``bilinear_naive``,
    ``cplx.py:1062-1090``): an outer-product kernel + the complex affine tcgen05 kernel on
    ``[.., in1 * in2]`` features."""
    b_re, b_im = (None, None) if bias is None else (bias.real, bias.imag)
    re, im = _ops.cplx_bilinear(input1.real, input1.imag, input2.real, input2.imag, weight.real,
                                weight.imag, b_re, b_im, conjugate)
    return Cplx(re, im)


bilinear_naive = bilinear
bilinear_cat = bilinear


def conv2d(input, weight, bias=None, stride=1, padding=0, dilation=1, groups=1,
           padding_mode="zeros"):
    """Complex 2-d cross-correlation (no kernel flip / conjugation), ``B x C x H x W``."""
    from . import conv_ops
    return conv_ops.cplx_convnd(2, input, weight, bias, stride, padding, dilation, groups,
                                padding_mode)


def conv1d(input, weight, bias=None, stride=1, padding=0, dilation=1, groups=1,
           padding_mode="zeros"):
    """Complex 1-d cross-correlation, ``B x C x L``."""
    from . import conv_ops
    return conv_ops.cplx_convnd(1, input, weight, bias, stride, padding, dilation, groups,
                                padding_mode)
