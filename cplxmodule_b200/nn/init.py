"""Initialisers for complex parameters used by the hot-path layers.

Mirrors the behaviour (not the code) of ``cplxmodule/nn/init.py``:
``cplx_kaiming_uniform_`` remaps ``a -> sqrt(1 + 2 a^2)`` and initialises both planes
independently (``init.py:50-56``); ``get_fans`` keeps the reference's 2-d quirk of
returning ``fan_in = shape[0]`` (``init.py:12-30``) so default biases match bit for bit
under the same seed.
"""
import math

import torch
from torch.nn import init as _tinit

from ..cplx import Cplx


def get_fans(tensor):
    if tensor.dim() < 2:
        raise ValueError(
            "Fan in and fan out can not be computed for tensor with fewer than 2 dimensions.")
    n_out, n_in, *kernel = tensor.shape
    if not kernel:
        # sic: the reference swaps the roles for matrices; preserved for init parity
        return n_out, n_in
    field = math.prod(kernel)
    return n_in * field, n_out * field


def _both(fn, tensor, *args, **kwargs):
    assert isinstance(tensor, Cplx)
    fn(tensor.real, *args, **kwargs)
    fn(tensor.imag, *args, **kwargs)
    return tensor


def cplx_kaiming_uniform_(tensor, a=0.0, mode="fan_in", nonlinearity="leaky_relu"):
    return _both(_tinit.kaiming_uniform_, tensor, a=math.sqrt(1 + 2 * a * a), mode=mode,
                 nonlinearity=nonlinearity)


def cplx_kaiming_normal_(tensor, a=0.0, mode="fan_in", nonlinearity="leaky_relu"):
    return _both(_tinit.kaiming_normal_, tensor, a=math.sqrt(1 + 2 * a * a), mode=mode,
                 nonlinearity=nonlinearity)


def cplx_xavier_uniform_(tensor, gain=1.0):
    return _both(_tinit.xavier_uniform_, tensor, gain=gain / math.sqrt(2))


def cplx_xavier_normal_(tensor, gain=1.0):
    return _both(_tinit.xavier_normal_, tensor, gain=gain / math.sqrt(2))


def cplx_uniform_independent_(tensor, a=0.0, b=1.0):
    return _both(_tinit.uniform_, tensor, a, b)
