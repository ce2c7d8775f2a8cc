"""Fixed-sparsity ("masked") layers: the third stage of the reference's
dense -> variational -> masked pipeline (``cplxmodule/nn/masked/{base,real,complex}.py``,
``nn/relevance/README.md:77-89``).  Linear layers: ONE C-ABI call in which the mask is applied
where the weights are staged for the GEMM (``cplxk_linear_masked_fwd``: inside the operand
pre-pass on the fp32 tensor-core path -- no ``weight * mask`` tensor exists).  Conv / bilinear
layers: the accelerated kernel on ``weight * mask`` (their weights are a few KB).  Masks arrive
from ``relevance.compute_ard_masks`` through ``deploy_masks`` or
``load_state_dict(..., strict=False)`` under the key ``<layer>.mask``."""
import torch

from .. import conv_ops, cplx, ops
from .modules.conv import CplxConv1d, CplxConv2d
from .modules.linear import CplxBilinear, CplxLinear


class BaseMasked(torch.nn.Module):
    """Holds an optional ``mask`` buffer shaped like ``.weight``; must come last among the
    bases (before ``torch.nn.Module``) so that the layer's own ``__init__`` runs first."""

    def __init__(self):
        super().__init__()
        self.register_buffer("mask", None)

    @property
    def is_sparse(self):
        return isinstance(self.mask, torch.Tensor)

    def mask_(self, mask):
        """Install (tensor) or remove (None) the mask; device / dtype / broadcasting follow
        ``.weight``."""
        if mask is None:
            if self.is_sparse:
                del self.mask
                self.register_buffer("mask", None)
            return self
        if not isinstance(mask, torch.Tensor):
            raise TypeError(f"`mask` must be either a Tensor or `None`. Got {type(mask).__name__}.")
        w = self.weight
        mask = mask.detach().to(w.device, w.dtype).expand(w.shape).contiguous()
        self.register_buffer("mask", mask)
        return self

    def __setattr__(self, name, value):
        if name == "mask":
            self.mask_(value)
        else:
            super().__setattr__(name, value)

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys,
                              unexpected_keys, error_msgs):
        key = prefix + "mask"
        rest = {k: v for k, v in state_dict.items() if k != key}
        super()._load_from_state_dict(rest, prefix, local_metadata, strict, missing_keys,
                                      unexpected_keys, error_msgs)
        reported = key in missing_keys
        if key in state_dict:
            if reported:
                missing_keys.remove(key)
            self.mask_(state_dict[key])
        elif strict and not reported:
            missing_keys.append(key)
        elif not strict and reported:
            missing_keys.remove(key)


class MaskedWeightMixin:
    @property
    def weight_masked(self):
        if not self.is_sparse:
            raise RuntimeError(f"`{type(self).__name__}` has no sparsity mask. Please, either set "
                               "a mask attribute, or call `deploy_masks()`.")
        return self.weight * self.mask

    def _effective_weight(self):
        return self.weight_masked if self.is_sparse else self.weight

    __sparsity_ignore__ = ("mask",)

    def sparsity(self, **kwargs):
        w = self.weight
        planes = (w.real, w.imag) if isinstance(w, cplx.Cplx) else (w,)
        n_dropped = float(planes[0].numel() - self.mask.sum().item()) if self.is_sparse else 0.0
        return [(id(p), n_dropped) for p in planes]


class LinearMasked(MaskedWeightMixin, torch.nn.Linear, BaseMasked):
    def forward(self, input):
        if not self.is_sparse:
            return ops.real_linear(input, self.weight, self.bias)
        return ops.real_linear_masked(input, self.weight, self.mask, self.bias)


class CplxLinearMasked(MaskedWeightMixin, CplxLinear, BaseMasked):
    def forward(self, input):
        if not self.is_sparse:
            return cplx.linear(input, self.weight, self.bias)
        w, b = self.weight, self.bias
        b_re, b_im = (None, None) if b is None else (b.real, b.imag)
        re, im = ops.cplx_linear_masked(input.real, input.imag, w.real, w.imag, self.mask, b_re, b_im)
        return cplx.Cplx(re, im)


class BilinearMasked(MaskedWeightMixin, torch.nn.Bilinear, BaseMasked):
    """nn/masked/real.py:69-72"""

    def forward(self, input1, input2):
        return ops.real_bilinear(input1, input2, self._effective_weight(), self.bias)


class CplxBilinearMasked(MaskedWeightMixin, CplxBilinear, BaseMasked):
    """nn/masked/complex.py:38-40"""

    def forward(self, input1, input2):
        return cplx.bilinear(input1, input2, self._effective_weight(), self.bias, self.conjugate)


class _RealConvMasked(MaskedWeightMixin):
    _nd = None

    def forward(self, input):
        if isinstance(self.padding, str) or self.padding_mode != "zeros":
            raise ValueError("only numeric zero padding is supported by the CUDA conv path")
        return conv_ops.real_convnd(self._nd, input, self._effective_weight(), self.bias, self.stride,
                                    self.padding, self.dilation, self.groups)


class Conv1dMasked(_RealConvMasked, torch.nn.Conv1d, BaseMasked):
    """nn/masked/real.py:30-40"""
    _nd = 1


class Conv2dMasked(_RealConvMasked, torch.nn.Conv2d, BaseMasked):
    """nn/masked/real.py:43-53"""
    _nd = 2


class CplxConv1dMasked(MaskedWeightMixin, CplxConv1d, BaseMasked):
    def forward(self, input):
        return cplx.conv1d(input, self._effective_weight(), self.bias, self.stride, self.padding,
                           self.dilation, self.groups, self.padding_mode)


class CplxConv2dMasked(MaskedWeightMixin, CplxConv2d, BaseMasked):
    def forward(self, input):
        return cplx.conv2d(input, self._effective_weight(), self.bias, self.stride, self.padding,
                           self.dilation, self.groups, self.padding_mode)


def is_sparse(module):
    return isinstance(module, BaseMasked) and module.is_sparse


def named_masks(module, prefix=""):
    for name, mod in module.named_modules(prefix=prefix):
        if isinstance(mod, BaseMasked):
            yield name, mod.mask


def deploy_masks(module, *, state_dict=None, prefix="", reset=False):
    """Set the masks listed in ``state_dict`` (``<name>.mask`` -> tensor or None); ``reset=True``
    also clears the masks of layers that are not listed."""
    if not isinstance(state_dict, dict) or not isinstance(module, torch.nn.Module):
        return module
    for name, mod in module.named_modules(prefix=prefix):
        if isinstance(mod, BaseMasked):
            key = name + ("." if name else "") + "mask"
            if key in state_dict:
                mod.mask = state_dict[key]
            elif reset:
                mod.mask = None
    return module


def binarize_masks(state_dict, masks):
    """Fold (possibly soft) masks into the weights and make the masks 0/1."""
    with torch.no_grad():
        new_state, new_masks = {}, {}
        for name, par in state_dict.items():
            if "weight" in name:
                key = name.rsplit("weight", 1)[0] + "mask"
                if key in masks:
                    par = par * masks[key].to(par)
                    par[par == 0] = 0          # drop the sign of negative zeros
            new_state[name] = par
        for name, mask in masks.items():
            new_masks[name] = torch.ne(mask, 0).to(mask)
    return new_state, new_masks
