"""Complex parameters and the complex-to-complex module base.

Same state-dict format as the reference (``weight.real`` / ``weight.imag`` ...,
``cplxmodule/nn/modules/base.py:8-130``) so checkpoints move freely between the two.
"""
import functools

import torch

from ...cplx import Cplx


class CplxParameter(torch.nn.ParameterDict):
    """A complex parameter stored as two real ``nn.Parameter`` planes."""

    def __init__(self, cplx):
        if not isinstance(cplx, Cplx):
            raise TypeError(f"`{type(self).__name__}` accepts only Cplx tensors.")
        super().__init__({"real": torch.nn.Parameter(cplx.real),
                          "imag": torch.nn.Parameter(cplx.imag)})

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys,
                              unexpected_keys, error_msgs):
        re_key, im_key, whole = prefix + "real", prefix + "imag", prefix[:-1]
        has_re, has_im = re_key in state_dict, im_key in state_dict
        if not has_re and not has_im:
            if whole not in state_dict:
                missing_keys.append(whole)
                return
            # a real-valued tensor saved under the parameter's own name: promote R -> C
            value = state_dict[whole]
            state_dict = {re_key: value, im_key: torch.zeros_like(value)}
        elif has_re != has_im:
            error_msgs.append("Complex parameter requires both `.real` and `.imag` parts. "
                              f"Missing `{im_key if has_re else re_key}`.")
            return
        extra = [k for k in state_dict if k.startswith(prefix) and k not in (re_key, im_key)]
        if strict and extra:
            error_msgs.append(f"Complex parameter disallows redundant key(s) in state_dict: {extra}.")
        unexpected_keys.extend(extra)
        super()._load_from_state_dict({re_key: state_dict[re_key], im_key: state_dict[im_key]},
                                      prefix, local_metadata, strict, missing_keys, [], error_msgs)

    def extra_repr(self):
        return ", ".join(str(s) for s in self["real"].shape)

    @property
    def data(self):
        return Cplx(self["real"].data, self["imag"].data)


class CplxParameterAccessor:
    """Reading a ``CplxParameter`` attribute of a module yields a ``Cplx`` view of it."""

    def __getattr__(self, name):
        attr = super().__getattr__(name)
        if isinstance(attr, CplxParameter):
            return Cplx(attr["real"], attr["imag"])
        return attr


class BaseRealToCplx(torch.nn.Module):
    pass


class BaseCplxToReal(torch.nn.Module):
    pass


def _split_from_callable(fn):
    class SplitFn(CplxToCplx):
        def __init__(self, *args, **kwargs):
            super().__init__()
            self.args, self.kwargs = args, kwargs

        def forward(self, input):
            return input.apply(fn, *self.args, **self.kwargs)

    SplitFn.__name__ = f"CplxSplitFunc{getattr(fn, '__name__', 'fn').title()}"
    return SplitFn


def _split_from_module(Module):
    class SplitLayer(Module, CplxToCplx):
        def forward(self, input):
            return input.apply(super().forward)

    SplitLayer.__name__ = f"CplxSplitLayer{Module.__name__}"
    return SplitLayer


class _CplxToCplxMeta(type):
    """``CplxToCplx[torch.nn.ReLU]`` / ``CplxToCplx[torch.tanh]``: split activations."""

    @functools.lru_cache(maxsize=None)
    def __getitem__(cls, base):
        if isinstance(base, type) and issubclass(base, torch.nn.Module):
            if issubclass(base, (CplxToCplx, BaseRealToCplx)):
                return base
            return CplxToCplx if base is torch.nn.Module else _split_from_module(base)
        if callable(base):
            return _split_from_callable(base)
        raise TypeError("Expecting either a torch.nn.Module subclass, or a callable for "
                        f"promotion. Got `{type(base)}`.")


class CplxToCplx(CplxParameterAccessor, torch.nn.Module, metaclass=_CplxToCplxMeta):
    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        # real -> complex promotion of checkpoints written by real-valued layers: torch hands
        # each child only the keys under "<name>.", so the bare "<name>" entry is rewritten here
        for name, child in self._modules.items():
            key = prefix + name
            if (isinstance(child, CplxParameter) and key in state_dict
                    and key + ".real" not in state_dict and key + ".imag" not in state_dict):
                value = state_dict.pop(key)
                state_dict[key + ".real"] = value
                state_dict[key + ".imag"] = torch.zeros_like(value)
        super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)


def is_from_cplx(module):
    if isinstance(module, type):
        return issubclass(module, (CplxToCplx, BaseCplxToReal))
    if isinstance(module, torch.nn.Sequential):
        return is_from_cplx(module[0])
    return isinstance(module, (CplxToCplx, BaseCplxToReal))


def is_to_cplx(module):
    if isinstance(module, type):
        return issubclass(module, (CplxToCplx, BaseRealToCplx))
    if isinstance(module, torch.nn.Sequential):
        return is_to_cplx(module[-1])
    return isinstance(module, (CplxToCplx, BaseRealToCplx))


def is_cplx_to_cplx(module):
    return is_from_cplx(module) and is_to_cplx(module)
