"""Penalty collection over a network (reference: ``cplxmodule/nn/relevance/base.py``).

``isinstance(mod, BaseARD)`` is the discovery protocol, as in the reference
(``base.py:132-133``).  Layers of this package additionally expose
``_penalty_reduced(reduction)`` which runs the fused KL kernel (per-element penalty
and reduction in ONE pass over the parameters, nothing of shape ``weight.shape`` is
written); foreign ``BaseARD`` modules fall back to their own ``.penalty`` tensor.
"""
import torch


class BaseARD(torch.nn.Module):
    r"""Trait of layers with a variational-approximation penalty (KL term of the ELBO)."""

    @property
    def penalty(self):
        # NB inside torch.nn.Module a raising property would be swallowed by __getattr__
        raise NotImplementedError("Derived classes must compute their own penalty.")

    def relevance(self, **kwargs):
        raise NotImplementedError(
            "Derived classes must implement a float mask of relevant coefficients.")


def _check_reduction(reduction):
    if reduction is not None and reduction not in ("mean", "sum"):
        raise ValueError(f"`reduction` must be either `None`, `sum` or `mean`. Got {reduction}.")


def named_penalties(module, reduction="sum", prefix=""):
    """Yield ``(name, penalty)`` for every variational submodule (shared modules once)."""
    _check_reduction(reduction)
    for name, mod in module.named_modules(prefix=prefix):
        if not isinstance(mod, BaseARD):
            continue
        fused = getattr(mod, "_penalty_reduced", None)
        if fused is not None:
            yield name, fused(reduction)
            continue
        penalty = mod.penalty
        if reduction == "sum":
            penalty = penalty.sum()
        elif reduction == "mean":
            penalty = penalty.mean()
        yield name, penalty


def penalties(module, reduction="sum"):
    """Iterate over the penalties only: ``loss = nll + C * sum(penalties(model))``."""
    for _, penalty in named_penalties(module, reduction=reduction):
        yield penalty


def named_relevance(module, prefix="", **kwargs):
    for name, mod in module.named_modules(prefix=prefix):
        if isinstance(mod, BaseARD):
            yield name, mod.relevance(**kwargs).detach()


def compute_ard_masks(module, *, prefix="", **kwargs):
    """Dict of ``<name>.mask`` relevance masks (compatible with the reference's ``nn.masked``)."""
    if not isinstance(module, torch.nn.Module):
        return {}
    return {name + ("." if name else "") + "mask": mask
            for name, mask in named_relevance(module, prefix=prefix, **kwargs)}
