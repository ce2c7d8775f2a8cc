"""Real-valued variational dropout / ARD linear layers.

Reference: ``cplxmodule/nn/relevance/real/{base,vd,ard}.py``.  Parameters and
state-dict keys are those of ``torch.nn.Linear`` plus ``log_sigma2`` (init -10).
"""
import torch

from ... import _native as nv
from ... import ops
from .base import BaseARD


class _RealGaussianLinear(torch.nn.Linear):
    """Local-reparameterisation forward: training draws ``mu + eps * sqrt(max(s2, 1e-8))``
    with ``s2 = x^2 . exp(log_sigma2)^T`` inside one fused kernel; eval returns ``mu``."""

    _kl_kind = None

    def __init__(self, in_features, out_features, bias=True):
        super().__init__(in_features, out_features, bias=bias)
        self.log_sigma2 = torch.nn.Parameter(torch.empty(*self.weight.shape))
        self.reset_variational_parameters()

    def reset_variational_parameters(self):
        self.log_sigma2.data.fill_(-10.0)

    def forward(self, input, eps=None):
        if not self.training:
            return ops.real_linear(input, self.weight, self.bias)
        kl_req = ops.kl_request(self._kl_kind, self.log_sigma2.shape[0])
        out = ops.real_linear_vd(input, self.weight, self.bias, self.log_sigma2, eps=eps,
                                 kl_req=kl_req)
        cache = self.__dict__.get("_kl_cache")
        if cache is None:
            cache = self.__dict__["_kl_cache"] = ops.FusedKLCache()
        cache.put((self.weight, self.log_sigma2), kl_req)
        return out

    @property
    def log_alpha(self):
        return ops.log_alpha(self.weight, None, self.log_sigma2)

    @property
    def penalty(self):
        return ops.kl(self._kl_kind, self.weight, None, self.log_sigma2, None)

    def _penalty_reduced(self, reduction):
        pre = None
        cache = self.__dict__.get("_kl_cache")
        if cache is not None and reduction in ("sum", "mean"):
            pre = cache.take((self.weight, self.log_sigma2))
        return ops.kl(self._kl_kind, self.weight, None, self.log_sigma2, reduction, precomputed=pre)

    def relevance(self, *, threshold, **kwargs):
        with torch.no_grad():
            return ops.log_alpha(self.weight, None, self.log_sigma2, threshold=threshold)

    __sparsity_ignore__ = ("log_sigma2",)

    def sparsity(self, *, threshold, **kwargs):
        n_relevant = float(self.relevance(threshold=threshold).sum().item())
        return [(id(self.weight), self.weight.numel() - n_relevant)]


class LinearVD(_RealGaussianLinear, BaseARD):
    """Variational dropout, softplus-sigmoid KL approximation of arXiv:1701.05369."""
    _kl_kind = nv.KL_REAL_VD


class LinearARD(_RealGaussianLinear, BaseARD):
    """Automatic relevance determination: ``0.5 * softplus(-log_alpha)``."""
    _kl_kind = nv.KL_REAL_ARD
