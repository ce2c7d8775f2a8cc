"""CPU oracle for the hot path -- TEST INFRASTRUCTURE, never on the product path.

Each function restates one reference function (ivannz/cplxmodule @ f199c57) with the
same torch / scipy calls the reference makes on CPU, so that timing it equals timing
the reference's CPU path, and cites the file:line it follows.

Pinning status: PINNED.  ``tests/test_oracle.py`` checks every function here
  (a) bit-for-bit against the live reference imported from /root/reference (when that
      tree is present, i.e. in the build container), and
  (b) against the committed fixtures ``tests/golden/*.npz`` which were generated from the
      live reference by ``oracle/make_golden.py`` (runs anywhere, incl. the GPU box).
The reference's own tests pin only Ei (tests/test_relevance.py:40-49), linear
(tests/test_cplx.py:251-269) and conv (tests/test_cplx.py:272-429) via float64 closed
forms; those closed forms are re-checked in tests/test_oracle.py as well.
"""
import math

import numpy as np
import scipy.special
import torch
import torch.nn.functional as F

EULER_GAMMA = float(np.euler_gamma)


# ----------------------------------------------------------------------------- noise
def cplx_randn(*size, dtype=None, generator=None):
    """cplx.randn, cplxmodule/cplx.py:544-550: ONE randn(2,*size)/sqrt(2); [0]->re, [1]->im."""
    normal = torch.randn(2, *size, dtype=dtype, generator=generator) / math.sqrt(2)
    return normal[0], normal[1]


# ---------------------------------------------------------------------------- linear
def cplx_linear(x_re, x_im, w_re, w_im, b_re=None, b_im=None):
    """cplx.linear == linear_naive, cplxmodule/cplx.py:634-648,698."""
    re = F.linear(x_re, w_re) - F.linear(x_im, w_im)
    im = F.linear(x_re, w_im) + F.linear(x_im, w_re)
    if b_re is not None:
        re, im = re + b_re, im + b_im
    return re, im


def cplx_linear_vd(x_re, x_im, w_re, w_im, b_re, b_im, log_sigma2, eps_re, eps_im):
    """CplxLinearGaussian.forward (training), nn/relevance/complex/base.py:43-56,
    with the noise ``cplx.randn_like(s2)`` passed in as (eps_re, eps_im)."""
    mu_re, mu_im = cplx_linear(x_re, x_im, w_re, w_im, b_re, b_im)
    s2 = F.linear(x_re * x_re + x_im * x_im, torch.exp(log_sigma2), None)
    sd = torch.sqrt(torch.clamp(s2, 1e-8))
    return mu_re + eps_re * sd, mu_im + eps_im * sd


def real_linear_vd(x, w, b, log_sigma2, eps):
    """LinearGaussian.forward (training), nn/relevance/real/base.py:43-49."""
    mu = F.linear(x, w, b)
    s2 = F.linear(x * x, torch.exp(log_sigma2), None)
    return mu + eps * torch.sqrt(torch.clamp(s2, 1e-8))


# -------------------------------------------------------------------------- bilinear
def cplx_bilinear(x1_re, x1_im, x2_re, x2_im, w_re, w_im, b_re=None, b_im=None, conjugate=True):
    """cplx.bilinear == bilinear_naive, cplxmodule/cplx.py:1062-1090."""
    n_out = int(w_re.shape[0])
    ww = torch.cat([w_re, w_im], dim=0)
    au, av = F.bilinear(x1_re, x2_re, ww, bias=None), F.bilinear(x1_re, x2_im, ww, bias=None)
    bu, bv = F.bilinear(x1_im, x2_re, ww, bias=None), F.bilinear(x1_im, x2_im, ww, bias=None)
    if conjugate:
        pp, qq = au + bv, av - bu
    else:
        pp, qq = au - bv, av + bu
    re = pp[..., :n_out] - qq[..., n_out:]
    im = pp[..., n_out:] + qq[..., :n_out]
    if b_re is not None:
        re, im = re + b_re, im + b_im
    return re, im


def cplx_bilinear_vd(x1_re, x1_im, x2_re, x2_im, w_re, w_im, b_re, b_im, log_sigma2, eps_re, eps_im,
                     conjugate=True):
    """CplxBilinearGaussian.forward (training), nn/relevance/complex/base.py:70-84."""
    mu_re, mu_im = cplx_bilinear(x1_re, x1_im, x2_re, x2_im, w_re, w_im, b_re, b_im, conjugate)
    s2 = F.bilinear(x1_re * x1_re + x1_im * x1_im, x2_re * x2_re + x2_im * x2_im,
                    torch.exp(log_sigma2), None)
    sd = torch.sqrt(torch.clamp(s2, 1e-8))
    return mu_re + eps_re * sd, mu_im + eps_im * sd


def real_bilinear_vd(x1, x2, w, b, log_sigma2, eps):
    """BilinearGaussian.forward (training), nn/relevance/real/base.py:66-80; eps=None: the mean."""
    mu = F.bilinear(x1, x2, w, b)
    if eps is None:
        return mu
    s2 = F.bilinear(x1 * x1, x2 * x2, torch.exp(log_sigma2), None)
    return mu + eps * torch.sqrt(torch.clamp(s2, 1e-8))


# ------------------------------------------------------------------------------ conv
def cplx_conv2d(x_re, x_im, w_re, w_im, b_re=None, b_im=None, stride=1, padding=0, dilation=1):
    """cplx.conv2d -> convnd -> convnd_quick (groups == 1), cplxmodule/cplx.py:729-742,770-838."""
    n_out = w_re.shape[0]
    ww = torch.cat([w_re, w_im], dim=0)
    wr = F.conv2d(x_re, ww, None, stride, padding, dilation, 1)
    wi = F.conv2d(x_im, ww, None, stride, padding, dilation, 1)
    re = wr[:, :n_out] - wi[:, n_out:]
    im = wr[:, n_out:] + wi[:, :n_out]
    if b_re is not None:
        re, im = re + b_re.reshape(-1, 1, 1), im + b_im.reshape(-1, 1, 1)
    return re, im


def cplx_conv2d_grouped(x_re, x_im, w_re, w_im, b_re=None, b_im=None, stride=1, padding=0, dilation=1,
                        groups=1):
    """cplx.conv2d -> convnd -> convnd_naive (groups > 1), cplxmodule/cplx.py:717-726,790-800."""
    conv = lambda a, w: F.conv2d(a, w, None, stride, padding, dilation, groups)
    re = conv(x_re, w_re) - conv(x_im, w_im)
    im = conv(x_re, w_im) + conv(x_im, w_re)
    if b_re is not None:
        re, im = re + b_re.reshape(-1, 1, 1), im + b_im.reshape(-1, 1, 1)
    return re, im


def cplx_conv2d_vd(x_re, x_im, w_re, w_im, b_re, b_im, log_sigma2, eps_re, eps_im, stride=1,
                   padding=0, dilation=1, groups=1):
    """CplxConvNdGaussianMixin._forward_impl, nn/relevance/complex/base.py:120-135."""
    if groups == 1:
        mu_re, mu_im = cplx_conv2d(x_re, x_im, w_re, w_im, b_re, b_im, stride, padding, dilation)
    else:
        mu_re, mu_im = cplx_conv2d_grouped(x_re, x_im, w_re, w_im, b_re, b_im, stride, padding,
                                           dilation, groups)
    s2 = F.conv2d(x_re * x_re + x_im * x_im, torch.exp(log_sigma2), None, stride, padding,
                  dilation, groups)
    sd = torch.sqrt(torch.clamp(s2, 1e-8))
    return mu_re + eps_re * sd, mu_im + eps_im * sd


def real_conv2d_vd(x, w, b, log_sigma2, eps, stride=1, padding=0, dilation=1, groups=1):
    """ConvNdGaussianMixin._forward_impl with F.conv2d, nn/relevance/real/base.py:149-163
    (training mode; ``eps`` is the ``torch.randn_like(s2)`` draw).  ``eps=None``: the mean only,
    i.e. the eval-mode forward (:150-152)."""
    mu = F.conv2d(x, w, b, stride, padding, dilation, groups)
    if eps is None:
        return mu
    s2 = F.conv2d(x * x, torch.exp(log_sigma2), None, stride, padding, dilation, groups)
    return mu + eps * torch.sqrt(torch.clamp(s2, 1e-8))


def real_conv1d_vd(x, w, b, log_sigma2, eps, stride=1, padding=0, dilation=1, groups=1):
    """Same with F.conv1d (Conv1dGaussian.forward, nn/relevance/real/base.py:166-177)."""
    mu = F.conv1d(x, w, b, stride, padding, dilation, groups)
    if eps is None:
        return mu
    s2 = F.conv1d(x * x, torch.exp(log_sigma2), None, stride, padding, dilation, groups)
    return mu + eps * torch.sqrt(torch.clamp(s2, 1e-8))


# -------------------------------------------------------------------------------- KL
def log_alpha_real(w, log_sigma2):
    """GaussianMixin.log_alpha, nn/relevance/real/base.py:23-26."""
    return log_sigma2 - 2 * torch.log(abs(w) + 1e-12)


def log_alpha_cplx(w_re, w_im, log_sigma2):
    """GaussianMixin.log_alpha, nn/relevance/complex/base.py:27-31 with
    abs(Cplx) = torch.norm(stack([re, im]), p=2, dim=0), cplx.py:183-192."""
    modulus = torch.norm(torch.stack([w_re, w_im], dim=0), p=2, dim=0, keepdim=False)
    return log_sigma2 - 2 * torch.log(modulus + 1e-12)


class _Expi(torch.autograd.Function):
    """ExpiFunction, nn/relevance/complex/vd.py:15-44: forward = host scipy in the input dtype
    (:31-36), backward = grad * exp(x) / x (:38-41)."""

    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        x_np = x.detach().cpu().numpy()
        return torch.from_numpy(scipy.special.expi(x_np, dtype=x_np.dtype)).to(x.device)   # vd.py:36

    @staticmethod
    def backward(ctx, grad_output):
        x, = ctx.saved_tensors
        return grad_output * torch.exp(x) / x


def expi(x):
    return _Expi.apply(x)


def penalty_real_vd(log_alpha):
    """RealVDMixin.penalty, nn/relevance/real/vd.py:74-76."""
    n = -log_alpha
    return F.softplus(n) / 2 + 0.63576 * torch.sigmoid(1.48695 * n - 1.87320)


def penalty_real_ard(log_alpha):
    """RealARDMixin.penalty, nn/relevance/real/ard.py:39."""
    return 0.5 * F.softplus(-log_alpha)


def penalty_cplx_vd(log_alpha):
    """CplxVDMixin.penalty, nn/relevance/complex/vd.py:95-99 (in the dtype of log_alpha,
    i.e. with the reference's fp32 cancellation when log_alpha is fp32)."""
    n = -log_alpha
    return EULER_GAMMA + n - expi(-torch.exp(n))


def penalty_cplx_ard(log_alpha):
    """CplxARDMixin.penalty, nn/relevance/complex/ard.py:39."""
    return F.softplus(-log_alpha)


def penalty_cplx_vd_exact64(log_alpha):
    """Float64 closed form of the same quantity: Ein(t) = gamma + ln t + E1(t), t = exp(-la),
    evaluated without cancellation (power series for t <= 1)."""
    la = log_alpha.detach().cpu().double().numpy()
    t = np.exp(-la)
    out = np.empty_like(t)
    small = t <= 1.0
    ts = t[small]
    acc = np.zeros_like(ts)
    term = np.ones_like(ts)
    for k in range(1, 40):
        term = term * (-ts) / k          # (-t)^k / k!
        acc = acc - term / k             # sum (-1)^(k+1) t^k / (k k!)
    out[small] = acc
    tl = t[~small]
    out[~small] = EULER_GAMMA + np.log(tl) + scipy.special.exp1(tl)
    return torch.from_numpy(out)


def penalty_cplx_vd_approx(log_alpha):
    """CplxVDApproxMixin.penalty, nn/relevance/extensions/complex.py:97-99."""
    n = -log_alpha
    return F.softplus(n) + 0.57810 * torch.sigmoid(1.36526 * n - 1.45926)


def penalty_cplx_vd_scalefree(w_re, w_im, log_sigma2):
    """CplxVDScaleFreeMixin.penalty, nn/relevance/extensions/complex.py:40-43 (needs the
    parameters, not just log_alpha)."""
    log_abs_w = torch.log(torch.norm(torch.stack([w_re, w_im], dim=0), p=2, dim=0) + 1e-12)
    n_log_alpha = 2 * log_abs_w - log_sigma2
    return log_abs_w - log_sigma2 - 0.5 * expi(-torch.exp(n_log_alpha))


PENALTY = {
    "real_vd": penalty_real_vd,
    "real_ard": penalty_real_ard,
    "cplx_vd": penalty_cplx_vd,
    "cplx_ard": penalty_cplx_ard,
}


def layer_penalty(kind, w_re, w_im, log_sigma2, reduction="sum"):
    """named_penalties body, nn/relevance/base.py:132-141."""
    if kind == "cplx_vd_scalefree":
        p = penalty_cplx_vd_scalefree(w_re, w_im, log_sigma2)
    else:
        la = log_alpha_cplx(w_re, w_im, log_sigma2) if kind.startswith("cplx") else \
            log_alpha_real(w_re, log_sigma2)
        p = (penalty_cplx_vd_approx if kind == "cplx_vd_approx" else PENALTY[kind])(la)
    if reduction == "sum":
        return p.sum()
    if reduction == "mean":
        return p.mean()
    return p


# ------------------------------------------------------- the headline step, CPU side
def cplx_linear_vd_step(x_re, x_im, w_re, w_im, b_re, b_im, log_sigma2, generator=None):
    """One unit of BASELINE.json's metric on CPU: training-mode CplxLinearVD forward
    (noise drawn like cplx.randn_like) + sum(penalties(model)), as in
    tests/test_relevance.py:65-68."""
    M, N = x_re.shape[0], w_re.shape[0]
    eps_re, eps_im = cplx_randn(M, N, dtype=x_re.dtype, generator=generator)
    y = cplx_linear_vd(x_re, x_im, w_re, w_im, b_re, b_im, log_sigma2, eps_re, eps_im)
    kl = layer_penalty("cplx_vd", w_re, w_im, log_sigma2, "sum")
    return y, kl
