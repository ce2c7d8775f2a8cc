"""Tensor-level entry points of the hot path: every function here ends in exactly
one call through the C ABI (``_native``).  ``torch`` is used for device memory,
streams and the autograd seam only -- there is no torch/CPU implementation of
these ops in the product path.

The autograd seam follows the reference's only custom op, ``ExpiFunction``
(cplxmodule/nn/relevance/complex/vd.py:15-44).
"""
import ctypes

import torch

from . import _native as nv

_state = {"noise": "torch", "math": "auto", "prepare": True}
_MATH = {"auto": nv.MATH_AUTO, "tensor": nv.MATH_TENSOR, "simt": nv.MATH_SIMT}
_NOISE = {"torch": nv.NOISE_PHILOX_TORCH, "fast": nv.NOISE_PHILOX_FAST}


def set_noise_mode(mode):
    """``"torch"``: the fused epilogue regenerates bit-for-bit the Philox stream
    ``cplx.randn_like`` / ``torch.randn_like`` would draw on the CUDA device (so a
    seeded run reproduces the reference executed on the same GPU); ``"fast"``: a
    private counter layout that spends one Philox call per four normals."""
    if mode not in _NOISE:
        raise ValueError(f"noise mode must be one of {sorted(_NOISE)}")
    _state["noise"] = mode


def set_math_mode(mode):
    """``"auto"`` (tensor cores when alignment allows), ``"tensor"``, ``"simt"`` (exact fp32)."""
    if mode not in _MATH:
        raise ValueError(f"math mode must be one of {sorted(_MATH)}")
    _state["math"] = mode


def set_operand_prepass(enabled):
    """True (default): |x|^2 and exp(log_sigma2) are written once to a scratch workspace by an
    elementwise launch and streamed by TMA; False: produced inside the GEMM kernel's smem
    pipeline (single launch)."""
    _state["prepare"] = bool(enabled)


def get_noise_mode():
    return _state["noise"]


def get_math_mode():
    return _state["math"]


_BACKWARD_MSG = (
    "cplxmodule_b200: backward of the fused {} kernel is not implemented yet "
    "(forward + KL are the accelerated path; there is deliberately no silent torch fallback)."
)


def _flat2d(t, K):
    return None if t is None else nv.plane(t.reshape(-1, K))


def _forward_raw(x_re, x_im, w_re, w_im, b_re, b_im, log_sigma2, eps_re, eps_im, noise):
    """Shared launcher. Returns (y_re, y_im|None). ``noise`` is None for the plain map."""
    dev = nv.require_cuda(x_re, x_im, w_re, w_im, b_re, b_im, log_sigma2, eps_re, eps_im)
    cplx = x_im is not None
    dt = w_re.dtype
    code = nv.dtype_code(dt)
    N, K = w_re.shape
    if x_re.shape[-1] != K:
        raise RuntimeError(
            f"size mismatch: input has {x_re.shape[-1]} features, weight is {tuple(w_re.shape)}")
    lead = x_re.shape[:-1]
    xr = _flat2d(x_re.to(dt), K)
    xi = _flat2d(x_im.to(dt), K) if cplx else None
    M = xr.shape[0]
    wr, wi = nv.plane(w_re), nv.plane(w_im)
    br, bi = nv.plane(b_re, dt), nv.plane(b_im, dt)
    y_re = torch.empty((M, N), dtype=dt, device=dev)
    y_im = torch.empty((M, N), dtype=dt, device=dev) if cplx else None
    lib = nv.lib()
    math = _MATH[_state["math"]]
    with torch.cuda.device(dev):
        st = nv.stream_ptr(dev)
        if log_sigma2 is None:
            nv.check(lib.cplxk_linear_fwd(nv.ptr(xr), nv.ptr(xi), nv.ptr(wr), nv.ptr(wi),
                                          nv.ptr(br), nv.ptr(bi), nv.ptr(y_re), nv.ptr(y_im),
                                          M, N, K, code, math, st))
        else:
            ls2 = nv.plane(log_sigma2, dt)
            if noise == nv.NOISE_INJECT:
                er, ei = _flat2d(eps_re.to(dt), N), (_flat2d(eps_im.to(dt), N) if cplx else None)
                seed = offset = threads = 0
            else:
                er = ei = None
                numel = (2 if cplx else 1) * M * N
                gen, seed, offset, threads, inc = nv.philox_plan(dev, max(numel, 1))
            ws, ws_bytes = None, 0
            if _state["prepare"] and math != nv.MATH_SIMT:
                ws_bytes = lib.cplxk_linear_vd_workspace_bytes(M, N, K, code)
                ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            nv.check(lib.cplxk_linear_vd_fwd(nv.ptr(xr), nv.ptr(xi), nv.ptr(wr), nv.ptr(wi),
                                             nv.ptr(br), nv.ptr(bi), nv.ptr(ls2), nv.ptr(er),
                                             nv.ptr(ei), noise, seed, offset, threads,
                                             nv.ptr(y_re), nv.ptr(y_im), M, N, K, code, math,
                                             nv.ptr(ws), ws_bytes, st))
            if noise != nv.NOISE_INJECT:
                gen.set_offset(offset + inc)
    y_re = y_re.reshape(*lead, N)
    if cplx:
        y_im = y_im.reshape(*lead, N)
    return y_re, y_im


class _CplxLinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x_re, x_im, w_re, w_im, b_re, b_im):
        return _forward_raw(x_re, x_im, w_re, w_im, b_re, b_im, None, None, None, None)

    @staticmethod
    def backward(ctx, *grads):
        raise NotImplementedError(_BACKWARD_MSG.format("complex linear"))


class _RealLinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b):
        return _forward_raw(x, None, w, None, b, None, None, None, None, None)[0]

    @staticmethod
    def backward(ctx, *grads):
        raise NotImplementedError(_BACKWARD_MSG.format("linear"))


class _CplxLinearVDFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x_re, x_im, w_re, w_im, b_re, b_im, log_sigma2, eps_re, eps_im, noise):
        return _forward_raw(x_re, x_im, w_re, w_im, b_re, b_im, log_sigma2, eps_re, eps_im, noise)

    @staticmethod
    def backward(ctx, *grads):
        raise NotImplementedError(_BACKWARD_MSG.format("complex variational linear"))


class _RealLinearVDFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, log_sigma2, eps, noise):
        return _forward_raw(x, None, w, None, b, None, log_sigma2, eps, None, noise)[0]

    @staticmethod
    def backward(ctx, *grads):
        raise NotImplementedError(_BACKWARD_MSG.format("variational linear"))


def cplx_linear(x_re, x_im, w_re, w_im, b_re=None, b_im=None):
    """y = x W^T + b on split planes (reference: cplx.linear_naive, cplx.py:634-648)."""
    return _CplxLinearFn.apply(x_re, x_im, w_re, w_im, b_re, b_im)


def real_linear(x, w, b=None):
    return _RealLinearFn.apply(x, w, b)


def _noise_args(eps):
    if eps is None:
        return None, None, _NOISE[_state["noise"]]
    return eps[0], eps[1], nv.NOISE_INJECT


def cplx_linear_vd(x_re, x_im, w_re, w_im, b_re, b_im, log_sigma2, eps=None):
    """Fused local-reparameterisation forward (reference: CplxLinearGaussian.forward,
    nn/relevance/complex/base.py:43-56). ``eps=(eps_re, eps_im)`` injects the noise
    (each ~ N(0, 1/2)); ``None`` draws it inside the kernel."""
    er, ei, mode = _noise_args(eps)
    return _CplxLinearVDFn.apply(x_re, x_im, w_re, w_im, b_re, b_im, log_sigma2, er, ei, mode)


def real_linear_vd(x, w, b, log_sigma2, eps=None):
    """Reference: LinearGaussian.forward, nn/relevance/real/base.py:43-49."""
    er, _, mode = _noise_args((eps, None) if eps is not None else None)
    return _RealLinearVDFn.apply(x, w, b, log_sigma2, er, mode)


# ------------------------------------------------------------------------- KL path
def kl_penalty(kind, w_re, w_im, log_sigma2, reduction="sum"):
    """Penalty of one variational layer. ``reduction`` in {"sum", "mean", None}
    (reference: named_penalties, nn/relevance/base.py:88-141). The reduced variants
    never materialise the [N, K] penalty tensor."""
    dev = nv.require_cuda(w_re, w_im, log_sigma2)
    dt = w_re.dtype
    code = nv.dtype_code(dt)
    wr, wi, ls2 = nv.plane(w_re), nv.plane(w_im), nv.plane(log_sigma2, dt)
    n = wr.numel()
    lib = nv.lib()
    with torch.cuda.device(dev):
        st = nv.stream_ptr(dev)
        if reduction is None:
            out = torch.empty_like(wr)
            nv.check(lib.cplxk_kl(kind, nv.ptr(wr), nv.ptr(wi), nv.ptr(ls2), n, code, nv.ptr(out),
                                  None, 1.0, None, 0, st))
            return out
        if reduction not in ("sum", "mean"):
            raise ValueError(f"`reduction` must be either `None`, `sum` or `mean`. Got {reduction}.")
        ws = nv.kl_workspace(dev)
        out = torch.empty((), dtype=torch.float32, device=dev)
        scale = 1.0 if reduction == "sum" else 1.0 / max(n, 1)
        nv.check(lib.cplxk_kl(kind, nv.ptr(wr), nv.ptr(wi), nv.ptr(ls2), n, code, None,
                              nv.ptr(out), scale, nv.ptr(ws), ws.numel() * 8, st))
    return out.to(dt) if dt != torch.float32 else out


class _KLFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, kind, reduction, w_re, w_im, log_sigma2):
        return kl_penalty(kind, w_re, w_im, log_sigma2, reduction)

    @staticmethod
    def backward(ctx, *grads):
        raise NotImplementedError(_BACKWARD_MSG.format("KL"))


def kl(kind, w_re, w_im, log_sigma2, reduction="sum"):
    return _KLFn.apply(kind, reduction, w_re, w_im, log_sigma2)


def log_alpha(w_re, w_im, log_sigma2, threshold=None):
    """``log_sigma2 - 2 log(|w| + 1e-12)`` (nn/relevance/{real,complex}/base.py) or, with a
    ``threshold``, the relevance mask ``(log_alpha <= threshold)`` as floats."""
    dev = nv.require_cuda(w_re, w_im, log_sigma2)
    dt = w_re.dtype
    code = nv.dtype_code(dt)
    wr, wi, ls2 = nv.plane(w_re), nv.plane(w_im), nv.plane(log_sigma2, dt)
    out = torch.empty_like(wr)
    lib = nv.lib()
    with torch.cuda.device(dev):
        st = nv.stream_ptr(dev)
        if threshold is None:
            nv.check(lib.cplxk_log_alpha(nv.ptr(wr), nv.ptr(wi), nv.ptr(ls2), wr.numel(), code,
                                         nv.ptr(out), 0.0, None, st))
        else:
            nv.check(lib.cplxk_log_alpha(nv.ptr(wr), nv.ptr(wi), nv.ptr(ls2), wr.numel(), code,
                                         None, float(threshold), nv.ptr(out), st))
    return out


def randn_philox_torch(n, seed, offset, threads, scale=1.0, device="cuda"):
    """Test hook: the epilogue's torch-layout normal generator as a standalone fill."""
    device = torch.device(device)
    out = torch.empty(n, dtype=torch.float32, device=device)
    with torch.cuda.device(device):
        nv.check(nv.lib().cplxk_randn_philox_torch(nv.ptr(out), n, seed, offset, threads,
                                                   ctypes.c_float(scale), nv.stream_ptr(device)))
    return out
