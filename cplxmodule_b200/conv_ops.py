"""Host side of the complex convolution path (reference: ``cplx.convnd``,
``cplxmodule/cplx.py:770-800`` and ``CplxConvNdGaussianMixin._forward_impl``,
``nn/relevance/complex/base.py:120-135``).  One C-ABI call per group."""
import torch
import torch.nn.functional as F

from . import _native as nv
from . import ops
from .cplx import Cplx, cat


def _tuple(v, n):
    if isinstance(v, int):
        return (v,) * n
    v = tuple(int(e) for e in v)
    if len(v) != n:
        raise ValueError(f"expected {n} values, got {v}")
    return v


def _circular_pad(input, padding):
    """symmetric_circular_padding, cplx.py:701-714: (pad+1)//2 in front, pad//2 behind."""
    expanded = []
    for pad in reversed(padding):  # F.pad lists the last dimension first
        expanded.extend(((pad + 1) // 2, pad // 2))
    return input.apply(F.pad, tuple(expanded), mode="circular")


class _ConvFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x_re, x_im, w_re, w_im, b_re, b_im, ls2, eps_re, eps_im, noise, geom):
        return _conv2d_raw(x_re, x_im, w_re, w_im, b_re, b_im, ls2, eps_re, eps_im, noise, geom)

    @staticmethod
    def backward(ctx, *grads):
        raise NotImplementedError("cplxmodule_b200: backward of the fused conv kernel is not "
                                  "implemented yet (no silent torch fallback).")


def _conv2d_raw(x_re, x_im, w_re, w_im, b_re, b_im, ls2, eps_re, eps_im, noise, geom):
    stride, padding, dilation = geom
    dev = nv.require_cuda(x_re, x_im, w_re, w_im, b_re, b_im, ls2, eps_re, eps_im)
    dt = w_re.dtype
    code = nv.dtype_code(dt)
    cplx = x_im is not None
    B, C, H, W = x_re.shape
    O, Cw, kh, kw = w_re.shape
    if Cw != C:
        raise RuntimeError(f"expected input with {Cw} channels, got {C}")
    Ho = (H + 2 * padding[0] - dilation[0] * (kh - 1) - 1) // stride[0] + 1
    Wo = (W + 2 * padding[1] - dilation[1] * (kw - 1) - 1) // stride[1] + 1
    if Ho <= 0 or Wo <= 0:
        raise RuntimeError("kernel size can't be greater than actual input size")
    math = nv.MATH_SIMT if ops.get_math_mode() == "simt" else (
        nv.MATH_TENSOR if ops.get_math_mode() == "tensor" else nv.MATH_AUTO)
    # torch.channels_last activations (NCHW shape, NHWC strides) are consumed in place by the
    # tensor-core path and the output keeps that memory format, as F.conv2d would
    gran = 8 if dt == torch.float32 else 16
    cl = (cplx and math != nv.MATH_SIMT and C % gran == 0 and C > 1
          and x_re.dtype == dt and x_im.dtype == dt
          and x_re.is_contiguous(memory_format=torch.channels_last) and not x_re.is_contiguous()
          and x_im.is_contiguous(memory_format=torch.channels_last))
    if cl:
        xr, xi = x_re, x_im
    else:
        xr, xi = nv.plane(x_re, dt), nv.plane(x_im, dt)
    wr, wi = nv.plane(w_re), nv.plane(w_im)
    br, bi = nv.plane(b_re, dt), nv.plane(b_im, dt)
    l2 = nv.plane(ls2, dt)
    fmt = torch.channels_last if cl else torch.contiguous_format
    y_re = torch.empty((B, O, Ho, Wo), dtype=dt, device=dev, memory_format=fmt)
    y_im = torch.empty_like(y_re) if cplx else None
    er = ei = None
    seed = offset = threads = 0
    gen = None
    mode = nv.NOISE_INJECT
    if ls2 is not None:
        mode = noise
        if noise == nv.NOISE_INJECT:
            er, ei = nv.plane(eps_re, dt), nv.plane(eps_im, dt)
        else:
            numel = (2 if cplx else 1) * y_re.numel()
            gen, seed, offset, threads, inc = nv.philox_plan(dev, max(numel, 1), noise == nv.NOISE_PHILOX_TORCH)
    ws, ws_bytes = None, 0
    if cplx and math != nv.MATH_SIMT:
        ws_bytes = nv.lib().cplxk_conv2d_workspace_bytes(B, C, H, W, O, kh, kw, code,
                                                         1 if ls2 is not None else 0)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    def call(xr, xi, y_re, y_im, cl):
        with torch.cuda.device(dev):
            return nv.lib().cplxk_conv2d_fwd(
                nv.ptr(xr), nv.ptr(xi), nv.ptr(wr), nv.ptr(wi), nv.ptr(br), nv.ptr(bi), nv.ptr(l2),
                nv.ptr(er), nv.ptr(ei), mode, seed, offset, threads, nv.ptr(y_re), nv.ptr(y_im),
                B, C, H, W, O, kh, kw, stride[0], stride[1], padding[0], padding[1],
                dilation[0], dilation[1], code, math, 1 if cl else 0, nv.ptr(ws), ws_bytes,
                nv.stream_ptr(dev))

    rc = call(xr, xi, y_re, y_im, cl)
    if cl and rc == nv.ERR_UNSUPPORTED:
        # geometry outside the implicit-GEMM kernel's TMA box limits: the other CUDA kernels
        # take NCHW planes, so re-lay the activations once (still no CPU path)
        xr, xi = nv.plane(x_re, dt), nv.plane(x_im, dt)
        y_re = torch.empty((B, O, Ho, Wo), dtype=dt, device=dev)
        y_im = torch.empty_like(y_re)
        rc = call(xr, xi, y_re, y_im, False)
    nv.check(rc)
    if gen is not None:
        gen.set_offset(offset + inc)
    return y_re, y_im


def cplx_convnd(nd, input, weight, bias=None, stride=1, padding=0, dilation=1, groups=1,
                padding_mode="zeros", log_sigma2=None, eps=None):
    stride, padding, dilation = _tuple(stride, nd), _tuple(padding, nd), _tuple(dilation, nd)
    if padding_mode == "circular":
        input = _circular_pad(input, padding)
        padding = (0,) * nd
    elif padding_mode != "zeros":
        raise ValueError("padding_mode must be 'zeros' or 'circular'.")
    if input.dim() != nd + 2:
        raise RuntimeError(f"expected a {nd + 2}-d complex input, got {input.dim()}-d")
    if nd == 1:  # conv1d == conv2d with a unit height
        lift = lambda t: None if t is None else t.unsqueeze(2)
        x_re, x_im = lift(input.real), lift(input.imag)
        w_re, w_im = lift(weight.real), lift(weight.imag)
        ls2 = lift(log_sigma2)
        e = None if eps is None else (lift(eps.real), lift(eps.imag))
        geom = ((1,) + stride, (0,) + padding, (1,) + dilation)
    else:
        x_re, x_im, w_re, w_im, ls2 = input.real, input.imag, weight.real, weight.imag, log_sigma2
        e = None if eps is None else (eps.real, eps.imag)
        geom = (stride, padding, dilation)
    b_re, b_im = (None, None) if bias is None else (bias.real, bias.imag)
    noise = nv.NOISE_INJECT if e is not None else ops._NOISE[ops.get_noise_mode()]

    def run(xr, xi, wr, wi, br, bi, l2, ee):
        er, ei = (None, None) if ee is None else ee
        return _ConvFn.apply(xr, xi, wr, wi, br, bi, l2, er, ei, noise, geom)

    if groups == 1:
        re, im = run(x_re, x_im, w_re, w_im, b_re, b_im, ls2, e)
    else:
        if ls2 is not None and e is None and noise == nv.NOISE_PHILOX_TORCH:
            raise NotImplementedError(
                "grouped variational conv with the torch-exact noise stream is not supported; "
                "use set_noise_mode('fast') or pass eps")
        cin, cout = x_re.shape[1] // groups, w_re.shape[0] // groups
        outs = []
        for gi in range(groups):
            ci, co = slice(gi * cin, (gi + 1) * cin), slice(gi * cout, (gi + 1) * cout)
            ee = None if e is None else (e[0][:, co], e[1][:, co])
            outs.append(Cplx(*run(x_re[:, ci], x_im[:, ci], w_re[co], w_im[co],
                                  None if b_re is None else b_re[co],
                                  None if b_im is None else b_im[co],
                                  None if ls2 is None else ls2[co], ee)))
        out = cat(outs, dim=1)
        re, im = out.real, out.imag
    if nd == 1:
        re, im = re.squeeze(2), im.squeeze(2)
    return Cplx(re, im)


def real_convnd(nd, input, weight, bias=None, stride=1, padding=0, dilation=1, groups=1,
                log_sigma2=None, eps=None):
    """Real-valued cross-correlation / its variational forward through the same C-ABI entry
    (reference: ``F.conv{1,2}d`` as used by ``ConvNdGaussianMixin._forward_impl``,
    ``nn/relevance/real/base.py:149-163``).  Zero padding only, like the reference layers."""
    stride, padding, dilation = _tuple(stride, nd), _tuple(padding, nd), _tuple(dilation, nd)
    if input.dim() != nd + 2:
        raise RuntimeError(f"expected a {nd + 2}-d input, got {input.dim()}-d")
    if nd == 1:
        lift = lambda t: None if t is None else t.unsqueeze(2)
        x, w, ls2, e = lift(input), lift(weight), lift(log_sigma2), lift(eps)
        geom = ((1,) + stride, (0,) + padding, (1,) + dilation)
    else:
        x, w, ls2, e = input, weight, log_sigma2, eps
        geom = (stride, padding, dilation)
    noise = nv.NOISE_INJECT if e is not None else ops._NOISE[ops.get_noise_mode()]

    def run(x_, w_, b_, l2_, e_):
        return _ConvFn.apply(x_, None, w_, None, b_, None, l2_, e_, None, noise, geom)[0]

    if groups == 1:
        out = run(x, w, bias, ls2, e)
    else:
        if ls2 is not None and e is None and noise == nv.NOISE_PHILOX_TORCH:
            raise NotImplementedError(
                "grouped variational conv with the torch-exact noise stream is not supported; "
                "use set_noise_mode('fast') or pass eps")
        cin, cout = x.shape[1] // groups, w.shape[0] // groups
        outs = []
        for gi in range(groups):
            ci, co = slice(gi * cin, (gi + 1) * cin), slice(gi * cout, (gi + 1) * cout)
            outs.append(run(x[:, ci], w[co], None if bias is None else bias[co],
                            None if ls2 is None else ls2[co], None if e is None else e[:, co]))
        out = torch.cat(outs, dim=1)
    return out.squeeze(2) if nd == 1 else out
