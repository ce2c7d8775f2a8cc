"""Host side of the complex convolution path (reference: ``cplx.convnd``,
``cplxmodule/cplx.py:770-800`` and ``CplxConvNdGaussianMixin._forward_impl``,
``nn/relevance/complex/base.py:120-135``).  One C-ABI call per layer call, groups included."""
import numpy as np
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
    """Autograd seam of the conv path.  The backward runs on the same CUDA kernels:
    * dx = g (*) conj(W): the forward conv kernel on the (zero-dilated, for stride > 1) output
      gradient with the spatially flipped, channel-swapped, conjugated kernel;
    * dW = g^T . conj(im2col(x)): a GEMM through ``ops._gemm`` (tcgen05 when the shape allows) on
      an ``unfold`` of the input (data movement only), accumulated over batch chunks;
    * db = column sums of g; variational part as for the linear layer (``ops._vd_backward_extra``):
      s2 is recomputed by the conv kernel, the forward's noise regenerated from its Philox
      coordinates.
    The reference obtains all of this from torch autograd over ``cplx.py:729-742`` and
    ``nn/relevance/complex/base.py:120-135``."""

    @staticmethod
    def forward(ctx, x_re, x_im, w_re, w_im, b_re, b_im, ls2, eps_re, eps_im, noise, geom, groups=1):
        y_re, y_im, aux = _conv2d_raw(x_re, x_im, w_re, w_im, b_re, b_im, ls2, eps_re, eps_im, noise,
                                      geom, groups)
        if any(ctx.needs_input_grad):
            dt = w_re.dtype
            ctx.geom, ctx.noise, ctx.philox = geom, aux.get("noise", noise), aux.get("philox")
            ctx.groups = groups
            ctx.planes = (x_re.to(dt), None if x_im is None else x_im.to(dt), w_re, w_im,
                          None if ls2 is None else ls2.to(dt), aux.get("eps_re"), aux.get("eps_im"))
            ctx.x_dtype = x_re.dtype
            ctx.has_bias = b_re is not None
        return y_re, y_im

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_re, g_im):
        x_re, x_im, w_re, w_im, ls2, eps_re, eps_im = ctx.planes
        cplx = x_im is not None
        dt, dev = w_re.dtype, w_re.device
        need = ctx.needs_input_grad
        B, C, H, W = x_re.shape
        O, _, kh, kw = w_re.shape
        G = ctx.groups
        stride, padding, dilation = ctx.geom
        Ho = (H + 2 * padding[0] - dilation[0] * (kh - 1) - 1) // stride[0] + 1
        Wo = (W + 2 * padding[1] - dilation[1] * (kw - 1) - 1) // stride[1] + 1
        zeros = lambda: torch.zeros((B, O, Ho, Wo), dtype=dt, device=dev)
        g_re = zeros() if g_re is None else g_re.to(dt).contiguous()
        g_im = (zeros() if g_im is None else g_im.to(dt).contiguous()) if cplx else None

        dx_re = dx_im = dw_re = dw_im = db_re = db_im = dls2 = None
        if need[0] or (cplx and need[1]):
            dx_re, dx_im = _conv_dgrad(g_re, g_im, w_re, w_im, (H, W), ctx.geom, G)
        if need[2] or (cplx and need[3]):
            dw_re, dw_im = _conv_wgrad_grouped(g_re, g_im, x_re, x_im, (kh, kw), ctx.geom, G)
        if ctx.has_bias and (need[4] or (cplx and need[5])):
            rows = lambda g: g.permute(0, 2, 3, 1).reshape(-1, O).contiguous()
            db_re = ops._colsum(rows(g_re))
            db_im = ops._colsum(rows(g_im)) if cplx else None

        if ls2 is not None and (need[0] or need[1] or need[6]):
            if ctx.noise == nv.NOISE_PHILOX_FAST:
                raise NotImplementedError(
                    "backward of the variational conv with the private 'fast' noise layout is not "
                    "supported; use the torch-exact layout (default) or pass eps")
            E = ops._eltwise(ops.TR_EXP, ls2)
            q = ops._eltwise(ops.TR_ABS2, x_re, x_im) if cplx else ops._eltwise(ops.TR_SQR, x_re)
            s2, _, _ = _conv2d_raw(q, None, E, None, None, None, None, None, None, nv.NOISE_INJECT,
                                   ctx.geom, G)
            gs2 = torch.empty_like(s2)
            seed, offset, threads = ctx.philox or (0, 0, 0)
            with nv.device_guard(dev):
                nv.check(nv.lib().cplxk_vd_grad_s2(
                    nv.ptr(g_re), nv.ptr(g_im), nv.ptr(s2), nv.ptr(eps_re), nv.ptr(eps_im), ctx.noise,
                    seed, offset, threads, nv.ptr(gs2), s2.numel() // Wo, Wo, nv.dtype_code(dt),
                    nv.stream_ptr(dev)))
            if need[0] or need[1]:
                dq, _ = _conv_dgrad(gs2, None, E, None, (H, W), ctx.geom, G)
                with nv.device_guard(dev):
                    nv.check(nv.lib().cplxk_vd_grad_input(
                        nv.ptr(dx_re), nv.ptr(dx_im), nv.ptr(x_re.contiguous()),
                        nv.ptr(None if x_im is None else x_im.contiguous()), nv.ptr(dq), dq.numel(),
                        nv.dtype_code(dt), nv.stream_ptr(dev)))
            if need[6]:
                dE, _ = _conv_wgrad_grouped(gs2, None, q, None, (kh, kw), ctx.geom, G)
                dls2 = torch.empty_like(dE)
                with nv.device_guard(dev):
                    nv.check(nv.lib().cplxk_mul_exp(nv.ptr(dE), nv.ptr(ls2.contiguous()), nv.ptr(dls2),
                                                    dE.numel(), nv.dtype_code(dt), 0, nv.stream_ptr(dev)))
        cast = lambda t: None if t is None else (t if t.dtype == ctx.x_dtype else t.to(ctx.x_dtype))
        return (cast(dx_re), cast(dx_im), dw_re, dw_im, db_re, db_im, dls2, None, None, None, None, None)


def _conv_dgrad(g_re, g_im, w_re, w_im, in_hw, geom, groups=1):
    """dx = g (*) conj(W) as a forward conv: zero-dilate g by the stride, pad by d(k-1) - p,
    correlate with W'[c, o, r, s] = conj(W[o, c, kh-1-r, kw-1-s]) at the forward's dilation
    (grouped: the same inside every group, still one launch)."""
    stride, padding, dilation = geom
    H, W = in_hw
    B, O, Ho, Wo = g_re.shape
    _, Cg, kh, kw = w_re.shape
    ph, pw = dilation[0] * (kh - 1) - padding[0], dilation[1] * (kw - 1) - padding[1]
    if ph < 0 or pw < 0:
        raise NotImplementedError("conv backward needs padding <= dilation * (kernel_size - 1)")

    # the forward drops (H + 2p - d(k-1) - 1) mod s trailing positions from its last window
    # start, but input rows up to the end of that window still receive gradient: extend the
    # dilated gradient by that remainder (zeros) instead of padding asymmetrically
    gh, gw = (Ho - 1) * stride[0] + 1, (Wo - 1) * stride[1] + 1
    eh = H - (gh + dilation[0] * (kh - 1) - 2 * padding[0])
    ew = W - (gw + dilation[1] * (kw - 1) - 2 * padding[1])

    def dilate(g):
        if g is None or (stride == (1, 1) and eh == 0 and ew == 0):
            return g
        out = g.new_zeros((B, O, gh + max(eh, 0), gw + max(ew, 0)))
        out[:, :, :gh:stride[0], :gw:stride[1]] = g      # scatter (data movement)
        return out

    def flip(w):       # [G*Og, Cg, kh, kw] -> [G*Cg, Og, kh, kw], taps reversed
        if w is None:
            return None
        w = w.flip(2, 3).reshape(groups, O // groups, Cg, kh, kw).transpose(1, 2)
        return w.reshape(groups * Cg, O // groups, kh, kw).contiguous()
    wr, wi = flip(w_re), flip(w_im)
    if wi is not None:
        wi = ops._eltwise(ops.TR_NEG, wi)                 # conj
    dx_re, dx_im, _ = _conv2d_raw(dilate(g_re), dilate(g_im), wr, wi, None, None, None, None, None,
                                  nv.NOISE_INJECT, ((1, 1), (ph, pw), dilation), groups)

    def fit(t):      # (defensive) crop to the input size
        if t is None or tuple(t.shape[2:]) == (H, W):
            return t
        return t[:, :, :H, :W].contiguous()
    return fit(dx_re), fit(dx_im)


def _conv_wgrad(g_re, g_im, x_re, x_im, khw, geom, max_bytes=1 << 29):
    """dW[o, (c, r, s)] = sum over (b, pixel) g[b, o, pixel] conj(x_patch[b, (c, r, s), pixel]): one
    K-major GEMM per batch chunk on an unfold (im2col gather) of the input planes."""
    stride, padding, dilation = geom
    B, O, Ho, Wo = g_re.shape
    C = x_re.shape[1]
    kh, kw = khw
    L, ck = Ho * Wo, C * kh * kw
    per_image = ck * L * x_re.element_size() * (2 if x_im is not None else 1)
    chunk = max(1, min(B, max_bytes // max(per_image, 1)))
    dw_re = dw_im = None
    unfold = lambda t: F.unfold(t, (kh, kw), dilation=dilation, padding=padding, stride=stride)
    for b0 in range(0, B, chunk):
        sl = slice(b0, min(B, b0 + chunk))
        nb = sl.stop - sl.start
        cols = lambda t: unfold(t[sl]).permute(1, 0, 2).reshape(ck, nb * L).contiguous()
        rows = lambda g: g[sl].reshape(nb, O, L).permute(1, 0, 2).reshape(O, nb * L).contiguous()
        p_re = cols(x_re)
        p_im = None if x_im is None else ops._eltwise(ops.TR_NEG, cols(x_im))     # conj
        re, im = ops._gemm(rows(g_re), None if g_im is None else rows(g_im), p_re, p_im)
        if dw_re is None:
            dw_re, dw_im = re, im
        else:
            dw_re.add_(re)
            if im is not None:
                dw_im.add_(im)
    shape = (O, C, kh, kw)
    return dw_re.reshape(shape), None if dw_im is None else dw_im.reshape(shape)


def _conv_wgrad_grouped(g_re, g_im, x_re, x_im, khw, geom, groups):
    """Grouped layers: the weight gradient of a group only sees its own channel slices -- one
    GEMM (chain) per group."""
    if groups == 1:
        return _conv_wgrad(g_re, g_im, x_re, x_im, khw, geom)
    cin, cout = x_re.shape[1] // groups, g_re.shape[1] // groups
    re, im = [], []
    for gi in range(groups):
        ci, co = slice(gi * cin, (gi + 1) * cin), slice(gi * cout, (gi + 1) * cout)
        a, b = _conv_wgrad(g_re[:, co].contiguous(), None if g_im is None else g_im[:, co].contiguous(),
                           x_re[:, ci].contiguous(), None if x_im is None else x_im[:, ci].contiguous(),
                           khw, geom)
        re.append(a), im.append(b)
    return torch.cat(re, 0), None if g_im is None else torch.cat(im, 0)


def _conv2d_raw_per_group(x_re, x_im, w_re, w_im, b_re, b_im, ls2, eps_re, eps_im, noise, geom, groups):
    """One C-ABI call per group: only for what the single-launch grouped kernel cannot take
    (exact-fp32 'simt' math mode, geometries outside the TMA box limits).  The layer's noise is
    still ONE draw (torch layout), injected slice by slice."""
    cplx = x_im is not None
    aux = {}
    if ls2 is not None and noise != nv.NOISE_INJECT:
        Ho, Wo = _out_hw(x_re, w_re, geom)
        eps_re, eps_im = _draw_noise(cplx, (x_re.shape[0], w_re.shape[0], Ho, Wo), x_re.device, w_re.dtype)
        noise = nv.NOISE_INJECT
    cin, cout = x_re.shape[1] // groups, w_re.shape[0] // groups
    sl = lambda t, s, d=0: None if t is None else (t[s] if d == 0 else t[:, s])
    re, im = [], []
    for gi in range(groups):
        ci, co = slice(gi * cin, (gi + 1) * cin), slice(gi * cout, (gi + 1) * cout)
        a, b, _ = _conv2d_raw(sl(x_re, ci, 1), sl(x_im, ci, 1), sl(w_re, co), sl(w_im, co), sl(b_re, co),
                              sl(b_im, co), sl(ls2, co), sl(eps_re, co, 1), sl(eps_im, co, 1), noise, geom)
        re.append(a), im.append(b)
    aux.update(noise=noise, philox=(0, 0, 0), eps_re=eps_re, eps_im=eps_im)
    return torch.cat(re, 1), torch.cat(im, 1) if cplx else None, aux


def _conv2d_raw(x_re, x_im, w_re, w_im, b_re, b_im, ls2, eps_re, eps_im, noise, geom, groups=1):
    stride, padding, dilation = geom
    dev = nv.require_cuda(x_re, x_im, w_re, w_im, b_re, b_im, ls2, eps_re, eps_im)
    dt = w_re.dtype
    code = nv.dtype_code(dt)
    cplx = x_im is not None
    B, C, H, W = x_re.shape
    O, Cw, kh, kw = w_re.shape
    if groups < 1 or C % groups or O % groups:
        raise ValueError("in_channels and out_channels must be divisible by groups")
    if Cw * groups != C:
        raise RuntimeError(f"expected input with {Cw * groups} channels, got {C}")
    Ho = (H + 2 * padding[0] - dilation[0] * (kh - 1) - 1) // stride[0] + 1
    Wo = (W + 2 * padding[1] - dilation[1] * (kw - 1) - 1) // stride[1] + 1
    if Ho <= 0 or Wo <= 0:
        raise RuntimeError("kernel size can't be greater than actual input size")
    math = ops._MATH[ops.get_math_mode()]
    # torch.channels_last activations (NCHW shape, NHWC strides) are consumed in place by the
    # tensor-core path and the output keeps that memory format, as F.conv2d would
    gran = 8 if dt == torch.float32 else 16
    cl = (cplx and groups == 1 and math != nv.MATH_SIMT and C % gran == 0 and C > 1
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
    if y_re.numel() == 0:      # empty batch / no output channels: empty outputs, as F.conv2d gives
        return y_re, y_im, {"philox": (0, 0, 0), "eps_re": None, "eps_im": None}
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
    if ls2 is not None and math != nv.MATH_SIMT and not cl and ops.conv_vd_composed(cplx, dt):
        # mean conv + variance conv on the fast plain kernels, then ONE in-place launch for the
        # noise: y += eps sqrt(max(s2, 1e-8))  (ops.set_conv_vd_mode)
        y_re = y_im = None
        y_re, y_im, _ = _conv2d_raw(xr, xi, wr, wi, br, bi, None, None, None, nv.NOISE_INJECT, geom, groups)
        E = ops._eltwise(ops.TR_EXP, l2)
        s2 = _variance_conv2d(xr, xi, E, geom, groups) if cplx else None
        if s2 is None:
            q = ops._eltwise(ops.TR_ABS2, xr, xi) if cplx else ops._eltwise(ops.TR_SQR, xr)
            s2, _, _ = _conv2d_raw(q, None, E, None, None, None, None, None, None, nv.NOISE_INJECT, geom, groups)
            del q
        with nv.device_guard(dev):
            nv.check(nv.lib().cplxk_vd_combine(
                nv.ptr(y_re), nv.ptr(y_im), nv.ptr(s2), nv.ptr(er), nv.ptr(ei), mode, seed, offset,
                threads, y_re.numel(), code, nv.stream_ptr(dev)))
        if mode != nv.NOISE_INJECT:
            nv.philox_advance(gen, offset, inc)
        return y_re, y_im, {"philox": (seed, offset, threads), "eps_re": er, "eps_im": ei}
    ws, ws_bytes = None, 0
    if math != nv.MATH_SIMT:
        # complex AND real planes run the tcgen05 implicit GEMM (real: one A tile, one accumulator)
        ws_bytes = nv.lib().cplxk_conv2d_workspace_bytes_g(B, C, H, W, O, kh, kw, groups, 1 if cplx else 0,
                                                           code, 1 if ls2 is not None else 0)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    def call(xr, xi, y_re, y_im, cl):
        with nv.device_guard(dev):
            return nv.lib().cplxk_conv2d_fwd_g(
                nv.ptr(xr), nv.ptr(xi), nv.ptr(wr), nv.ptr(wi), nv.ptr(br), nv.ptr(bi), nv.ptr(l2),
                nv.ptr(er), nv.ptr(ei), mode, seed, offset, threads, nv.ptr(y_re), nv.ptr(y_im),
                B, C, H, W, O, kh, kw, stride[0], stride[1], padding[0], padding[1],
                dilation[0], dilation[1], groups, code, math, 1 if cl else 0, nv.ptr(ws), ws_bytes,
                nv.stream_ptr(dev))

    rc = call(xr, xi, y_re, y_im, cl)
    if cl and rc == nv.ERR_UNSUPPORTED:
        # geometry outside the implicit-GEMM kernel's TMA box limits: the other CUDA kernels
        # take NCHW planes, so re-lay the activations once (still no CPU path)
        xr, xi = nv.plane(x_re, dt), nv.plane(x_im, dt)
        y_re = torch.empty((B, O, Ho, Wo), dtype=dt, device=dev)
        y_im = torch.empty_like(y_re)
        rc = call(xr, xi, y_re, y_im, False)
    if groups > 1 and rc == nv.ERR_UNSUPPORTED:
        return _conv2d_raw_per_group(x_re, x_im, w_re, w_im, b_re, b_im, ls2, eps_re, eps_im, noise,
                                     geom, groups)
    nv.check(rc)
    if mode != nv.NOISE_INJECT:
        nv.philox_advance(gen, offset, inc)
    return y_re, y_im, {"philox": (seed, offset, threads), "eps_re": er, "eps_im": ei}


def _variance_conv2d(x_re, x_im, E, geom, groups):
    """s2 = conv(|x_re + i x_im|^2, E) in ONE C-ABI call: the real-plane tensor-core kernel whose
    transposing pre-pass squares the input on the way (no |x|^2 plane is written or re-read;
    ``cplxk_conv2d_fwd_g`` with an imaginary input plane but real weights and output).  Returns None
    where that form is not available (geometry / alignment), the caller then forms |x|^2 itself.
    Reference: ``F.conv(abs(input)**2, exp(log_sigma2))``, nn/relevance/complex/base.py:100-117."""
    stride, padding, dilation = geom
    dev, dt = x_re.device, E.dtype
    B, C, H, W = x_re.shape
    O, _, kh, kw = E.shape
    Ho, Wo = _out_hw(x_re, E, geom)
    if ops._MATH[ops.get_math_mode()] == nv.MATH_SIMT or min(B, O, Ho, Wo) <= 0:
        return None
    code = nv.dtype_code(dt)
    ws_bytes = nv.lib().cplxk_conv2d_workspace_bytes_g(B, C, H, W, O, kh, kw, groups, 0, code, 0)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    s2 = torch.empty((B, O, Ho, Wo), dtype=dt, device=dev)
    with nv.device_guard(dev):
        rc = nv.lib().cplxk_conv2d_fwd_g(
            nv.ptr(x_re), nv.ptr(x_im), nv.ptr(E), None, None, None, None, None, None, nv.NOISE_INJECT,
            0, 0, 0, nv.ptr(s2), None, B, C, H, W, O, kh, kw, stride[0], stride[1], padding[0],
            padding[1], dilation[0], dilation[1], groups, code, nv.MATH_TENSOR, 0, nv.ptr(ws), ws_bytes,
            nv.stream_ptr(dev))
    if rc in (nv.ERR_UNSUPPORTED, nv.ERR_WORKSPACE):
        return None
    nv.check(rc)
    return s2


def _out_hw(x, w, geom):
    stride, padding, dilation = geom
    H, W = x.shape[2], x.shape[3]
    kh, kw = w.shape[2], w.shape[3]
    return ((H + 2 * padding[0] - dilation[0] * (kh - 1) - 1) // stride[0] + 1,
            (W + 2 * padding[1] - dilation[1] * (kw - 1) - 1) // stride[1] + 1)


def _draw_noise(cplx, shape, device, dtype):
    """The draw the reference makes for a whole layer -- ``cplx.randn_like(s2)`` = ONE
    ``randn(2, *shape) / sqrt(2)`` (cplx.py:544-550) or ``torch.randn_like(s2)`` -- produced by
    the library's torch-layout Philox kernel (bit-equal to torch's stream for the same generator
    state; the generator is advanced like torch would).  Used where one launch cannot cover the
    layer's output (grouped convolutions): the groups then read their channel slices of it."""
    numel = (2 if cplx else 1) * int(torch.Size(shape).numel())
    gen, seed, offset, threads, inc = nv.philox_plan(device, max(numel, 1), True)
    # torch evaluates `randn(...) / sqrt(2)` on CUDA as a product with the fp32 reciprocal
    inv_sqrt2 = float(np.float32(1.0) / np.float32(2.0 ** 0.5))
    flat = ops.randn_philox_torch(numel, seed, offset, threads, inv_sqrt2 if cplx else 1.0, device)
    nv.philox_advance(gen, offset, inc)
    flat = flat.to(dtype)
    if cplx:
        planes = flat.view(2, *shape)
        return planes[0], planes[1]
    return flat.view(*shape), None


def _grad_wanted(*tensors):
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors)


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
    if (ls2 is not None and noise == nv.NOISE_PHILOX_FAST
            and _grad_wanted(x_re, x_im, w_re, w_im, ls2)):
        # the private 'fast' layout of the conv kernels is not regenerated by the backward: under
        # autograd the forward uses the torch-exact layout instead (same distribution)
        noise = nv.NOISE_PHILOX_TORCH
    er, ei = (None, None) if e is None else e
    re, im = _ConvFn.apply(x_re, x_im, w_re, w_im, b_re, b_im, ls2, er, ei, noise, geom, groups)
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
    if ls2 is not None and noise == nv.NOISE_PHILOX_FAST and _grad_wanted(x, w, ls2):
        noise = nv.NOISE_PHILOX_TORCH       # see cplx_convnd
    out = _ConvFn.apply(x, None, w, None, bias, None, ls2, e, None, noise, geom, groups)[0]
    return out.squeeze(2) if nd == 1 else out
