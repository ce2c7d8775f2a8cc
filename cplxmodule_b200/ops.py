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

_state = {"noise": "torch", "math": "auto", "prepare": True, "fuse_kl": True, "kl_shard": None,
          "conv_vd": "auto"}
_MATH = {"auto": nv.MATH_AUTO, "tensor": nv.MATH_TENSOR, "simt": nv.MATH_SIMT, "exact": nv.MATH_SIMT,
         "tf32": nv.MATH_TENSOR_TF32}
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
    """Arithmetic of the GEMM-shaped part of every forward / gradient call on fp32 planes:

    * ``"auto"`` (default) / ``"tensor"``: tcgen05 tensor cores on per-row power-of-two scaled
      fp16 copies of the operands (tf32's 11-bit significand, round-to-nearest, fp32
      accumulation; about 3e-4 max-norm relative error; entries more than 2^-24 below their
      row maximum lose bits).  ``"auto"`` falls back to the exact kernel for shapes TMA cannot
      take, ``"tensor"`` raises instead;
    * ``"tf32"``: tensor cores on tf32 operands, i.e. what torch does under
      ``torch.backends.cuda.matmul.allow_tf32 = True`` (no per-row scale, half the MMA rate);
    * ``"simt"`` / ``"exact"``: fp32 FMA on the CUDA cores -- the arithmetic of the reference
      under torch's default ``allow_tf32 = False`` (about 1e-6).
    bf16 planes always run bf16 operands with fp32 accumulation."""
    if mode not in _MATH:
        raise ValueError(f"math mode must be one of {sorted(_MATH)}")
    _state["math"] = mode


def set_conv_vd_mode(mode):
    """Variational convolution forward on the tensor-core path (NCHW planes):

    * ``"composed"``: mean conv and variance conv through the fast plain kernels (CTA-pair /
      scaled-fp16 complex kernel, real-plane kernel) + ONE in-place launch that draws the noise by
      Philox call (three calls per four complex outputs) and adds ``eps * sqrt(max(s2, 1e-8))``;
    * ``"fused"``: the single kernel with three accumulators and the noise in its epilogue (no
      scratch planes; always used for channels-last inputs);
    * ``"auto"`` (default): composed for NCHW planes.  Measured on config-4-sized work
      (``tools/convvd_probe.py``, 256 x 64 x 128^2, 3 x 3, 64 -> 64, torch-exact noise, ms composed /
      fused, after the plain kernels' row mode): complex fp32 5.12 / 5.83, complex bf16 4.09 / 5.49,
      real fp32 3.36 / 4.74, real bf16 2.33 / 4.64.
    Same values either way (same noise stream, same tolerance)."""
    if mode not in ("auto", "composed", "fused"):
        raise ValueError("conv VD mode must be 'auto', 'composed' or 'fused'")
    _state["conv_vd"] = mode


def conv_vd_composed(cplx, dtype):
    mode = _state["conv_vd"]
    if mode == "auto":
        return True
    return mode == "composed"


def set_operand_prepass(enabled):
    """True (default): |x|^2 and exp(log_sigma2) are written once to a scratch workspace by an
    elementwise launch and streamed by TMA; False: produced inside the GEMM kernel's smem
    pipeline (single launch)."""
    _state["prepare"] = bool(enabled)


def set_kl_fusion(enabled):
    """True (default): a training-mode forward of a variational linear layer also produces the
    layer's KL sum (its operand pre-pass reads every weight anyway) and the next
    ``penalties(..., reduction="sum"|"mean")`` over unchanged parameters returns it instead of
    running the stand-alone KL pass.  ``"unchecked"``: the same without the device-side
    fingerprint check of ``FusedKLCache`` (two tiny launches per layer and step)."""
    _state["fuse_kl"] = "unchecked" if enabled == "unchecked" else bool(enabled)


def set_kl_shard(rank=None, world=None):
    """Multi-GPU: the KL by-product of a forward covers only this rank's block of weight rows
    (``distributed.row_shard``); ``distributed.sharded_penalties`` all-reduces the partial sums.
    ``set_kl_shard()`` (no arguments) returns to whole-layer sums."""
    _state["kl_shard"] = None if rank is None or world is None or world <= 1 else (int(rank), int(world))


def kl_request(kind, n_rows):
    """The ``kl_req`` a layer passes to its training forward (None: nothing to ask for)."""
    if kind is None or not _state["fuse_kl"]:
        return None
    req = {"kind": kind}
    if _state["kl_shard"] is not None:
        rank, world = _state["kl_shard"]
        base, extra = divmod(n_rows, world)
        lo = rank * base + min(rank, extra)
        req["rows"] = (lo, lo + base + (1 if rank < extra else 0))
    return req


def set_sm_reserve(n_sms):
    """Leave ``n_sms`` SMs out of the persistent GEMM grid so that kernels of other streams (a
    collective, a KL shard kernel) do not wait for its last tile (0: use every SM)."""
    nv.check(nv.lib().cplxk_set_sm_reserve(int(n_sms or 0)))


def get_noise_mode():
    return _state["noise"]


def get_math_mode():
    return _state["math"]


def _flat2d(t, K):
    """[..., K] -> dense [M, K]; M is spelled out so that empty batches reshape too"""
    if t is None:
        return None
    rows = 1
    for d in t.shape[:-1]:
        rows *= int(d)
    return nv.plane(t.reshape(rows, K))


def _check_features(x, w):
    if x.shape[-1] != w.shape[-1]:
        raise RuntimeError(
            f"size mismatch: input has {x.shape[-1]} features, weight is {tuple(w.shape)}")


def _forward_raw(x_re, x_im, w_re, w_im, b_re, b_im, log_sigma2, eps_re, eps_im, noise,
                 want_s2=False, math=None, kl_req=None):
    """Shared launcher. Returns (y_re, y_im|None, aux). ``noise`` is None for the plain map;
    ``aux`` = dict(s2=..., philox=(seed, offset, threads)) for the variational forward.
    ``kl_req`` = {"kind": k[, "rows": (lo, hi)]}: ask the operand pre-pass for the layer's KL sum
    (over the weight rows ``[lo, hi)``: a rank's shard) as a by-product; when the path taken
    produced it, ``kl_req["sum"]`` is set to the 0-d float32 result and ``kl_req["event"]`` to a
    CUDA event recorded right after the pre-pass (the sum is final there)."""
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
    aux = {}
    if M == 0 or N == 0:
        # empty batch / layer: empty outputs, as F.linear gives; nothing is drawn (torch's normal_
        # does not advance its generator for zero elements) and no KL by-product is produced
        if log_sigma2 is not None:
            aux = {"s2": torch.empty((M, N), dtype=dt, device=dev) if want_s2 else None,
                   "philox": (0, 0, 0), "eps": (None, None), "x": (xr, xi)}
        return y_re.reshape(*lead, N), (y_im.reshape(*lead, N) if cplx else None), aux
    lib = nv.lib()
    math = _MATH[_state["math"]] if math is None else math
    with nv.device_guard(dev):
        st = nv.stream_ptr(dev)
        if log_sigma2 is None:
            ws_bytes = 0
            if _state["prepare"] and math != nv.MATH_SIMT:
                ws_bytes = lib.cplxk_linear_workspace_bytes(M, N, K, code)
            if ws_bytes:
                ws = nv.workspace(dev, ws_bytes)
                nv.check(lib.cplxk_linear_fwd_ws(nv.ptr(xr), nv.ptr(xi), nv.ptr(wr), nv.ptr(wi),
                                                 nv.ptr(br), nv.ptr(bi), nv.ptr(y_re), nv.ptr(y_im),
                                                 M, N, K, code, math, nv.ptr(ws), ws_bytes, st))
            else:
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
                gen, seed, offset, threads, inc = nv.philox_plan(dev, max(numel, 1), noise == nv.NOISE_PHILOX_TORCH)
            ws, ws_bytes = None, 0
            if _state["prepare"] and math != nv.MATH_SIMT:
                ws_bytes = lib.cplxk_linear_vd_workspace_bytes(M, N, K, code)
                ws = nv.workspace(dev, ws_bytes)
            s2 = torch.empty((M, N), dtype=dt, device=dev) if want_s2 else None
            if (kl_req is not None and ws is not None and _state["fuse_kl"]
                    and lib.cplxk_linear_vd_fuses_kl(M, N, K, code, math)):
                # the pre-pass also fingerprints the parameters it reads (FusedKLCache compares at hand-out)
                fp = (torch.empty(1, dtype=torch.int64, device=dev)
                      if _state["fuse_kl"] != "unchecked" else None)
                kl_sum = torch.empty((), dtype=torch.float32, device=dev)
                kl_ws = nv.kl_workspace(dev)
                done = ctypes.c_int(0)
                lo, hi = kl_req.get("rows") or (0, -1)
                ev = None
                if "rows" in kl_req:     # multi-GPU shard: a collective on a side stream will wait for it
                    ev = torch.cuda.Event()
                    ev.record(torch.cuda.current_stream(dev))   # materialises the handle; re-recorded by the call
                nv.check(lib.cplxk_linear_vd_fwd_kl(
                    nv.ptr(xr), nv.ptr(xi), nv.ptr(wr), nv.ptr(wi), nv.ptr(br), nv.ptr(bi),
                    nv.ptr(ls2), nv.ptr(er), nv.ptr(ei), noise, seed, offset, threads,
                    nv.ptr(y_re), nv.ptr(y_im), M, N, K, code, math, nv.ptr(s2), nv.ptr(ws),
                    ws_bytes, kl_req["kind"], nv.ptr(kl_sum), nv.ptr(kl_ws), kl_ws.numel() * 8,
                    lo, hi, None if ev is None else ctypes.c_void_p(ev.cuda_event), nv.ptr(fp),
                    ctypes.byref(done), st))
                if done.value:
                    kl_req["sum"], kl_req["event"], kl_req["fp"] = kl_sum, ev, fp
            else:
                nv.check(lib.cplxk_linear_vd_fwd(nv.ptr(xr), nv.ptr(xi), nv.ptr(wr), nv.ptr(wi),
                                                 nv.ptr(br), nv.ptr(bi), nv.ptr(ls2), nv.ptr(er),
                                                 nv.ptr(ei), noise, seed, offset, threads,
                                                 nv.ptr(y_re), nv.ptr(y_im), M, N, K, code, math,
                                                 nv.ptr(s2), nv.ptr(ws), ws_bytes, st))
            if noise != nv.NOISE_INJECT:
                nv.philox_advance(gen, offset, inc)
            aux = {"s2": s2, "philox": (seed, offset, threads), "eps": (er, ei), "x": (xr, xi)}
    y_re = y_re.reshape(*lead, N)
    if cplx:
        y_im = y_im.reshape(*lead, N)
    return y_re, y_im, aux


def _masked_forward_raw(x_re, x_im, w_re, w_im, mask, b_re, b_im):
    """y = x (W * mask)^T + b in ONE C-ABI call (cplxk_linear_masked_fwd): on the fp32 tensor-core
    path the mask is applied inside the operand pre-pass, nothing shaped like W is written."""
    dev = nv.require_cuda(x_re, x_im, w_re, w_im, mask, b_re, b_im)
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
    mk = nv.plane(mask.expand(N, K), dt)
    br, bi = nv.plane(b_re, dt), nv.plane(b_im, dt)
    y_re = torch.empty((M, N), dtype=dt, device=dev)
    y_im = torch.empty((M, N), dtype=dt, device=dev) if cplx else None
    lib = nv.lib()
    math = _MATH[_state["math"]]
    with nv.device_guard(dev):
        ws_bytes = lib.cplxk_linear_masked_workspace_bytes(M, N, K, code)
        ws = nv.workspace(dev, ws_bytes)
        nv.check(lib.cplxk_linear_masked_fwd(nv.ptr(xr), nv.ptr(xi), nv.ptr(wr), nv.ptr(wi), nv.ptr(mk),
                                             nv.ptr(br), nv.ptr(bi), nv.ptr(y_re), nv.ptr(y_im),
                                             M, N, K, code, math, nv.ptr(ws), ws_bytes,
                                             nv.stream_ptr(dev)))
    return y_re.reshape(*lead, N), (y_im.reshape(*lead, N) if cplx else None)


def _guard(ctx, *tensors):
    """Hand the ORIGINAL input tensors to autograd's version tracking (``save_for_backward``):
    ``backward`` touches ``ctx.saved_tensors`` first, so an in-place update of an input or a
    parameter between forward and backward raises torch's usual error instead of silently
    differentiating changed values.  The planes the kernels read are views of these tensors
    whenever no dtype conversion or densification was needed (no extra memory)."""
    ctx.save_for_backward(*[t for t in tensors if isinstance(t, torch.Tensor)])


# ------------------------------------------------------------------ backward helpers
TR_COPY, TR_NEG, TR_EXP, TR_ABS2, TR_SQR, TR_MUL = 0, 1, 2, 3, 4, 5


def _transpose(t, op=TR_COPY, t2=None):
    """out[cols, rows] = op(t[rows, cols]) through the C ABI (makes a K-major GEMM operand)."""
    rows, cols = t.shape
    out = torch.empty((cols, rows), dtype=t.dtype, device=t.device)
    with torch.cuda.device(t.device):
        nv.check(nv.lib().cplxk_transpose2d(nv.ptr(t), nv.ptr(t2), nv.ptr(out), rows, cols,
                                            nv.dtype_code(t.dtype), op, nv.stream_ptr(t.device)))
    return out


def _eltwise(op, a, b=None):
    """op(a [, b]) elementwise through the C ABI (TR_NEG, TR_EXP, TR_SQR, TR_ABS2)."""
    a = a.contiguous()
    b = None if b is None else b.contiguous()
    out = torch.empty_like(a)
    with torch.cuda.device(a.device):
        nv.check(nv.lib().cplxk_eltwise(op, nv.ptr(a), nv.ptr(b), nv.ptr(out), a.numel(),
                                        nv.dtype_code(a.dtype), nv.stream_ptr(a.device)))
    return out


def _colsum(g):
    M, N = g.shape
    out = torch.empty(N, dtype=g.dtype, device=g.device)
    with torch.cuda.device(g.device):
        nv.check(nv.lib().cplxk_colsum(nv.ptr(g), nv.ptr(out), M, N, nv.dtype_code(g.dtype),
                                       nv.stream_ptr(g.device)))
    return out


def _gemm(a_re, a_im, p_re, p_im):
    """(a_re + i a_im) . (p_re + i p_im)^T  (or the real product) on the tensor cores; gradient
    GEMMs contract over whatever the batch size is, so shapes the TMA path cannot take fall
    through to the exact-fp32 kernel even when the user forced ``"tensor"`` for the forward."""
    math = {"simt": nv.MATH_SIMT, "exact": nv.MATH_SIMT, "tf32": nv.MATH_TENSOR_TF32}.get(
        _state["math"], nv.MATH_AUTO)
    re, im, _ = _forward_raw(a_re, a_im, p_re, p_im, None, None, None, None, None, None, math=math)
    return re, im


def _grad2d(g, like_lead, N, dt, dev):
    M = 1
    for s_ in like_lead:
        M *= s_
    if g is None:
        return torch.zeros((M, N), dtype=dt, device=dev)
    return nv.plane(g.reshape(M, N).to(dt))


def _linear_backward(ctx, g_re, g_im, need_x, need_w, need_b):
    """Gradients of y = x W^T + b for split-complex (or real) planes."""
    xr, xi, wr, wi = ctx.saved_planes
    cplx = xi is not None
    dx_re = dx_im = dw_re = dw_im = db_re = db_im = None
    if need_x:
        if cplx:  # dx = g . conj(W): P = (U^T, -V^T)
            dx_re, dx_im = _gemm(g_re, g_im, _transpose(wr), _transpose(wi, TR_NEG))
        else:
            dx_re, _ = _gemm(g_re, None, _transpose(wr), None)
    if need_w:
        if cplx:  # dW = g^T . conj(x)
            dw_re, dw_im = _gemm(_transpose(g_re), _transpose(g_im), _transpose(xr),
                                 _transpose(xi, TR_NEG))
        else:
            dw_re, _ = _gemm(_transpose(g_re), None, _transpose(xr), None)
    if need_b:
        db_re = _colsum(g_re)
        db_im = _colsum(g_im) if cplx else None
    return dx_re, dx_im, dw_re, dw_im, db_re, db_im


def _vd_backward_extra(ctx, g_re, g_im, dx_re, dx_im, need_x, need_ls2):
    """Variational part: through s2 = |x|^2 . exp(log_sigma2)^T and y = mu + eps sqrt(max(s2, 1e-8))."""
    xr, xi, wr, wi = ctx.saved_planes
    cplx = xi is not None
    s2, ls2 = ctx.s2, ctx.ls2
    M, N = s2.shape
    dev, dt = s2.device, s2.dtype
    gs2 = torch.empty_like(s2)
    seed, offset, threads = ctx.philox
    er, ei = ctx.eps
    with nv.device_guard(dev):
        nv.check(nv.lib().cplxk_vd_grad_s2(nv.ptr(g_re), nv.ptr(g_im), nv.ptr(s2), nv.ptr(er),
                                           nv.ptr(ei), ctx.noise, seed, offset, threads,
                                           nv.ptr(gs2), M, N, nv.dtype_code(dt), nv.stream_ptr(dev)))
    dls2 = None
    if need_x:   # dq = g_s2 . E ; dx += 2 x dq
        dq, _ = _gemm(gs2, None, _transpose(ls2, TR_EXP), None)
        with nv.device_guard(dev):
            nv.check(nv.lib().cplxk_vd_grad_input(nv.ptr(dx_re), nv.ptr(dx_im), nv.ptr(xr),
                                                  nv.ptr(xi), nv.ptr(dq), dq.numel(),
                                                  nv.dtype_code(dt), nv.stream_ptr(dev)))
    if need_ls2:  # dE = g_s2^T . |x|^2 ; d log_sigma2 = dE * E
        qT = _transpose(xr, TR_ABS2, xi) if cplx else _transpose(xr, TR_SQR)
        dE, _ = _gemm(_transpose(gs2), None, qT, None)
        dls2 = torch.empty_like(dE)
        with nv.device_guard(dev):
            nv.check(nv.lib().cplxk_mul_exp(nv.ptr(dE), nv.ptr(ls2), nv.ptr(dls2), dE.numel(),
                                            nv.dtype_code(dt), 0, nv.stream_ptr(dev)))
    return dls2


def _save_linear(ctx, x_re, x_im, w_re, w_im):
    dt = w_re.dtype
    K = w_re.shape[1]
    ctx.saved_planes = (_flat2d(x_re.to(dt), K), None if x_im is None else _flat2d(x_im.to(dt), K),
                        nv.plane(w_re), nv.plane(w_im))
    ctx.lead = tuple(x_re.shape[:-1])
    ctx.x_dtype = x_re.dtype


def _shape_back(t, lead, last, dtype=None):
    if t is None:
        return None
    t = t.reshape(*lead, last)
    return t if dtype is None or t.dtype == dtype else t.to(dtype)


class _CplxLinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x_re, x_im, w_re, w_im, b_re, b_im):
        if any(ctx.needs_input_grad):
            _save_linear(ctx, x_re, x_im, w_re, w_im)
            _guard(ctx, x_re, x_im, w_re, w_im)
        re, im, _ = _forward_raw(x_re, x_im, w_re, w_im, b_re, b_im, None, None, None, None)
        return re, im

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_re, g_im):
        ctx.saved_tensors
        xr, xi, wr, wi = ctx.saved_planes
        N, K = wr.shape
        g_re, g_im = (_grad2d(g, ctx.lead, N, wr.dtype, wr.device) for g in (g_re, g_im))
        n = ctx.needs_input_grad
        dx_re, dx_im, dw_re, dw_im, db_re, db_im = _linear_backward(
            ctx, g_re, g_im, n[0] or n[1], n[2] or n[3], n[4] or n[5])
        return (_shape_back(dx_re, ctx.lead, K, ctx.x_dtype), _shape_back(dx_im, ctx.lead, K, ctx.x_dtype),
                dw_re, dw_im, db_re, db_im)


class _RealLinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b):
        if any(ctx.needs_input_grad):
            _save_linear(ctx, x, None, w, None)
            _guard(ctx, x, w)
        return _forward_raw(x, None, w, None, b, None, None, None, None, None)[0]

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        ctx.saved_tensors
        xr, _, wr, _ = ctx.saved_planes
        N, K = wr.shape
        g = _grad2d(g, ctx.lead, N, wr.dtype, wr.device)
        n = ctx.needs_input_grad
        dx, _, dw, _, db, _ = _linear_backward(ctx, g, None, n[0], n[1], n[2])
        return _shape_back(dx, ctx.lead, K, ctx.x_dtype), dw, db


def _save_vd(ctx, aux, log_sigma2, noise):
    ctx.s2, ctx.philox, ctx.eps = aux["s2"], aux["philox"], aux["eps"]
    ctx.ls2 = nv.plane(log_sigma2, aux["s2"].dtype)
    ctx.noise = noise


class _CplxLinearVDFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x_re, x_im, w_re, w_im, b_re, b_im, log_sigma2, eps_re, eps_im, noise,
                kl_req=None):
        need = any(ctx.needs_input_grad)
        if need:
            _save_linear(ctx, x_re, x_im, w_re, w_im)
            _guard(ctx, x_re, x_im, w_re, w_im, log_sigma2, eps_re, eps_im)
        re, im, aux = _forward_raw(x_re, x_im, w_re, w_im, b_re, b_im, log_sigma2, eps_re, eps_im,
                                   noise, want_s2=need, kl_req=kl_req)
        if need:
            _save_vd(ctx, aux, log_sigma2, noise)
        return re, im

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_re, g_im):
        ctx.saved_tensors
        xr, xi, wr, wi = ctx.saved_planes
        N, K = wr.shape
        g_re, g_im = (_grad2d(g, ctx.lead, N, wr.dtype, wr.device) for g in (g_re, g_im))
        n = ctx.needs_input_grad
        need_x = n[0] or n[1]
        dx_re, dx_im, dw_re, dw_im, db_re, db_im = _linear_backward(
            ctx, g_re, g_im, need_x, n[2] or n[3], n[4] or n[5])
        dls2 = _vd_backward_extra(ctx, g_re, g_im, dx_re, dx_im, need_x, n[6])
        return (_shape_back(dx_re, ctx.lead, K, ctx.x_dtype), _shape_back(dx_im, ctx.lead, K, ctx.x_dtype),
                dw_re, dw_im, db_re, db_im, dls2, None, None, None, None)


class _RealLinearVDFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, log_sigma2, eps, noise, kl_req=None):
        need = any(ctx.needs_input_grad)
        if need:
            _save_linear(ctx, x, None, w, None)
            _guard(ctx, x, w, log_sigma2, eps)
        y, _, aux = _forward_raw(x, None, w, None, b, None, log_sigma2, eps, None, noise,
                                 want_s2=need, kl_req=kl_req)
        if need:
            _save_vd(ctx, aux, log_sigma2, noise)
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        ctx.saved_tensors
        xr, _, wr, _ = ctx.saved_planes
        N, K = wr.shape
        g = _grad2d(g, ctx.lead, N, wr.dtype, wr.device)
        n = ctx.needs_input_grad
        dx, _, dw, _, db, _ = _linear_backward(ctx, g, None, n[0], n[1], n[2])
        dls2 = _vd_backward_extra(ctx, g, None, dx, None, n[0], n[3])
        return _shape_back(dx, ctx.lead, K, ctx.x_dtype), dw, db, dls2, None, None, None


def _wants_grad(*tensors):
    """False: nothing would be recorded anyway -- call the launcher directly (torch's
    autograd.Function.apply costs more host time than the small-shape kernels take)."""
    if not torch.is_grad_enabled():
        return False
    return any(t is not None and t.requires_grad for t in tensors)


def cplx_linear(x_re, x_im, w_re, w_im, b_re=None, b_im=None):
    """y = x W^T + b on split planes (reference: cplx.linear_naive, cplx.py:634-648)."""
    _check_features(x_re, w_re)
    if not _wants_grad(x_re, x_im, w_re, w_im, b_re, b_im):
        return _forward_raw(x_re, x_im, w_re, w_im, b_re, b_im, None, None, None, None)[:2]
    return _CplxLinearFn.apply(x_re, x_im, w_re, w_im, b_re, b_im)


def real_linear(x, w, b=None):
    _check_features(x, w)
    if not _wants_grad(x, w, b):
        return _forward_raw(x, None, w, None, b, None, None, None, None, None)[0]
    return _RealLinearFn.apply(x, w, b)


class _LinearMaskedFn(torch.autograd.Function):
    """y = x (W * mask)^T + b; gradients: dx = g conj(W * mask), dW = (g^T conj(x)) * mask."""

    @staticmethod
    def forward(ctx, x_re, x_im, w_re, w_im, mask, b_re, b_im):
        if any(ctx.needs_input_grad):
            _save_linear(ctx, x_re, x_im, w_re, w_im)
            ctx.mask = nv.plane(mask.expand(w_re.shape), w_re.dtype)
            _guard(ctx, x_re, x_im, w_re, w_im, mask)
        re, im = _masked_forward_raw(x_re, x_im, w_re, w_im, mask, b_re, b_im)
        ctx.cplx = x_im is not None
        return (re, im) if ctx.cplx else re

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, *grads):
        ctx.saved_tensors
        xr, xi, wr, wi = ctx.saved_planes
        N, K = wr.shape
        g_re = _grad2d(grads[0], ctx.lead, N, wr.dtype, wr.device)
        g_im = _grad2d(grads[1], ctx.lead, N, wr.dtype, wr.device) if ctx.cplx else None
        n = ctx.needs_input_grad
        mk = ctx.mask
        # the masked weight only exists here, where a gradient w.r.t. the input is wanted
        ctx.saved_planes = (xr, xi, _eltwise(TR_MUL, wr, mk), None if wi is None else _eltwise(TR_MUL, wi, mk))
        dx_re, dx_im, dw_re, dw_im, db_re, db_im = _linear_backward(
            ctx, g_re, g_im, n[0] or n[1], n[2] or n[3], n[5] or n[6])
        if dw_re is not None:
            dw_re = _eltwise(TR_MUL, dw_re, mk)
            dw_im = None if dw_im is None else _eltwise(TR_MUL, dw_im, mk)
        return (_shape_back(dx_re, ctx.lead, K, ctx.x_dtype), _shape_back(dx_im, ctx.lead, K, ctx.x_dtype),
                dw_re, dw_im, None, db_re, db_im)


def cplx_linear_masked(x_re, x_im, w_re, w_im, mask, b_re=None, b_im=None):
    """Reference: CplxLinearMasked.forward = cplx.linear(input, weight * mask, bias)
    (nn/masked/complex.py:33-35, nn/masked/base.py:135-149)."""
    return _LinearMaskedFn.apply(x_re, x_im, w_re, w_im, mask, b_re, b_im)


def real_linear_masked(x, w, mask, b=None):
    """Reference: LinearMasked.forward = F.linear(input, weight * mask, bias) (nn/masked/real.py:25-27)."""
    return _LinearMaskedFn.apply(x, None, w, None, mask, b, None)


# ------------------------------------------------------------------------ bilinear
class _OuterFn(torch.autograd.Function):
    """z[..., p * d2 + q] = conj?(x1[..., p]) * x2[..., q]  (cplxk_outer_fwd / _bwd)."""

    @staticmethod
    def forward(ctx, x1_re, x1_im, x2_re, x2_im, conjugate):
        dev = nv.require_cuda(x1_re, x1_im, x2_re, x2_im)
        cplx = x1_im is not None
        dt = x1_re.dtype
        code = nv.dtype_code(dt)
        d1, d2 = x1_re.shape[-1], x2_re.shape[-1]
        lead = torch.broadcast_shapes(x1_re.shape[:-1], x2_re.shape[:-1])
        flat = lambda t, d: None if t is None else nv.plane(t.expand(*lead, d).reshape(-1, d), dt)
        a_re, a_im, b_re, b_im = flat(x1_re, d1), flat(x1_im, d1), flat(x2_re, d2), flat(x2_im, d2)
        B = a_re.shape[0]
        z_re = torch.empty((B, d1 * d2), dtype=dt, device=dev)
        z_im = torch.empty_like(z_re) if cplx else None
        with nv.device_guard(dev):
            nv.check(nv.lib().cplxk_outer_fwd(nv.ptr(a_re), nv.ptr(a_im), nv.ptr(b_re), nv.ptr(b_im),
                                              nv.ptr(z_re), nv.ptr(z_im), B, d1, d2,
                                              1 if conjugate else 0, code, nv.stream_ptr(dev)))
        if any(ctx.needs_input_grad):
            ctx.planes = (a_re, a_im, b_re, b_im)
            ctx.dims = (B, d1, d2, bool(conjugate), tuple(lead), tuple(x1_re.shape), tuple(x2_re.shape))
            _guard(ctx, x1_re, x1_im, x2_re, x2_im)
        ctx.cplx = cplx
        z_re = z_re.reshape(*lead, d1 * d2)
        return (z_re, z_im.reshape(*lead, d1 * d2)) if cplx else z_re

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, *grads):
        ctx.saved_tensors
        a_re, a_im, b_re, b_im = ctx.planes
        B, d1, d2, conj, lead, shp1, shp2 = ctx.dims
        dev, dt = a_re.device, a_re.dtype
        g_re = _grad2d(grads[0], lead, d1 * d2, dt, dev)
        g_im = _grad2d(grads[1], lead, d1 * d2, dt, dev) if ctx.cplx else None
        n = ctx.needs_input_grad
        want1, want2 = n[0] or n[1], n[2] or n[3]
        new = lambda d, want: torch.empty((B, d), dtype=dt, device=dev) if want else None
        d1_re, d2_re = new(d1, want1), new(d2, want2)
        d1_im = new(d1, want1) if ctx.cplx else None
        d2_im = new(d2, want2) if ctx.cplx else None
        with nv.device_guard(dev):
            nv.check(nv.lib().cplxk_outer_bwd(nv.ptr(g_re), nv.ptr(g_im), nv.ptr(a_re), nv.ptr(a_im),
                                              nv.ptr(b_re), nv.ptr(b_im), nv.ptr(d1_re), nv.ptr(d1_im),
                                              nv.ptr(d2_re), nv.ptr(d2_im), B, d1, d2, 1 if conj else 0,
                                              nv.dtype_code(dt), nv.stream_ptr(dev)))

        def back(t, d, shape):   # undo the broadcast of the leading dims
            if t is None:
                return None
            t = t.reshape(*lead, d)
            return t.sum_to_size(shape) if tuple(t.shape) != tuple(shape) else t
        return (back(d1_re, d1, shp1), back(d1_im, d1, shp1), back(d2_re, d2, shp2),
                back(d2_im, d2, shp2), None)


def outer_features(x1_re, x1_im, x2_re, x2_im, conjugate=True):
    return _OuterFn.apply(x1_re, x1_im, x2_re, x2_im, conjugate)


def cplx_bilinear(x1_re, x1_im, x2_re, x2_im, w_re, w_im, b_re=None, b_im=None, conjugate=True,
                  log_sigma2=None, eps=None, kl_req=None):
    """y_j = x1^{H|T} A_j x2 + b_j (cplx.bilinear, cplxmodule/cplx.py:1062-1090) as the complex
    affine map of the outer-product features; with ``log_sigma2`` the local-reparameterisation
    forward of CplxBilinearGaussian (nn/relevance/complex/base.py:70-84)."""
    O = w_re.shape[0]
    z_re, z_im = outer_features(x1_re, x1_im, x2_re, x2_im, conjugate)
    w2 = lambda t: t.reshape(O, -1)
    if log_sigma2 is None:
        return cplx_linear(z_re, z_im, w2(w_re), w2(w_im), b_re, b_im)
    return cplx_linear_vd(z_re, z_im, w2(w_re), w2(w_im), b_re, b_im, w2(log_sigma2), eps=eps,
                          kl_req=kl_req)


def real_bilinear(x1, x2, w, b=None, log_sigma2=None, eps=None, kl_req=None):
    """torch.nn.functional.bilinear / BilinearGaussian.forward (nn/relevance/real/base.py:52-80)."""
    O = w.shape[0]
    z = outer_features(x1, None, x2, None, False)
    if log_sigma2 is None:
        return real_linear(z, w.reshape(O, -1), b)
    return real_linear_vd(z, w.reshape(O, -1), b, log_sigma2.reshape(O, -1), eps=eps, kl_req=kl_req)


def _noise_args(eps):
    if eps is None:
        return None, None, _NOISE[_state["noise"]]
    return eps[0], eps[1], nv.NOISE_INJECT


def cplx_linear_vd(x_re, x_im, w_re, w_im, b_re, b_im, log_sigma2, eps=None, kl_req=None):
    """Fused local-reparameterisation forward (reference: CplxLinearGaussian.forward,
    nn/relevance/complex/base.py:43-56). ``eps=(eps_re, eps_im)`` injects the noise
    (each ~ N(0, 1/2)); ``None`` draws it inside the kernel.  ``kl_req``: see ``_forward_raw``."""
    _check_features(x_re, w_re)
    er, ei, mode = _noise_args(eps)
    if not _wants_grad(x_re, x_im, w_re, w_im, b_re, b_im, log_sigma2, er, ei):
        return _forward_raw(x_re, x_im, w_re, w_im, b_re, b_im, log_sigma2, er, ei, mode,
                            kl_req=kl_req)[:2]
    return _CplxLinearVDFn.apply(x_re, x_im, w_re, w_im, b_re, b_im, log_sigma2, er, ei, mode,
                                 kl_req)


def real_linear_vd(x, w, b, log_sigma2, eps=None, kl_req=None):
    """Reference: LinearGaussian.forward, nn/relevance/real/base.py:43-49."""
    _check_features(x, w)
    er, _, mode = _noise_args((eps, None) if eps is not None else None)
    if not _wants_grad(x, w, b, log_sigma2, er):
        return _forward_raw(x, None, w, None, b, None, log_sigma2, er, None, mode, kl_req=kl_req)[0]
    return _RealLinearVDFn.apply(x, w, b, log_sigma2, er, mode, kl_req)


_stale_flags = {}     # device index -> pinned int32[1] the guard kernel raises on a mismatch


def _stale_flag(dev):
    flag = _stale_flags.get(dev.index)
    if flag is None:
        flag = _stale_flags[dev.index] = torch.zeros(1, dtype=torch.int32).pin_memory()
    return flag


def _check_stale(dev):
    flag = _stale_flags.get(dev.index)
    if flag is not None and int(flag[0]) != 0:
        flag[0] = 0
        raise RuntimeError(
            "cplxmodule_b200: the parameters of a variational layer were modified through `.data` "
            "(or another write autograd does not see) between its forward and penalties(); the KL "
            "sum handed out for that step was NaN.  Edit parameters before the forward, or call "
            "cplxmodule_b200.set_kl_fusion(False).")


class FusedKLCache:
    """KL sum produced by the last training-mode forward of a layer, handed out once and only
    for unchanged parameters.  Two checks: on the host tensor identity, storage and autograd
    version counters (catches optimizer steps, ``load_state_dict``, ``.to()``); on the device a
    fingerprint of a strided sample of the parameters taken after the forward and again at
    hand-out (``cplxk_kl_guard``: catches writes through ``param.data``, which bump no version
    counter).  On a device mismatch the value handed out is NaN and the next call into the
    cache raises -- nothing synchronises."""

    def __init__(self):
        self._entry = None

    @staticmethod
    def _key(params):
        return tuple((id(p), p.data_ptr(), p._version, p.dtype) for p in params if p is not None)

    @staticmethod
    def _planes(params):
        # (w_re, w_im | None, log_sigma2) as dense [N, K] planes (views when already dense)
        w_re, w_im, ls2 = params if len(params) == 3 else (params[0], None, params[1])
        N = w_re.shape[0]
        flat = lambda t: None if t is None else nv.plane(t.detach()).reshape(N, -1)
        return flat(w_re), flat(w_im), flat(ls2.to(w_re.dtype) if ls2.dtype != w_re.dtype else ls2)

    def _guard(self, params, fp, fused, out):
        wr, wi, ls2 = self._planes(params)
        dev = wr.device
        with nv.device_guard(dev):
            kl_ws = nv.kl_workspace(dev)
            nv.check(nv.lib().cplxk_kl_guard(
                nv.ptr(wr), nv.ptr(wi), nv.ptr(ls2), wr.shape[0], wr.shape[1], nv.dtype_code(wr.dtype),
                None, nv.ptr(fp), nv.ptr(fused), nv.ptr(out),
                ctypes.c_void_p(_stale_flag(dev).data_ptr()), nv.ptr(kl_ws), kl_ws.numel() * 8,
                nv.stream_ptr(dev)))

    def put(self, params, kl_req):
        self._entry = None
        if kl_req is not None and "sum" in kl_req:
            _check_stale(kl_req["sum"].device)
            self._entry = (self._key(params), kl_req["sum"], kl_req.get("rows"), kl_req.get("event"),
                           kl_req.get("fp"))

    def take(self, params, rows=None, with_event=False):
        """The cached sum if it covers ``rows`` (None: the whole layer) of unchanged parameters."""
        entry, self._entry = self._entry, None
        if entry is not None and entry[0] == self._key(params) and entry[2] == rows:
            value, fp = entry[1], entry[4]
            _check_stale(value.device)
            if fp is not None:
                if entry[3] is not None:     # shard request: the sum is final at this event (after the pre-pass)
                    torch.cuda.current_stream(value.device).wait_event(entry[3])
                out = torch.empty_like(value)
                self._guard(params, fp, value, out)
                value = out
            return (value, entry[3]) if with_event else value
        return (None, None) if with_event else None


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
    with nv.device_guard(dev):
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
    def forward(ctx, kind, reduction, w_re, w_im, log_sigma2, precomputed=None):
        if any(ctx.needs_input_grad):
            dt = w_re.dtype
            ctx.kind, ctx.reduction = kind, reduction
            ctx.planes = (nv.plane(w_re), nv.plane(w_im), nv.plane(log_sigma2, dt))
            ctx.shape = tuple(w_re.shape)
            _guard(ctx, w_re, w_im, log_sigma2)
        if precomputed is not None and reduction in ("sum", "mean"):
            # [sum] from the forward's operand pre-pass (FusedKLCache); the list hides it from autograd
            out = precomputed[0]
            if reduction == "mean":
                out = out / max(w_re.numel(), 1)
            return out.to(w_re.dtype) if w_re.dtype != torch.float32 else out
        return kl_penalty(kind, w_re, w_im, log_sigma2, reduction)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad):
        ctx.saved_tensors
        wr, wi, ls2 = ctx.planes
        dev, dt = wr.device, wr.dtype
        n = wr.numel()
        grad = grad.contiguous()
        if grad.dtype not in (torch.float32, dt):
            grad = grad.float()
        per_elem = ctx.reduction is None
        scale = 1.0 / max(n, 1) if ctx.reduction == "mean" else 1.0
        d_wr, d_ls2 = torch.empty_like(wr), torch.empty_like(ls2)
        d_wi = torch.empty_like(wi) if wi is not None else None
        with nv.device_guard(dev):
            nv.check(nv.lib().cplxk_kl_bwd(ctx.kind, nv.ptr(wr), nv.ptr(wi), nv.ptr(ls2), n,
                                           nv.dtype_code(dt), nv.ptr(grad), 1 if per_elem else 0,
                                           1 if grad.dtype == torch.float32 else 0, scale,
                                           nv.ptr(d_wr), nv.ptr(d_wi), nv.ptr(d_ls2),
                                           nv.stream_ptr(dev)))
        return None, None, d_wr, d_wi, d_ls2, None


def kl(kind, w_re, w_im, log_sigma2, reduction="sum", precomputed=None):
    """``precomputed``: 0-d float32 sum from ``FusedKLCache.take`` (or None)."""
    if not _wants_grad(w_re, w_im, log_sigma2):
        if precomputed is not None and reduction in ("sum", "mean"):
            out = precomputed / max(w_re.numel(), 1) if reduction == "mean" else precomputed
            return out.to(w_re.dtype) if w_re.dtype != torch.float32 else out
        return kl_penalty(kind, w_re, w_im, log_sigma2, reduction)
    pre = None if precomputed is None else [precomputed]
    return _KLFn.apply(kind, reduction, w_re, w_im, log_sigma2, pre)


def _log_alpha_raw(w_re, w_im, log_sigma2, threshold=None):
    dev = nv.require_cuda(w_re, w_im, log_sigma2)
    dt = w_re.dtype
    code = nv.dtype_code(dt)
    wr, wi, ls2 = nv.plane(w_re), nv.plane(w_im), nv.plane(log_sigma2, dt)
    out = torch.empty_like(wr)
    lib = nv.lib()
    with nv.device_guard(dev):
        st = nv.stream_ptr(dev)
        if threshold is None:
            nv.check(lib.cplxk_log_alpha(nv.ptr(wr), nv.ptr(wi), nv.ptr(ls2), wr.numel(), code,
                                         nv.ptr(out), 0.0, None, st))
        else:
            nv.check(lib.cplxk_log_alpha(nv.ptr(wr), nv.ptr(wi), nv.ptr(ls2), wr.numel(), code,
                                         None, float(threshold), nv.ptr(out), st))
    return out


class _LogAlphaFn(torch.autograd.Function):
    """log_alpha = log_sigma2 - log(|w|^2): d/d log_sigma2 = 1, d/d w = -2 w / |w|^2.  The
    forward is the device kernel; the backward of this (off the hot path: user-defined
    regularisers on ``layer.log_alpha``) is three elementwise torch ops and is itself
    differentiable."""

    @staticmethod
    def forward(ctx, w_re, w_im, log_sigma2):
        if any(ctx.needs_input_grad):
            _guard(ctx, w_re, w_im, log_sigma2)
            ctx.cplx = w_im is not None
        return _log_alpha_raw(w_re, w_im, log_sigma2)

    @staticmethod
    def backward(ctx, grad):
        saved = ctx.saved_tensors
        w_re, w_im = saved[0], (saved[1] if ctx.cplx else None)
        # reference: ls2 - 2 log(|w| + 1e-12)  =>  d/dw = -2 w / (|w| (|w| + 1e-12)); the device
        # kernel evaluates log(|w|^2 + 1e-24): same value and gradient wherever |w| >> 1e-12
        r2 = w_re * w_re if w_im is None else w_re * w_re + w_im * w_im
        k = -2.0 * grad / (r2 + 1e-24)
        return k * w_re, (None if w_im is None else k * w_im), grad


def log_alpha(w_re, w_im, log_sigma2, threshold=None):
    """``log_sigma2 - 2 log(|w| + 1e-12)`` (nn/relevance/{real,complex}/base.py), differentiable in
    the weights and ``log_sigma2`` like the reference's property; with a ``threshold`` the
    relevance mask ``(log_alpha <= threshold)`` as floats (no graph, as in the reference)."""
    if threshold is not None:
        return _log_alpha_raw(w_re, w_im, log_sigma2, threshold)
    return _LogAlphaFn.apply(w_re, w_im, log_sigma2)


def kl_and_mask(kind, w_re, w_im, log_sigma2, threshold, reduction="sum"):
    """(KL sum, relevance mask) of a layer from ONE pass over its parameters
    (``cplxk_kl_mask``): what a sparsification schedule needs each time it logs the penalty and
    re-derives the masks.  No autograd graph."""
    dev = nv.require_cuda(w_re, w_im, log_sigma2)
    dt = w_re.dtype
    code = nv.dtype_code(dt)
    wr, wi, ls2 = nv.plane(w_re), nv.plane(w_im), nv.plane(log_sigma2, dt)
    n = wr.numel()
    if reduction not in ("sum", "mean"):
        raise ValueError(f"`reduction` must be `sum` or `mean`. Got {reduction}.")
    mask = torch.empty_like(wr)
    out = torch.empty((), dtype=torch.float32, device=dev)
    with nv.device_guard(dev):
        ws = nv.kl_workspace(dev)
        scale = 1.0 if reduction == "sum" else 1.0 / max(n, 1)
        nv.check(nv.lib().cplxk_kl_mask(kind, nv.ptr(wr), nv.ptr(wi), nv.ptr(ls2), n, code,
                                        float(threshold), nv.ptr(mask), nv.ptr(out), scale,
                                        nv.ptr(ws), ws.numel() * 8, nv.stream_ptr(dev)))
    return (out.to(dt) if dt != torch.float32 else out), mask


def randn_philox_torch(n, seed, offset, threads, scale=1.0, device="cuda"):
    """Test hook: the epilogue's torch-layout normal generator as a standalone fill."""
    device = torch.device(device)
    out = torch.empty(n, dtype=torch.float32, device=device)
    with torch.cuda.device(device):
        nv.check(nv.lib().cplxk_randn_philox_torch(nv.ptr(out), n, seed, offset, threads,
                                                   ctypes.c_float(scale), nv.stream_ptr(device)))
    return out
