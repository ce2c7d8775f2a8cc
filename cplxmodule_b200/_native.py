"""ctypes binding of ``csrc/libcplxk.so`` (C ABI declared in ``include/cplxk.h``).

This is the only module that touches the native library.  There is no CPU or
pure-torch implementation of the hot path behind it: if the library is missing
or the tensors are not on a CUDA (sm_100) device the calls raise.
"""
import ctypes
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CPLXK_LIB") or os.path.join(_HERE, "csrc", "libcplxk.so")  # CPLXK_LIB: A/B builds

F32, BF16 = 0, 1
MATH_AUTO, MATH_TENSOR, MATH_SIMT, MATH_TENSOR_TF32 = 0, 1, 2, 3
NOISE_INJECT, NOISE_PHILOX_TORCH, NOISE_PHILOX_FAST = 0, 1, 2
ERR_UNSUPPORTED, ERR_WORKSPACE = -5, -6
KL_REAL_VD, KL_REAL_ARD, KL_CPLX_VD, KL_CPLX_ARD = 0, 1, 2, 3
KL_CPLX_VD_APPROX, KL_CPLX_VD_SCALEFREE = 4, 5        # nn/relevance/extensions/complex.py

EXPORTS = (
    "cplxk_abi_version", "cplxk_strerror", "cplxk_device_info", "cplxk_set_sm_reserve", "cplxk_linear_fwd",
    "cplxk_linear_vd_prepare",
    "cplxk_linear_fwd_ws", "cplxk_linear_workspace_bytes",
    "cplxk_linear_vd_fwd", "cplxk_linear_vd_fwd_kl", "cplxk_linear_vd_workspace_bytes", "cplxk_kl_workspace_bytes", "cplxk_kl", "cplxk_log_alpha",
    "cplxk_conv2d_fwd", "cplxk_conv2d_workspace_bytes", "cplxk_conv2d_fwd_g",
    "cplxk_conv2d_workspace_bytes_g", "cplxk_randn_philox_torch",
    "cplxk_transpose2d", "cplxk_eltwise", "cplxk_colsum", "cplxk_vd_grad_s2", "cplxk_vd_grad_input",
    "cplxk_mul_exp", "cplxk_kl_bwd",
    "cplxk_linear_masked_fwd", "cplxk_linear_masked_workspace_bytes", "cplxk_kl_mask",
    "cplxk_outer_fwd", "cplxk_outer_bwd", "cplxk_kl_guard", "cplxk_linear_vd_fuses_kl", "cplxk_vd_combine",
)

_lock = threading.Lock()
_lib = None

_vp, _i64, _u64, _u32, _int = (ctypes.c_void_p, ctypes.c_int64, ctypes.c_uint64,
                               ctypes.c_uint32, ctypes.c_int)


def _declare(lib):
    lib.cplxk_abi_version.restype = _int
    lib.cplxk_strerror.restype = ctypes.c_char_p
    lib.cplxk_strerror.argtypes = [_int]
    lib.cplxk_device_info.argtypes = [ctypes.POINTER(_int)] * 3
    lib.cplxk_linear_fwd.argtypes = [_vp] * 8 + [_i64] * 3 + [_int, _int, _vp]
    lib.cplxk_linear_fwd_ws.argtypes = [_vp] * 8 + [_i64] * 3 + [_int, _int, _vp, ctypes.c_size_t, _vp]
    lib.cplxk_linear_workspace_bytes.restype = ctypes.c_size_t
    lib.cplxk_linear_workspace_bytes.argtypes = [_i64, _i64, _i64, _int]
    lib.cplxk_linear_vd_fwd.argtypes = ([_vp] * 9 + [_int, _u64, _u64, _u32] + [_vp] * 2
                                        + [_i64] * 3 + [_int, _int, _vp, _vp, ctypes.c_size_t, _vp])
    lib.cplxk_linear_vd_fwd_kl.argtypes = ([_vp] * 9 + [_int, _u64, _u64, _u32] + [_vp] * 2
                                           + [_i64] * 3 + [_int, _int, _vp, _vp, ctypes.c_size_t]
                                           + [_int, _vp, _vp, ctypes.c_size_t, _i64, _i64, _vp, _vp,
                                              ctypes.POINTER(_int), _vp])
    lib.cplxk_set_sm_reserve.argtypes = [_int]
    lib.cplxk_linear_vd_prepare.argtypes = ([_vp] * 5 + [_i64] * 3 + [_int, _vp, ctypes.c_size_t, _int, _vp, _vp,
                                            ctypes.c_size_t, _vp])
    lib.cplxk_linear_vd_workspace_bytes.restype = ctypes.c_size_t
    lib.cplxk_linear_vd_workspace_bytes.argtypes = [_i64, _i64, _i64, _int]
    lib.cplxk_kl_workspace_bytes.restype = ctypes.c_size_t
    lib.cplxk_kl.argtypes = [_int, _vp, _vp, _vp, _i64, _int, _vp, _vp, ctypes.c_double,
                             _vp, ctypes.c_size_t, _vp]
    lib.cplxk_log_alpha.argtypes = [_vp, _vp, _vp, _i64, _int, _vp, ctypes.c_float, _vp, _vp]
    lib.cplxk_conv2d_fwd.argtypes = ([_vp] * 9 + [_int, _u64, _u64, _u32] + [_vp] * 2
                                     + [_i64] * 13 + [_int, _int, _int, _vp, ctypes.c_size_t, _vp])
    lib.cplxk_conv2d_workspace_bytes.restype = ctypes.c_size_t
    lib.cplxk_conv2d_workspace_bytes.argtypes = [_i64] * 7 + [_int, _int]
    lib.cplxk_conv2d_fwd_g.argtypes = ([_vp] * 9 + [_int, _u64, _u64, _u32] + [_vp] * 2
                                       + [_i64] * 14 + [_int, _int, _int, _vp, ctypes.c_size_t, _vp])
    lib.cplxk_conv2d_workspace_bytes_g.restype = ctypes.c_size_t
    lib.cplxk_conv2d_workspace_bytes_g.argtypes = [_i64] * 8 + [_int, _int, _int]
    lib.cplxk_randn_philox_torch.argtypes = [_vp, _i64, _u64, _u64, _u32, ctypes.c_float, _vp]
    lib.cplxk_transpose2d.argtypes = [_vp, _vp, _vp, _i64, _i64, _int, _int, _vp]
    lib.cplxk_eltwise.argtypes = [_int, _vp, _vp, _vp, _i64, _int, _vp]
    lib.cplxk_colsum.argtypes = [_vp, _vp, _i64, _i64, _int, _vp]
    lib.cplxk_vd_grad_s2.argtypes = [_vp] * 5 + [_int, _u64, _u64, _u32, _vp, _i64, _i64, _int, _vp]
    lib.cplxk_vd_combine.argtypes = [_vp] * 5 + [_int, _u64, _u64, _u32, _i64, _int, _vp]
    lib.cplxk_vd_grad_input.argtypes = [_vp] * 5 + [_i64, _int, _vp]
    lib.cplxk_mul_exp.argtypes = [_vp, _vp, _vp, _i64, _int, _int, _vp]
    lib.cplxk_kl_bwd.argtypes = [_int, _vp, _vp, _vp, _i64, _int, _vp, _int, _int, ctypes.c_double,
                                 _vp, _vp, _vp, _vp]
    lib.cplxk_linear_vd_fuses_kl.argtypes = [_i64, _i64, _i64, _int, _int]
    lib.cplxk_kl_guard.argtypes = [_vp, _vp, _vp, _i64, _i64, _int, _vp, _vp, _vp, _vp, _vp, _vp,
                                   ctypes.c_size_t, _vp]
    lib.cplxk_linear_masked_workspace_bytes.restype = ctypes.c_size_t
    lib.cplxk_linear_masked_workspace_bytes.argtypes = [_i64, _i64, _i64, _int]
    lib.cplxk_linear_masked_fwd.argtypes = [_vp] * 9 + [_i64] * 3 + [_int, _int, _vp, ctypes.c_size_t, _vp]
    lib.cplxk_kl_mask.argtypes = [_int, _vp, _vp, _vp, _i64, _int, ctypes.c_float, _vp, _vp,
                                  ctypes.c_double, _vp, ctypes.c_size_t, _vp]
    lib.cplxk_outer_fwd.argtypes = [_vp] * 6 + [_i64] * 3 + [_int, _int, _vp]
    lib.cplxk_outer_bwd.argtypes = [_vp] * 10 + [_i64] * 3 + [_int, _int, _vp]
    for name in EXPORTS:
        getattr(lib, name)  # fail at load time, not at first use, if a symbol is missing


def lib():
    """Load (once) the native library; raise loudly when it is not there."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise RuntimeError(
                        f"cplxmodule_b200: native library {LIB_PATH} is missing. Build it with "
                        "`python -m cplxmodule_b200.build` (needs nvcc). There is no CPU or "
                        "pure-PyTorch fallback for the hot path."
                    )
                handle = ctypes.CDLL(LIB_PATH)
                _declare(handle)
                if handle.cplxk_abi_version() != 1:
                    raise RuntimeError("cplxmodule_b200: libcplxk.so ABI version mismatch; rebuild")
                _lib = handle
    return _lib


def check(rc):
    if rc != 0:
        raise RuntimeError(lib().cplxk_strerror(rc).decode())


def dtype_code(dtype):
    if dtype == torch.float32:
        return F32
    if dtype == torch.bfloat16:
        return BF16
    raise TypeError(f"cplxmodule_b200 kernels take float32 or bfloat16 planes, got {dtype}")


def require_cuda(*tensors):
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError(
                "cplxmodule_b200 runs its linear/conv/variational-dropout/KL path only as "
                "sm_100a CUDA kernels: got a CPU tensor and there is no CPU fallback. "
                "Move the module and its inputs to a B200 (`.cuda()`)."
            )
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError(f"tensors on different devices: {dev} vs {t.device}")
    return dev


def ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


class _NoGuard:
    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


_NO_GUARD = _NoGuard()


def device_guard(device):
    """``torch.cuda.device(device)`` only when ``device`` is not already current (the context
    manager costs several microseconds per call; single-GPU processes never need it)."""
    if device.index is None or device.index == torch.cuda.current_device():
        return _NO_GUARD
    return torch.cuda.device(device)


def stream_ptr(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def plane(t, dtype=None):
    """Dense row-major plane (the C ABI wants unit inner stride, 16-byte aligned rows)."""
    if t is None:
        return None
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


# --------------------------------------------------------------- philox bookkeeping
_props = {}          # device index -> (SM count, resident 256-thread blocks per SM)
_capture_noise = {"offset": 0}


def _launch_props(device):
    p = _props.get(device.index)
    if p is None:
        props = torch.cuda.get_device_properties(device)
        p = _props[device.index] = (props.multi_processor_count, props.max_threads_per_multi_processor // 256)
    return p


def philox_plan(device, numel, torch_exact=True):
    """(generator, seed, offset, threads, increment) replicating torch's CUDA ``normal_`` launch
    for ``numel`` floats (ATen/native/cuda/DistributionTemplates.h: calc_execution_policy).

    While a CUDA graph is being captured torch's generator may not be read or advanced from
    Python: the coordinates then come from a private counter (generator ``None``) and are baked
    into the captured launch -- every replay of the graph repeats that draw."""
    if torch_exact and numel >= 2 ** 31:
        # torch splits such tensors into 32-bit-indexable pieces with one generator advance each;
        # that stream is not reproduced -- ask for the private layout instead
        raise NotImplementedError(
            "torch-exact noise is limited to outputs below 2**31 elements; "
            "use cplxmodule_b200.set_noise_mode('fast') for this size")
    sms, blocks_per_sm = _launch_props(device)
    block = 256
    grid = min(sms * blocks_per_sm, (numel + block - 1) // block)
    grid = max(grid, 1)
    threads = block * grid
    increment = ((numel - 1) // (threads * 4) + 1) * 4
    if torch.cuda.is_current_stream_capturing():
        offset = _capture_noise["offset"]
        _capture_noise["offset"] = offset + increment
        return None, torch.initial_seed() & 0xFFFFFFFFFFFFFFFF, offset, threads, increment
    gen = torch.cuda.default_generators[device.index]
    offset = gen.get_offset()
    offset = (offset + 3) // 4 * 4
    return gen, gen.initial_seed(), offset, threads, increment


def philox_advance(gen, offset, increment):
    """Advance torch's generator past the draw a kernel made (nothing to do for the private
    coordinates used under graph capture)."""
    if gen is not None:
        gen.set_offset(offset + increment)


# ------------------------------------------------------------------ scratch workspace
_ws = {}


def workspace(device, nbytes):
    """Grow-only scratch buffer per (device, stream).  Kernels of one stream are ordered, so
    consecutive calls on it can share the buffer: no allocator round trip per call.  Under CUDA
    graph capture a fresh tensor is returned instead (it then belongs to the graph's pool)."""
    if nbytes <= 0:
        return None
    if torch.cuda.is_current_stream_capturing():
        return torch.empty(nbytes, dtype=torch.uint8, device=device)
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    buf = _ws.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = None
        _ws.pop(key, None)
        buf = torch.empty(nbytes, dtype=torch.uint8, device=device)
        _ws[key] = buf
    return buf


def release_workspaces():
    """Drop the cached scratch buffers (they are re-created on demand)."""
    _ws.clear()
    _kl_ws.clear()


# -------------------------------------------------------------------- KL workspace
_kl_ws = {}


def kl_workspace(device):
    if torch.cuda.is_current_stream_capturing():      # belongs to the graph's pool: never cached
        return torch.zeros((lib().cplxk_kl_workspace_bytes() + 7) // 8, dtype=torch.int64, device=device)
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    ws = _kl_ws.get(key)
    if ws is None:
        nbytes = lib().cplxk_kl_workspace_bytes()
        ws = torch.zeros((nbytes + 7) // 8, dtype=torch.int64, device=device)
        _kl_ws[key] = ws
    return ws
