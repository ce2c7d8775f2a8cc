"""CPU-side checks: API surface / state-dict format / init parity with the reference,
penalty collection logic, loud failure without CUDA, C-ABI export list."""
import ctypes
import os
import re

import pytest
import torch

import cplxmodule_b200 as cb
from cplxmodule_b200 import _native, cplx
from cplxmodule_b200.nn import CplxConv1d, CplxConv2d, CplxLinear, CplxParameter, RealToCplx
from cplxmodule_b200.nn.relevance import (BaseARD, CplxConv2dVD, CplxLinearARD, CplxLinearVD,
                                          LinearARD, LinearVD, compute_ard_masks, named_penalties,
                                          penalties)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HAVE_REF = os.path.isdir("/root/reference/cplxmodule")


def test_cabi_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "cplxk.h")).read()
    declared = set(re.findall(r"\b(cplxk_[a-z0-9_]+)\s*\(", header))
    declared -= {"cplxk_status", "cplxk_dtype", "cplxk_math", "cplxk_noise", "cplxk_kl_kind"}
    assert declared == set(_native.EXPORTS)
    lib = ctypes.CDLL(_native.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.cplxk_abi_version() == 1


def test_library_reports_errors_not_aborts_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _native.lib()
    rc = lib.cplxk_device_info(None, None, None)
    assert rc == -4 and b"CUDA error" in lib.cplxk_strerror(rc)
    assert lib.cplxk_kl(0, None, None, None, 5, 0, None, None, 1.0, None, 0, None) == -1
    # fp32 planes: three 16-bit operand planes per side + the two row-scale vectors
    assert lib.cplxk_linear_vd_workspace_bytes(4096, 4096, 4096, 0) == 6 * 4096 * 4096 * 2 + 2 * 4096 * 4
    assert lib.cplxk_linear_vd_workspace_bytes(4096, 4096, 4096, 1) == 2 * 4096 * 4096 * 2


def test_cpu_tensors_fail_loudly():
    layer = CplxLinearVD(8, 4)
    z = cplx.Cplx(torch.randn(3, 8), torch.randn(3, 8))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        layer(z)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        sum(penalties(layer))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        LinearVD(8, 4)(torch.randn(3, 8))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        CplxConv2d(2, 3, 3)(cplx.Cplx(torch.randn(1, 2, 5, 5), torch.randn(1, 2, 5, 5)))


def test_state_dict_format():
    assert sorted(CplxLinearVD(5, 3).state_dict()) == [
        "bias.imag", "bias.real", "log_sigma2", "weight.imag", "weight.real"]
    assert sorted(LinearARD(5, 3).state_dict()) == ["bias", "log_sigma2", "weight"]
    assert sorted(CplxConv2dVD(2, 3, 3, bias=False).state_dict()) == [
        "log_sigma2", "weight.imag", "weight.real"]
    m = CplxLinearARD(5, 3)
    assert isinstance(m.weight, cplx.Cplx) and m.weight.shape == (3, 5)
    assert isinstance(m._modules["weight"], CplxParameter)
    assert float(m.log_sigma2.min()) == -10.0 == float(m.log_sigma2.max())


def test_real_to_complex_promotion_on_load():
    dense = torch.nn.Linear(6, 4)
    vd = CplxLinearVD(6, 4)
    missing, unexpected = vd.load_state_dict(dense.state_dict(), strict=False)
    assert missing == ["log_sigma2"] and unexpected == []
    assert torch.equal(vd.weight.real, dense.weight) and float(vd.weight.imag.abs().max()) == 0.0
    src = CplxLinear(6, 4)
    vd.load_state_dict(src.state_dict(), strict=False)
    assert torch.equal(vd.weight.imag, src.weight.imag)
    bad = {k: v for k, v in src.state_dict().items() if k != "weight.imag"}
    with pytest.raises(RuntimeError, match="requires both"):
        CplxLinear(6, 4).load_state_dict(bad)


def test_cplx_container_semantics():
    a = cplx.Cplx(torch.randn(3, 4), torch.randn(3, 4))
    b = cplx.Cplx(torch.randn(3, 4), torch.randn(3, 4))
    p = a * b
    ref = torch.complex(a.real, a.imag) * torch.complex(b.real, b.imag)
    assert torch.allclose(p.real, ref.real) and torch.allclose(p.imag, ref.imag)
    q = a / b
    ref = torch.complex(a.real, a.imag) / torch.complex(b.real, b.imag)
    assert torch.allclose(q.real, ref.real, atol=1e-5) and torch.allclose(q.imag, ref.imag, atol=1e-5)
    assert torch.allclose(abs(a), torch.complex(a.real, a.imag).abs())
    assert cplx.Cplx(a) is a
    with pytest.raises(TypeError):
        cplx.Cplx([1.0])
    with pytest.raises(ValueError):
        cplx.Cplx(torch.zeros(2), torch.zeros(3))
    with pytest.raises(AttributeError):
        a.real = torch.zeros(3, 4)
    r = torch.randn(2, 10)
    z = cplx.from_interleaved_real(r)
    assert torch.equal(z.real, r[:, 0::2]) and torch.equal(cplx.to_interleaved_real(z), r)
    z = cplx.from_concatenated_real(r)
    assert torch.equal(z.imag, r[:, 5:]) and torch.equal(cplx.to_concatenated_real(z), r)
    assert RealToCplx()(r).shape == (2, 5)
    assert cplx.cat([a, b], dim=0).shape == (6, 4) and cplx.stack([a, b], dim=0).shape == (2, 3, 4)
    torch.manual_seed(3)
    n = cplx.randn(1000, 50)
    assert abs(float((n.real ** 2 + n.imag ** 2).mean()) - 1.0) < 0.02


class _Toy(BaseARD):
    def __init__(self, value):
        super().__init__()
        self.value = torch.nn.Parameter(torch.tensor(value))

    @property
    def penalty(self):
        return self.value * torch.ones(2, 3)

    def relevance(self, **kw):
        return torch.ones(2, 3)


def test_penalty_collection_walk():
    shared = _Toy(2.0)
    net = torch.nn.Sequential(shared, torch.nn.ReLU(), _Toy(1.0), shared)
    got = dict(named_penalties(net, reduction="sum"))
    assert list(got) == ["0", "2"]            # shared module visited once
    assert float(got["0"]) == 12.0 and float(got["2"]) == 6.0
    assert float(sum(penalties(net, reduction="mean"))) == 3.0
    assert [p.shape for p in penalties(net, reduction=None)] == [(2, 3), (2, 3)]
    with pytest.raises(ValueError):
        list(penalties(net, reduction="max"))
    assert sorted(compute_ard_masks(net)) == ["0.mask", "2.mask"]
    assert sorted(compute_ard_masks(net, prefix="enc")) == ["enc.0.mask", "enc.2.mask"]


def test_conv_module_ctor_matches_torch():
    m = CplxConv2d(4, 6, (3, 2), stride=2, padding=(1, 0), dilation=(1, 2), bias=False)
    assert m.weight.shape == (6, 4, 3, 2) and m.bias is None and m.stride == (2, 2)
    assert CplxConv1d(4, 6, 5).weight.shape == (6, 4, 5)
    with pytest.raises(ValueError):
        CplxConv2d(3, 6, 3, groups=2)
    with pytest.raises(ValueError):
        CplxConv2dVD(2, 2, 3, padding_mode="circular")


@pytest.mark.skipif(not HAVE_REF, reason="/root/reference not present")
def test_default_init_bit_parity_with_reference():
    from oracle.make_golden import import_reference
    import_reference()
    from cplxmodule.nn import CplxLinear as RefLinear, CplxConv2d as RefConv2d
    from cplxmodule.nn.relevance import CplxLinearVD as RefVD, LinearVD as RefRealVD
    for ours, ref, args in ((CplxLinear, RefLinear, (31, 17)), (CplxLinearVD, RefVD, (31, 17)),
                            (LinearVD, RefRealVD, (31, 17)), (CplxConv2d, RefConv2d, (3, 5, 3))):
        torch.manual_seed(99)
        a = ours(*args).state_dict()
        torch.manual_seed(99)
        b = ref(*args).state_dict()
        assert list(a) == list(b)
        for k in a:
            assert torch.equal(a[k], b[k]), k


def test_masked_layer_mask_semantics_and_state_dict():
    from cplxmodule_b200.nn.masked import (CplxLinearMasked, LinearMasked, binarize_masks,
                                           deploy_masks, named_masks)
    m = CplxLinearMasked(6, 4)
    assert not m.is_sparse and "mask" not in m.state_dict()
    with pytest.raises(RuntimeError, match="no sparsity mask"):
        m.weight_masked
    m.mask = torch.ones(6)                      # broadcast to the weight's shape
    assert m.is_sparse and m.mask.shape == (4, 6) and "mask" in m.state_dict()
    wm = m.weight_masked
    assert isinstance(wm, cplx.Cplx) and torch.equal(wm.real, m.weight.real)
    m.mask = None
    assert not m.is_sparse
    with pytest.raises(TypeError):
        m.mask = [1, 0]
    # dense -> VD -> masked transfer through state dicts (nn/relevance/README.md:77-89)
    vd = CplxLinearVD(6, 4)
    masks = {"mask": (torch.rand(4, 6) > 0.5).float()}
    state, masks = binarize_masks(vd.state_dict(), masks)
    missing, unexpected = m.load_state_dict(state, strict=False)
    assert "log_sigma2" in unexpected and not m.is_sparse
    deploy_masks(m, state_dict=masks)
    assert m.is_sparse and torch.equal(m.mask, masks["mask"])
    assert dict(named_masks(m))[""] is m.mask
    net = torch.nn.Sequential(LinearMasked(5, 3), torch.nn.ReLU(), LinearMasked(3, 2))
    deploy_masks(net, state_dict={"0.mask": torch.zeros(3, 5)})
    assert net[0].is_sparse and not net[2].is_sparse
    deploy_masks(net, state_dict={}, reset=True)
    assert not net[0].is_sparse
    sd = net.state_dict()
    sd["2.mask"] = torch.ones(2, 3)
    net.load_state_dict(sd, strict=False)
    assert net[2].is_sparse


def test_fused_kl_cache_bookkeeping_and_shard_requests():
    """host logic of the pre-pass KL by-product: which rows a forward is asked for under
    set_kl_shard, and when a cached sum may be handed out (rows, parameter identity, version)"""
    from cplxmodule_b200 import ops
    from cplxmodule_b200.distributed import row_shard
    try:
        assert ops.kl_request(None, 10) is None
        assert ops.kl_request(2, 10) == {"kind": 2}
        for world in (2, 3, 8):
            covered = []
            for rank in range(world):
                ops.set_kl_shard(rank, world)
                req = ops.kl_request(3, 4099)
                assert req["rows"] == row_shard(4099, rank, world)
                covered.append(req["rows"])
            assert covered[0][0] == 0 and covered[-1][1] == 4099
            assert all(a[1] == b[0] for a, b in zip(covered, covered[1:]))
        ops.set_kl_shard(0, 1)                      # a world of one is not sharded
        assert "rows" not in ops.kl_request(3, 16)
        ops.set_kl_fusion(False)
        assert ops.kl_request(3, 16) is None
    finally:
        ops.set_kl_shard()
        ops.set_kl_fusion(True)

    w, ls2 = torch.nn.Parameter(torch.randn(4, 3)), torch.nn.Parameter(torch.randn(4, 3))
    cache = ops.FusedKLCache()
    cache.put((w, ls2), {"kind": 0})                # the path taken produced nothing
    assert cache.take((w, ls2)) is None
    s = torch.tensor(1.5)
    cache.put((w, ls2), {"kind": 0, "sum": s})
    assert cache.take((w, ls2)) is s and cache.take((w, ls2)) is None      # handed out once
    cache.put((w, ls2), {"kind": 0, "sum": s, "rows": (0, 2)})
    assert cache.take((w, ls2)) is None             # a shard's partial sum is not the layer's
    cache.put((w, ls2), {"kind": 0, "sum": s, "rows": (0, 2)})
    assert cache.take((w, ls2), rows=(0, 2)) is s
    cache.put((w, ls2), {"kind": 0, "sum": s})
    with torch.no_grad():
        ls2.add_(1.0)                               # optimiser-style in-place update
    assert cache.take((w, ls2)) is None
    cache.put((w, ls2), {"kind": 0, "sum": s})
    assert cache.take((torch.nn.Parameter(w.detach().clone()), ls2)) is None   # other tensor


def test_conv_abi_argument_contract_without_gpu():
    """cplxk_conv2d_fwd_g rejects inconsistent plane sets BEFORE any CUDA call (include/cplxk.h): complex
    input needs complex weights and output; an imaginary input plane with REAL weights and output is
    the variance-operand form (real conv of |x|^2, complex/base.py:100-117) and exists on the
    tensor-core path only -- no workspace / channels-last / variational -> CPLXK_ERR_UNSUPPORTED."""
    lib = _native.lib()
    buf = (ctypes.c_float * 64)()                       # never dereferenced: every call below is refused
    p = ctypes.cast(buf, ctypes.c_void_p)
    geom = (1, 4, 8, 8, 4, 3, 3, 1, 1, 0, 0, 1, 1, 1)   # B C H W O kh kw sh sw ph pw dh dw groups

    def call(x_im, w_im, y_im, ls2=None, channels_last=0, workspace=None, ws_bytes=0, math=_native.MATH_AUTO,
             geometry=geom):
        return lib.cplxk_conv2d_fwd_g(p, x_im, p, w_im, None, None, ls2, None, None, _native.NOISE_INJECT, 0, 0, 0,
                                      p, y_im, *geometry, _native.F32, math, channels_last, workspace, ws_bytes, None)

    assert call(p, p, None) == -1                       # complex input and weights, real output
    assert call(p, None, p) == -1                       # complex output from real weights
    assert call(None, p, None) == -1                    # real input, complex weights
    assert call(p, None, None) == _native.ERR_UNSUPPORTED                 # variance form without a workspace
    assert call(p, None, None, channels_last=1, workspace=p, ws_bytes=1 << 20) == _native.ERR_UNSUPPORTED
    assert call(p, None, None, ls2=p, workspace=p, ws_bytes=1 << 20) == _native.ERR_UNSUPPORTED
    assert call(p, None, None, math=_native.MATH_SIMT) == _native.ERR_UNSUPPORTED
    bad = list(geom); bad[5] = 0                        # kh = 0
    assert call(p, p, p, geometry=tuple(bad)) == -1
    bad = list(geom); bad[13] = 3                       # C % groups != 0
    assert call(p, p, p, geometry=tuple(bad)) == -1
    empty = list(geom); empty[0] = 0                    # empty batch: nothing to do, no pointer is looked at
    assert lib.cplxk_conv2d_fwd_g(None, None, None, None, None, None, None, None, None, 0, 0, 0, 0, None, None,
                                  *empty, _native.F32, 0, 0, None, 0, None) == 0
