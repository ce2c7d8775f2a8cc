"""fp32 planes on the kind::f16 tensor-core path: per-row power-of-two scaled fp16 operands
(pre-pass of fwd_tc3.cu), persistent CTA-pair kernel, and the KL sum fused into the pre-pass.
All comparisons are against the float64 oracle on the same inputs and injected noise."""
import os

import pytest
import torch

import cplxmodule_b200 as cb
from cplxmodule_b200 import cplx, ops
from cplxmodule_b200.nn.relevance import CplxLinearARD, CplxLinearVD, LinearARD, LinearVD, penalties
from oracle import cplx_oracle as orc
from tests.conftest import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = 1e-3   # north_star: 1e-3 relative (max-norm) for fp32 planes


def _case(M, N, K, seed, row_scale_x=None, row_scale_w=None):
    g = torch.Generator().manual_seed(seed)
    x_re, x_im = torch.randn(M, K, generator=g), torch.randn(M, K, generator=g)
    w_re = torch.randn(N, K, generator=g) / K ** 0.5
    w_im = torch.randn(N, K, generator=g) / K ** 0.5
    if row_scale_x is not None:
        x_re, x_im = x_re * row_scale_x[:, None], x_im * row_scale_x[:, None]
    if row_scale_w is not None:
        w_re, w_im = w_re * row_scale_w[:, None], w_im * row_scale_w[:, None]
    b_re, b_im = torch.randn(N, generator=g), torch.randn(N, generator=g)
    ls2 = torch.empty(N, K).uniform_(-12, 2, generator=g)
    eps_re, eps_im = torch.randn(M, N, generator=g), torch.randn(M, N, generator=g)
    return x_re, x_im, w_re, w_im, b_re, b_im, ls2, eps_re, eps_im


def _run(case, cplx_=True):
    x_re, x_im, w_re, w_im, b_re, b_im, ls2, eps_re, eps_im = case
    d = lambda t: t.to(DEV)
    if cplx_:
        want = orc.cplx_linear_vd(*(t.double() for t in case))
        got = ops.cplx_linear_vd(d(x_re), d(x_im), d(w_re), d(w_im), d(b_re), d(b_im), d(ls2),
                                 eps=(d(eps_re), d(eps_im)))
        return got, want
    want = orc.real_linear_vd(x_re.double(), w_re.double(), b_re.double(), ls2.double(), eps_re.double())
    got = ops.real_linear_vd(d(x_re), d(w_re), d(b_re), d(ls2), eps=d(eps_re))
    return (got,), (want,)


@pytest.fixture(params=["persistent"])
def kernel(request):
    """one kernel serves the scaled-fp16 path (fwd_tc3.cu); the tile-per-cluster generation is gone"""
    yield request.param


@pytest.mark.parametrize("M,N,K", [(129, 8, 8), (256, 128, 64), (300, 384, 1000), (513, 130, 264),
                                   (1024, 640, 4096), (200, 72, 8200), (2000, 1000, 16384 + 8)])
@pytest.mark.parametrize("cplx_", [True, False])
def test_shapes(M, N, K, cplx_, kernel):
    """ragged tiles in M and N, K tails past the register-cached 8192 columns, real and complex"""
    got, want = _run(_case(M, N, K, seed=M + N + K), cplx_)
    for a, b in zip(got, want):
        assert rel_err(a, b) < TOL


def test_rows_of_wildly_different_magnitude(kernel):
    """every row carries its own power-of-two scale: a row of 1e-9 next to a row of 1e+9 keeps
    full precision RELATIVE TO ITS OWN outputs, which a per-tensor fp16 scale could not do
    (|x|^2 . exp(log_sigma2) must stay inside fp32, as in the reference, hence not wider)"""
    M, N, K = 384, 256, 512
    sx = torch.logspace(-9, 9, M)[torch.randperm(M, generator=torch.Generator().manual_seed(1))]
    sw = torch.logspace(-12, 12, N)[torch.randperm(N, generator=torch.Generator().manual_seed(2))]
    x_re, x_im, w_re, w_im, b_re, b_im, ls2, eps_re, eps_im = _case(M, N, K, 5, sx, sw)
    d = lambda t: t.to(DEV)
    mu = orc.cplx_linear(x_re.double(), x_im.double(), w_re.double(), w_im.double())
    # noise off (eps = 0) and no bias: the output is the mean GEMM alone, checked row by row and
    # column by column against the scale of that row / column
    zero = torch.zeros(M, N)
    got = ops.cplx_linear_vd(d(x_re), d(x_im), d(w_re), d(w_im), None, None, d(ls2), eps=(d(zero), d(zero)))
    for g, w in zip(got, mu):
        err = (g.double().cpu() - w).abs()
        scale = sx.double()[:, None] * sw.double()[None, :]
        assert float((err / scale).max()) < 4 * TOL * float((w.abs() / scale).max())
        assert torch.isfinite(g).all()


def test_zero_rows_and_tiny_values(kernel):
    M, N, K = 260, 136, 128
    case = list(_case(M, N, K, 11))
    case[0][3].zero_(); case[1][3].zero_()          # an all-zero input row
    case[2][7].zero_(); case[3][7].zero_()          # an all-zero weight row
    case[0][5] *= 1e-42; case[1][5] *= 1e-42        # a subnormal input row
    got, want = _run(case)
    for a, b in zip(got, want):
        assert torch.isfinite(a).all() and rel_err(a, b) < TOL
    # outputs of the zero weight row: bias + noise only
    assert rel_err(got[0][:, 7], want[0][:, 7]) < TOL


def test_inf_and_nan_propagate_like_the_reference(kernel):
    M, N, K = 256, 128, 64
    case = list(_case(M, N, K, 13))
    case[0][2, 3] = float("inf")
    case[2][5, 1] = float("nan")
    got, _ = _run(case)
    re = got[0].cpu()
    assert not torch.isfinite(re[2]).all()           # the row with the inf
    assert torch.isnan(re[:, 5]).all()               # the column of the nan weight
    keep = torch.ones(M, dtype=torch.bool); keep[2] = False
    cols = torch.ones(N, dtype=torch.bool); cols[5] = False
    assert torch.isfinite(re[keep][:, cols]).all()   # nothing else is contaminated


def test_matches_tf32_path_and_exact_fp32_kernel():
    """three independent implementations of the same forward on the device agree"""
    case = _case(512, 256, 1024, 17)
    got16, want = _run(case)
    ops.set_math_mode("tf32")          # tf32 operands (cplxk_math: CPLXK_MATH_TENSOR_TF32)
    try:
        got32, _ = _run(case)
    finally:
        ops.set_math_mode("auto")
    ops.set_math_mode("simt")
    try:
        exact, _ = _run(case)
    finally:
        ops.set_math_mode("auto")
    for a, b, c, w in zip(got16, got32, exact, want):
        assert rel_err(c, w) < 2e-5
        assert rel_err(a, w) < TOL and rel_err(b, w) < TOL
        # both carry an 11-bit significand in the mean GEMM; the scaled-fp16 path additionally
        # rounds the variance operands to bf16, the tf32 path keeps them tf32
        assert rel_err(a, w) < 5e-4


@pytest.mark.parametrize("cls,kind", [(CplxLinearVD, "cplx_vd"), (CplxLinearARD, "cplx_ard"),
                                      (LinearVD, "real_vd"), (LinearARD, "real_ard")])
@pytest.mark.parametrize("reduction", ["sum", "mean"])
def test_fused_kl_equals_standalone_pass(cls, kind, reduction):
    """penalties() after a training forward returns the pre-pass by-product: same value as the
    stand-alone KL kernel and as the float64 oracle; used once, invalidated by parameter updates"""
    torch.manual_seed(3)
    layer = cls(264, 200).to(DEV).train()
    with torch.no_grad():
        layer.log_sigma2.uniform_(-12, 4)
    is_cplx = kind.startswith("cplx")
    x = cplx.randn(300, 264, device=DEV) if is_cplx else torch.randn(300, 264, device=DEV)
    with torch.no_grad():
        cb.set_kl_fusion(False)
        layer(x)
        alone = float(sum(penalties(layer, reduction=reduction)))
        cb.set_kl_fusion(True)
        layer(x)
        assert layer._kl_cache._entry is not None            # the forward produced it
        fused = float(sum(penalties(layer, reduction=reduction)))
        assert layer._kl_cache._entry is None                # handed out once
        again = float(sum(penalties(layer, reduction=reduction)))
    w = layer.weight
    w_re, w_im = (w.real, w.imag) if is_cplx else (w, None)
    c = lambda t: None if t is None else t.detach().double().cpu()
    want = float(orc.layer_penalty(kind, c(w_re), c(w_im), c(layer.log_sigma2), reduction))
    assert abs(fused - alone) <= 1e-6 * abs(alone)
    assert again == alone
    assert abs(fused - want) <= 1e-4 * abs(want)
    # a parameter update between forward and penalties() must not return the stale sum
    with torch.no_grad():
        layer(x)
        layer.log_sigma2.add_(1.0)
        fresh = float(sum(penalties(layer, reduction=reduction)))
        cb.set_kl_fusion(False)
        ref = float(sum(penalties(layer, reduction=reduction)))
        cb.set_kl_fusion(True)
    assert fresh == ref and abs(fresh - alone) > 1e-3 * abs(alone)


def test_fused_kl_data_edit_is_caught_on_the_device():
    """a write through `.data` bumps no version counter: the hand-out is checked against a device
    fingerprint of the parameters -> NaN for that step (never the stale sum), and the next call
    into the cache raises; `set_kl_fusion("unchecked")` skips the check"""
    torch.manual_seed(5)
    layer = CplxLinearVD(264, 200).to(DEV).train()
    x = cplx.randn(300, 264, device=DEV)
    with torch.no_grad():
        layer(x)
        ok = sum(penalties(layer))
        assert torch.isfinite(ok)
        layer(x)
        v = layer.log_sigma2._version
        layer.log_sigma2.data.add_(1.0)
        assert layer.log_sigma2._version == v            # invisible to the host-side key
        stale = sum(penalties(layer))
        torch.cuda.synchronize()
        assert torch.isnan(stale)
        with pytest.raises(RuntimeError, match=r"\.data"):
            layer(x)
        # the flag is cleared by the raise; the layer works again and sees the edited parameters
        layer(x)
        fresh = float(sum(penalties(layer)))
        cb.set_kl_fusion(False)
        ref = float(sum(penalties(layer)))
        cb.set_kl_fusion("unchecked")
        layer(x)
        unchecked = float(sum(penalties(layer)))
        cb.set_kl_fusion(True)
    assert abs(fresh - ref) <= 1e-6 * abs(ref) and abs(unchecked - ref) <= 1e-6 * abs(ref)
    assert abs(fresh - float(ok)) > 1e-3 * abs(ref)


def test_fused_kl_keeps_gradients():
    torch.manual_seed(4)
    a = CplxLinearVD(256, 136).to(DEV).train()
    b = CplxLinearVD(256, 136).to(DEV).train()
    b.load_state_dict(a.state_dict())
    x = cplx.randn(260, 256, device=DEV)
    eps = cplx.randn(260, 136, device=DEV)
    outs = []
    for layer, fuse in ((a, True), (b, False)):
        cb.set_kl_fusion(fuse)
        y = layer(x, eps=eps)
        loss = (y.real ** 2 + y.imag ** 2).mean() + 1e-3 * sum(penalties(layer))
        loss.backward()
        outs.append(float(loss))
    cb.set_kl_fusion(True)
    assert abs(outs[0] - outs[1]) <= 1e-6 * abs(outs[1])
    for pa, pb in zip(a.parameters(), b.parameters()):
        assert torch.allclose(pa.grad, pb.grad, rtol=1e-5, atol=1e-7)


def test_row_sharded_fused_kl_partials_add_up():
    """multi-GPU layout on one device: each 'rank' asks its forward for the KL of its block of
    weight rows; the partial sums equal the stand-alone kernel on the slices and add up to the
    whole-layer KL (what the all-reduce computes)"""
    from cplxmodule_b200.distributed import _default_partial, row_shard
    torch.manual_seed(5)
    layer = CplxLinearVD(128, 203).to(DEV).train()
    with torch.no_grad():
        layer.log_sigma2.uniform_(-10, 3)
    x = cplx.randn(300, 128, device=DEV)
    world, parts = 3, []
    side = torch.cuda.Stream()
    with torch.no_grad():
        whole = float(ops.kl(layer._kl_kind, layer.weight.real, layer.weight.imag, layer.log_sigma2, "sum"))
        for rank in range(world):
            cb.set_kl_shard(rank, world)
            try:
                layer(x)
                lo, hi = row_shard(203, rank, world)
                assert layer._kl_cache._entry[2] == (lo, hi)
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    part = _default_partial(layer, lo, hi, side)
                torch.cuda.current_stream().wait_stream(side)
                assert layer._kl_cache._entry is None          # consumed
                w = layer.weight
                ref = float(ops.kl(layer._kl_kind, w.real[lo:hi], w.imag[lo:hi], layer.log_sigma2[lo:hi], "sum"))
                assert abs(float(part) - ref) <= 1e-6 * abs(ref)
                parts.append(float(part))
                # the whole-layer API must not hand out a shard's partial sum
                layer(x)
                assert abs(float(sum(penalties(layer))) - whole) <= 1e-6 * abs(whole)
            finally:
                cb.set_kl_shard()
    assert abs(sum(parts) - whole) <= 1e-6 * abs(whole)


def test_full_size_exact_properties():
    """BASELINE headline size (B = d = 4096), properties that hold BIT FOR BIT on the scaled-fp16
    path because every row carries a power-of-two scale of its own:
      * rescaling input rows / weight rows by powers of two rescales the mean exactly;
      * permuting the input rows permutes the output rows;
    and the KL by-product of the forward equals the stand-alone KL pass."""
    B = D = 4096
    torch.manual_seed(9)
    layer = CplxLinearVD(D, D).to(DEV).train()
    with torch.no_grad():
        layer.log_sigma2.uniform_(-12, 0)
    w = layer.weight
    x = cplx.randn(B, D, device=DEV)
    zero = torch.zeros(B, D, device=DEV)
    run = lambda xr, xi, wr, wi: ops.cplx_linear_vd(xr, xi, wr, wi, None, None, layer.log_sigma2,
                                                     eps=(zero, zero))
    with torch.no_grad():
        y0 = run(x.real, x.imag, w.real, w.imag)
        # per-row powers of two on x (2^-20 .. 2^20) and on W (2^-8 .. 2^8)
        ex = torch.randint(-20, 21, (B, 1), device=DEV).float()
        ew = torch.randint(-8, 9, (D, 1), device=DEV).float()
        sx, sw = torch.exp2(ex), torch.exp2(ew)
        y1 = run(x.real * sx, x.imag * sx, w.real * sw, w.imag * sw)
        for a, b in zip(y0, y1):
            assert torch.equal(a * sx * sw.t(), b)
        perm = torch.randperm(B, device=DEV)
        y2 = run(x.real[perm], x.imag[perm], w.real, w.imag)
        for a, b in zip(y0, y2):
            assert torch.equal(a[perm], b)
        # sampled rows against the float64 oracle
        rows = torch.arange(0, B, 97, device=DEV)
        c = lambda t: t.detach().double().cpu()
        want = orc.cplx_linear(c(x.real[rows]), c(x.imag[rows]), c(w.real), c(w.imag))
        assert rel_err(y0[0][rows], want[0]) < TOL and rel_err(y0[1][rows], want[1]) < TOL
        # fused KL == stand-alone KL at full size
        layer(x)
        fused = float(sum(penalties(layer)))
        alone = float(sum(penalties(layer)))
        assert abs(fused - alone) <= 1e-6 * abs(alone)
