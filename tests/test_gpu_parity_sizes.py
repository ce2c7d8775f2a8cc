"""Parity on exactly the paths and sizes bench.py times (BASELINE.json configs 2, 3, 5):

* the in-kernel torch-layout Philox noise at sizes where BOTH planes cross Philox slot and
  Box-Muller component boundaries (numel > 4 T, T = 256 * blocks of torch's launch) and runs
  wrap the thread index (N not a multiple of 8): fused == the same kernel fed with the draw
  `cplx.randn_like` / `torch.randn_like` makes on this device (cplxmodule/cplx.py:544-550,
  nn/relevance/complex/base.py:55, real/base.py:48), bit for bit;
* `CplxLinear` on bf16 planes through `lin_tc3_kernel<bf16>` (config 2, cplx.py:634-648) against
  the float64 oracle at 1e-2;
* `CplxLinearARD(8192, 8192)`, 8192 rows (config 5's per-GPU shard, complex/ard.py:39): sampled
  rows, and fused-KL == stand-alone kl_kernel == oracle.
"""
import pytest
import torch

import cplxmodule_b200 as cb
from cplxmodule_b200 import _native as nv
from cplxmodule_b200 import cplx, ops
from cplxmodule_b200.nn import CplxLinear
from cplxmodule_b200.nn.relevance import (CplxConv2dVD, CplxLinearARD, CplxLinearVD, LinearVD,
                                          penalties)
from oracle import cplx_oracle as orc
from tests.conftest import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _philox_T(numel):
    dev = torch.device("cuda", torch.cuda.current_device())
    gen, seed, offset, threads, inc = nv.philox_plan(dev, numel)
    return threads


def _fused_vs_inject(layer, x, draw):
    """(fused, inject, offsets equal): same seed, noise generated in the kernel vs injected."""
    gen = torch.cuda.default_generators[torch.cuda.current_device()]
    torch.manual_seed(4242)
    torch.randn(5, device=DEV)                     # start from a non-zero, non-multiple-of-4 offset
    state = gen.get_state()
    with torch.no_grad():
        fused = layer(x)
    off_fused = gen.get_offset()
    gen.set_state(state)
    eps = draw()
    off_ref = gen.get_offset()
    with torch.no_grad():
        inject = layer(x, eps=eps)
    return fused, inject, off_fused == off_ref


# (M, N, K): headline; ragged with N odd (every 8-run position relative to T occurs, rows are not
# 16-byte aligned); N % 8 == 4 (vector stores, runs straddling T by half)
@pytest.mark.parametrize("M,N,K", [(4096, 4096, 4096), (1100, 1111, 136), (1500, 900 + 4, 72)])
def test_cplx_fused_noise_bit_equal_beyond_first_slot(M, N, K):
    torch.manual_seed(3)
    T = _philox_T(2 * M * N)
    assert 2 * M * N > 4 * T, "both planes must cross slot AND component boundaries"
    layer = CplxLinearVD(K, N).to(DEV).train()
    with torch.no_grad():
        layer.log_sigma2.uniform_(-6, 1)
    x = cplx.randn(M, K, device=DEV)
    fused, inject, same_off = _fused_vs_inject(layer, x, lambda: cplx.randn(M, N, device=DEV))
    assert same_off
    assert torch.equal(fused.real, inject.real)
    assert torch.equal(fused.imag, inject.imag)
    # and the noise really is there (not all-zero eps on both sides)
    with torch.no_grad():
        mu = layer.eval()(x)
    layer.train()
    assert (fused.real - mu.real).abs().max().item() > 0


@pytest.mark.parametrize("M,N,K", [(1300, 1001, 64), (4096, 1024, 256)])
def test_real_fused_noise_bit_equal_beyond_first_slot(M, N, K):
    torch.manual_seed(4)
    T = _philox_T(M * N)
    assert M * N > 4 * T
    layer = LinearVD(K, N).to(DEV).train()
    with torch.no_grad():
        layer.log_sigma2.uniform_(-6, 1)
    x = torch.randn(M, K, device=DEV)
    fused, inject, same_off = _fused_vs_inject(layer, x, lambda: torch.randn(M, N, device=DEV))
    assert same_off and torch.equal(fused, inject)


@pytest.mark.parametrize("channels_last", [False, True])
def test_conv_fused_noise_bit_equal_beyond_first_slot(channels_last):
    """CplxConv2dVD with 2.8 M output elements (> 4 T = 1.2 M)."""
    torch.manual_seed(5)
    B, C, O, H = 4, 16, 16, 150
    layer = CplxConv2dVD(C, O, 3).to(DEV).train()
    with torch.no_grad():
        layer.log_sigma2.uniform_(-6, 1)
    x = cplx.randn(B, C, H, H, device=DEV)
    if channels_last:
        x = cplx.Cplx(x.real.contiguous(memory_format=torch.channels_last),
                      x.imag.contiguous(memory_format=torch.channels_last))
    Ho = H - 2
    assert 2 * B * O * Ho * Ho > 4 * _philox_T(2 * B * O * Ho * Ho)
    fused, inject, same_off = _fused_vs_inject(
        layer, x, lambda: cplx.randn(B, O, Ho, Ho, device=DEV))
    assert same_off
    assert torch.equal(fused.real, inject.real) and torch.equal(fused.imag, inject.imag)


# ------------------------------------------------------------------ config 2: CplxLinear bf16
def _bf16_case(M, N, K, bias=True, seed=0):
    torch.manual_seed(seed)
    bf = lambda t: t.to(DEV).bfloat16()
    x_re, x_im = bf(torch.randn(M, K)), bf(torch.randn(M, K))
    w_re, w_im = bf(torch.randn(N, K) / K ** 0.5), bf(torch.randn(N, K) / K ** 0.5)
    b_re, b_im = (bf(torch.randn(N)), bf(torch.randn(N))) if bias else (None, None)
    return x_re, x_im, w_re, w_im, b_re, b_im


@pytest.mark.parametrize("M,N,K", [(129, 8, 8), (130, 129, 72), (257, 200, 264), (300, 384, 1000),
                                   (513, 130, 4096), (1024, 1000, 520)])
@pytest.mark.parametrize("cplx_", [True, False])
def test_bf16_linear_lin_tc3_ragged(M, N, K, cplx_):
    """M > 128, K % 8 == 0: the persistent double-buffered CTA-pair kernel (fwd_lin3.cu) on
    bf16 planes -- partial row tiles, partial column tiles, K tails (TMA zero fill)."""
    x_re, x_im, w_re, w_im, b_re, b_im = _bf16_case(M, N, K, seed=M + N + K)
    c = lambda t: t.float().cpu().double()
    if cplx_:
        got = ops.cplx_linear(x_re, x_im, w_re, w_im, b_re, b_im)
        want = orc.cplx_linear(c(x_re), c(x_im), c(w_re), c(w_im), c(b_re), c(b_im))
        assert got[0].dtype == torch.bfloat16 and got[0].shape == (M, N)
        assert rel_err(got[0].float(), want[0]) < 1e-2 and rel_err(got[1].float(), want[1]) < 1e-2
    else:
        got = ops.real_linear(x_re, w_re, b_re)
        want = torch.nn.functional.linear(c(x_re), c(w_re), c(b_re))
        assert rel_err(got.float(), want) < 1e-2


def test_bf16_linear_config2_full_size_sampled_rows():
    """BASELINE.json configs[1]: CplxLinear 4096 -> 4096, bf16, batch 4096; 64 sampled rows
    against the float64 oracle (1e-2), module path (`CplxLinear.forward`)."""
    torch.manual_seed(22)
    B = D = 4096
    layer = CplxLinear(D, D).to(DEV).bfloat16()
    x = cplx.randn(B, D, device=DEV).to(torch.bfloat16)
    with torch.no_grad():
        y = layer(x)
    assert y.real.dtype == torch.bfloat16
    rows = torch.randperm(B)[:64].to(DEV)
    c = lambda t: t.detach().float().cpu().double()
    w, b = layer.weight, layer.bias
    want = orc.cplx_linear(c(x.real[rows]), c(x.imag[rows]), c(w.real), c(w.imag), c(b.real),
                           c(b.imag))
    assert rel_err(y.real[rows].float(), want[0]) < 1e-2
    assert rel_err(y.imag[rows].float(), want[1]) < 1e-2
    # size-independent property: conj symmetry  f(conj x; conj W, conj b) = conj f(x; W, b)
    with torch.no_grad():
        yc_re, yc_im = ops.cplx_linear(x.real, -x.imag, w.real, -w.imag, b.real, -b.imag)
    assert torch.equal(yc_re, y.real) and torch.equal(yc_im, -y.imag)


# -------------------------------------------------------------- config 5: CplxLinearARD 8192^2
def test_ard_config5_shard_8192():
    """BASELINE.json configs[4], one GPU's shard: CplxLinearARD(8192, 8192) on 8192 rows.  K = 8192
    is where the operand pre-pass stops holding a row in registers (second pass re-reads it)."""
    torch.manual_seed(55)
    B = D = 8192
    layer = CplxLinearARD(D, D).to(DEV).train()
    with torch.no_grad():
        layer.log_sigma2.uniform_(-12, 2)
    x = cplx.randn(B, D, device=DEV)
    eps = cplx.randn(B, D, device=DEV)
    with torch.no_grad():
        y = layer(x, eps=eps)
        kl_fused = sum(penalties(layer))              # by-product of the forward's pre-pass
        cb.set_kl_fusion(False)
        try:
            kl_alone = sum(penalties(layer))          # kl_kernel<CPLX_ARD>
        finally:
            cb.set_kl_fusion(True)
    rows = torch.randperm(B)[:32].to(DEV)
    c = lambda t: t.detach().cpu().double()
    w, b = layer.weight, layer.bias
    want = orc.cplx_linear_vd(c(x.real[rows]), c(x.imag[rows]), c(w.real), c(w.imag), c(b.real),
                              c(b.imag), c(layer.log_sigma2), c(eps.real[rows]), c(eps.imag[rows]))
    assert rel_err(y.real[rows], want[0]) < 1e-3 and rel_err(y.imag[rows], want[1]) < 1e-3
    kl_ref = orc.layer_penalty("cplx_ard", c(w.real), c(w.imag), c(layer.log_sigma2), "sum").item()
    assert abs(kl_fused.item() - kl_ref) / abs(kl_ref) < 1e-5
    assert abs(kl_alone.item() - kl_ref) / abs(kl_ref) < 1e-5
    assert abs(kl_fused.item() - kl_alone.item()) / abs(kl_ref) < 1e-6
    # fused torch-layout noise at this size (134 M normals) against the injected draw, sampled:
    # full-tensor equality costs another 1 GB; compare the first, a middle and the last rows
    del y, eps
    gen = torch.cuda.default_generators[torch.cuda.current_device()]
    torch.manual_seed(99)
    state = gen.get_state()
    with torch.no_grad():
        fused = layer(x)
    gen.set_state(state)
    eps = cplx.randn(B, D, device=DEV)
    with torch.no_grad():
        inject = layer(x, eps=eps)
    assert torch.equal(fused.real, inject.real) and torch.equal(fused.imag, inject.imag)


# ------------------------------------------- several waves of tiles per persistent CTA pair
@pytest.mark.parametrize("M,N,K", [(2500, 1300, 264), (1000, 3000 + 8, 512), (4096, 4096, 1024),
                                   (20000, 136, 256), (136 * 2, 20000, 256)])
@pytest.mark.parametrize("cplx_", [True, False])
def test_multi_wave_shapes_every_row(M, N, K, cplx_):
    """more tile pairs than one wave of the persistent grid, ragged in M and N, tall-skinny and
    short-wide: EVERY output row vs the float64 oracle (the full-size tests sample rows); the KL
    by-product of the pre-pass vs the stand-alone kernel; bit-reproducible."""
    torch.manual_seed(M + N + K)
    cls = CplxLinearVD if cplx_ else LinearVD
    layer = cls(K, N).to(DEV).train()
    with torch.no_grad():
        layer.log_sigma2.uniform_(-10, 0)
    x = cplx.randn(M, K, device=DEV) if cplx_ else torch.randn(M, K, device=DEV)
    eps = cplx.randn(M, N, device=DEV) if cplx_ else torch.randn(M, N, device=DEV)
    with torch.no_grad():
        y1 = layer(x, eps=eps)
        kl1 = sum(penalties(layer))
        y2 = layer(x, eps=eps)
        kl2 = sum(penalties(layer))
        cb.set_kl_fusion(False)
        try:
            kl_alone = sum(penalties(layer))
        finally:
            cb.set_kl_fusion(True)
    c = lambda t: t.detach().cpu().double()
    if cplx_:
        assert torch.equal(y1.real, y2.real) and torch.equal(y1.imag, y2.imag)
        w, b = layer.weight, layer.bias
        want = orc.cplx_linear_vd(c(x.real), c(x.imag), c(w.real), c(w.imag), c(b.real), c(b.imag),
                                  c(layer.log_sigma2), c(eps.real), c(eps.imag))
        assert rel_err(y1.real, want[0]) < 1e-3 and rel_err(y1.imag, want[1]) < 1e-3
    else:
        assert torch.equal(y1, y2)
        want = orc.real_linear_vd(c(x), c(layer.weight), c(layer.bias), c(layer.log_sigma2), c(eps))
        assert rel_err(y1, want) < 1e-3
    assert torch.equal(kl1, kl2)
    assert abs(kl1.item() - kl_alone.item()) <= 2e-6 * abs(kl_alone.item())


def test_repeated_calls_share_the_workspace_and_are_bit_stable():
    """40 back-to-back calls on one stream share the cached scratch workspace (and the KL
    reduction's ticket workspace): all results identical to the first."""
    torch.manual_seed(77)
    M, N, K = 3000, 2500, 512
    layer = CplxLinearVD(K, N).to(DEV).train()
    with torch.no_grad():
        layer.log_sigma2.uniform_(-10, 0)
    x, eps = cplx.randn(M, K, device=DEV), cplx.randn(M, N, device=DEV)
    with torch.no_grad():
        first = layer(x, eps=eps)
        kl0 = sum(penalties(layer))
        for _ in range(40):
            y = layer(x, eps=eps)
            kl = sum(penalties(layer))
            assert torch.equal(y.real, first.real) and torch.equal(y.imag, first.imag)
            assert torch.equal(kl, kl0)
