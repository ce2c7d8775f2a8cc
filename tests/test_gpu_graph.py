"""CUDA-graph capture of the user-level step `y = layer(x); kl = sum(penalties(layer))`
(tests/test_relevance.py:62-73 shape of the loop): nothing on the path may synchronise, read a
device value on the host or consult torch's generator while capturing; replays are bit-identical
(the noise coordinates are baked into the captured launch)."""
import pytest
import torch

import cplxmodule_b200 as cb
from cplxmodule_b200 import cplx
from cplxmodule_b200.nn import relevance as rel

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _capture(fn):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s), torch.no_grad():
        for _ in range(3):                      # warm-up on the side stream, as torch asks
            fn()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.no_grad(), torch.cuda.graph(g):
        out = fn()
    return g, out


@pytest.mark.parametrize("noise", ["fast", "torch"])
@pytest.mark.parametrize("case", ["config1_real", "cplx_small", "cplx_persistent", "conv_vd"])
def test_step_is_graph_capturable_and_replays_bit_identically(case, noise):
    torch.manual_seed(3)
    if case == "config1_real":                  # BASELINE config 1: LinearVD 784 -> 256, batch 128
        layer, x = rel.LinearVD(784, 256).to(DEV).train(), torch.randn(128, 784, device=DEV)
    elif case == "cplx_small":
        layer, x = rel.CplxLinearVD(96, 40).to(DEV).train(), cplx.randn(50, 96, device=DEV)
    elif case == "cplx_persistent":             # pre-pass (+ fused KL) + persistent CTA-pair kernel, PDL edge
        layer, x = rel.CplxLinearARD(512, 384).to(DEV).train(), cplx.randn(300, 512, device=DEV)
    else:
        layer, x = rel.CplxConv2dVD(8, 16, 3, padding=1).to(DEV).train(), cplx.randn(2, 8, 12, 12, device=DEV)
    with torch.no_grad():
        layer.log_sigma2.uniform_(-8, -2)
    cb.set_noise_mode(noise)
    try:
        def step():
            y = layer(x)
            return y, sum(rel.penalties(layer))
        g, (y, kl) = _capture(step)
        planes = (lambda t: (t.real, t.imag)) if isinstance(y, cplx.Cplx) else (lambda t: (t,))
        g.replay()
        torch.cuda.synchronize()
        first = [p.clone() for p in planes(y)] + [kl.clone()]
        assert all(torch.isfinite(t).all() for t in first)
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        again = list(planes(y)) + [kl]
        for a, b in zip(first, again):
            assert torch.equal(a, b)
        # the captured step computes what the eager step computes: same mean (eval forward) and KL
        with torch.no_grad():
            kl_eager = sum(rel.penalties(layer))
            mu = layer.eval()(x)
        layer.train()
        assert abs(kl.item() - kl_eager.item()) <= 1e-5 * abs(kl_eager.item())
        mu0, y0 = planes(mu)[0], first[0]
        # noise of the right order, not garbage: |y - mu| <= 6 sd, sd^2 <= fan_in max|x|^2 e^-2
        fan_in = layer.log_sigma2[0].numel()
        xmax = max(p.abs().max().item() for p in planes(x))
        dev = (y0 - mu0).abs().max().item()
        assert 0.0 < dev < 6.0 * (fan_in * 2 * xmax ** 2 * 0.1354) ** 0.5
        # a changed input is picked up by the next replay (the graph reads x in place)
        x0 = planes(x)[0] if isinstance(x, cplx.Cplx) else x
        x0.mul_(2.0)
        g.replay()
        torch.cuda.synchronize()
        assert not torch.equal(planes(y)[0], first[0])
    finally:
        cb.set_noise_mode("torch")


def test_eager_generator_is_untouched_by_capture():
    """capturing must neither read nor advance torch's CUDA generator"""
    torch.manual_seed(5)
    layer, x = rel.LinearVD(64, 32).to(DEV).train(), torch.randn(16, 64, device=DEV)
    gen = torch.cuda.default_generators[torch.cuda.current_device()]
    before = gen.get_offset()
    g, _ = _capture(lambda: layer(x))
    # (the three warm-up steps ran eagerly and did advance it; the capture itself did not)
    after_warm = gen.get_offset()
    g.replay(); g.replay()
    torch.cuda.synchronize()
    assert gen.get_offset() == after_warm and after_warm > before
