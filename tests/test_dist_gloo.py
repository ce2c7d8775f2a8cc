"""world_size-2 gloo test (CPU) of the multi-GPU layout: row-sharded KL + ONE all-reduce,
batch rows sharded with no data-path collective.  The shard kernel is replaced by the
oracle (allowed in tests) so the host-side sharding/reduction logic is what is exercised."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cplxmodule_b200.distributed import row_shard, sharded_penalties
from cplxmodule_b200.nn.relevance import CplxLinearARD, CplxLinearVD, LinearVD
from oracle import cplx_oracle as orc


def test_row_shard_partitions():
    for n in (0, 1, 7, 8, 4096, 8191):
        for world in (1, 2, 3, 8):
            spans = [row_shard(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _oracle_partial(mod, lo, hi):
    w = mod.weight
    kind = {CplxLinearVD: "cplx_vd", CplxLinearARD: "cplx_ard", LinearVD: "real_vd"}[type(mod)]
    if kind.startswith("cplx"):
        return orc.layer_penalty(kind, w.real[lo:hi].detach(), w.imag[lo:hi].detach(),
                                 mod.log_sigma2[lo:hi].detach())
    return orc.layer_penalty(kind, w[lo:hi].detach(), None, mod.log_sigma2[lo:hi].detach())


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)                       # replicated parameters
        net = torch.nn.Sequential(CplxLinearVD(24, 13), CplxLinearARD(13, 9), LinearVD(9, 5))
        for m in net:
            with torch.no_grad():
                m.log_sigma2.uniform_(-12, 2)
        names, vec = sharded_penalties(net, partial_fn=_oracle_partial)
        full = torch.stack([_oracle_partial(m, 0, m.log_sigma2.shape[0]) for m in net])
        # batch sharding: each rank owns its rows, outputs need no exchange
        torch.manual_seed(100)
        x = torch.randn(10, 9)
        lo, hi = row_shard(10, rank, world)
        w, b, ls2 = net[2].weight.detach(), net[2].bias.detach(), net[2].log_sigma2.detach()
        eps = torch.randn(10, 5)
        mine = orc.real_linear_vd(x[lo:hi], w, b, ls2, eps[lo:hi])
        whole = orc.real_linear_vd(x, w, b, ls2, eps)
        out.put((rank, names, vec.tolist(), full.tolist(),
                 bool(torch.equal(mine, whole[lo:hi]))))
    finally:
        dist.destroy_process_group()


def test_sharded_kl_allreduce_world2():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    results = [out.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    vecs = {}
    for rank, names, vec, full, rows_ok in results:
        assert names == ["0", "1", "2"] and rows_ok
        vecs[rank] = vec
        for a, b in zip(vec, full):
            assert abs(a - b) <= 1e-5 * abs(b)
    assert vecs[0] == vecs[1]                      # every rank holds the same reduced vector


def test_single_process_degenerates_to_full_sum():
    torch.manual_seed(1)
    net = torch.nn.Sequential(CplxLinearVD(6, 4))
    names, vec = sharded_penalties(net, partial_fn=_oracle_partial)
    assert names == ["0"] and torch.allclose(vec[0], _oracle_partial(net[0], 0, 4))


# ---------------------------------------------------------------- gradients
def _oracle_partial_grad(mod, lo, hi):
    """differentiable partial (torch autograd over the oracle): what ops.kl provides on the GPU"""
    w = mod.weight
    kind = {CplxLinearVD: "cplx_vd", CplxLinearARD: "cplx_ard", LinearVD: "real_vd"}[type(mod)]
    if kind.startswith("cplx"):
        return orc.layer_penalty(kind, w.real[lo:hi], w.imag[lo:hi], mod.log_sigma2[lo:hi])
    return orc.layer_penalty(kind, w[lo:hi], None, mod.log_sigma2[lo:hi])


def _grad_worker(rank, world, port, out, grad_reduce):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        net = torch.nn.Sequential(CplxLinearARD(12, 7), LinearVD(7, 5)).double()
        for m in net:
            with torch.no_grad():
                m.log_sigma2.uniform_(-6, 2)
        params = [p for p in net.parameters()]
        # reference: every rank differentiates the FULL KL (no sharding, no collective)
        full = sum(_oracle_partial_grad(m, 0, m.log_sigma2.shape[0]) for m in net)
        want = torch.autograd.grad(full, params, allow_unused=True)
        # sharded: each rank differentiates its rows, then the data-parallel gradient reduction
        _, vec = sharded_penalties(net, partial_fn=_oracle_partial_grad, grad_reduce=grad_reduce)
        got = torch.autograd.grad(vec.sum(), params, allow_unused=True)
        ok = abs(vec.sum().item() - full.item()) <= 1e-9 * abs(full.item())
        for g, w in zip(got, want):
            if w is None:
                ok = ok and (g is None or float(g.abs().max()) == 0.0)
                continue
            g = torch.zeros_like(w) if g is None else g.clone()
            dist.all_reduce(g, op=dist.ReduceOp.SUM)
            if grad_reduce == "mean":
                g /= world                                   # DDP's default averaging
            ok = ok and torch.allclose(g, w, rtol=1e-9, atol=1e-12)
        out.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("grad_reduce", ["mean", "sum"])
def test_sharded_kl_gradients_world2(grad_reduce):
    """after the replicas' gradients are combined (mean = DDP default, or sum) every rank holds the
    gradient of the FULL KL: the all-reduce in sharded_penalties is differentiable"""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_grad_worker, args=(r, 2, port, out, grad_reduce)) for r in range(2)]
    for p in procs:
        p.start()
    results = [out.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok in results)
