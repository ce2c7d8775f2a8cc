"""Multi-GPU layout of the path: one process per GPU, batch rows sharded across ranks with
NO data-path collective (the forward of different samples is independent given replicated
parameters); the KL term depends on parameters only, so it is row-sharded across ranks
(each rank streams 1/G of the weight rows) and finished by ONE all-reduce of a small vector
(one scalar per variational layer) over NCCL / NVLink.  The reference has no counterpart
(SURVEY.md section 5)."""
import torch
import torch.distributed as dist

from . import ops
from .nn.relevance.base import BaseARD


def row_shard(n_rows, rank, world):
    """Contiguous block partition of ``n_rows`` rows: first ``n_rows % world`` ranks get one extra."""
    base, extra = divmod(n_rows, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def _layer_planes(mod):
    w = mod.weight
    if hasattr(w, "real") and not isinstance(w, torch.Tensor):
        return w.real, w.imag
    return w, None


def _default_partial(mod, lo, hi, stream=None):
    """Partial KL sum of weight rows [lo, hi): the by-product of this step's forward when it was
    asked for that shard (``ops.set_kl_shard``; no extra pass over the parameters), else the
    stand-alone kernel on the row slice."""
    w_re, w_im = _layer_planes(mod)
    if hi <= lo:
        return torch.zeros((), dtype=torch.float32, device=w_re.device)
    cache = mod.__dict__.get("_kl_cache")
    wants_grad = torch.is_grad_enabled() and mod.log_sigma2.requires_grad
    if cache is not None and not wants_grad:     # the by-product carries no autograd graph
        params = (w_re, mod.log_sigma2) if w_im is None else (w_re, w_im, mod.log_sigma2)
        pre, event = cache.take(params, rows=(lo, hi), with_event=True)
        if pre is not None:
            if stream is not None and event is not None:
                stream.wait_event(event)     # final right after the pre-pass: do not wait for the GEMM
                pre.record_stream(stream)
            return pre
    return ops.kl(mod._kl_kind, w_re[lo:hi], None if w_im is None else w_im[lo:hi],
                  mod.log_sigma2[lo:hi], "sum").float()


class _AllReduceSum(torch.autograd.Function):
    """SUM all-reduce of the per-layer partial KL sums whose backward hands every rank's partial
    ``grad_scale`` times the upstream gradient.

    KL = sum over ranks of partial_r and rank r only differentiates ITS rows, so after the
    backward the gradient of the KL term lives on rank r for rows of r and is zero elsewhere; the
    data-parallel gradient reduction then puts it together.  With a MEAN reduction (DDP's
    default) each rank must contribute ``world`` times its part for the mean over ranks to be the
    full KL gradient (``grad_reduce="mean"`` -> ``grad_scale = world``); with a SUM reduction
    ``grad_scale = 1``.  ``dist.all_reduce`` itself has no autograd formula: called directly it
    would silently give every rank the gradient of its own rows only."""

    @staticmethod
    def forward(ctx, vec, group, grad_scale):
        ctx.grad_scale = grad_scale
        out = vec.detach().clone()
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
        return out

    @staticmethod
    def backward(ctx, grad):
        return grad * ctx.grad_scale, None, None


def sharded_penalties(module, group=None, partial_fn=None, stream=None, grad_reduce="mean"):
    """``sum``-reduced penalties of every variational layer, each computed on this rank's row
    shard and combined with a single all-reduce.  Returns ``(names, tensor[n_layers])``.

    Differentiable: ``loss = nll + klw * sharded_penalties(model)[1].sum()`` trains correctly
    under data parallelism when ``grad_reduce`` names how the gradients of the replicas are
    combined afterwards -- ``"mean"`` (``DistributedDataParallel``'s default averaging) or
    ``"sum"``; see ``_AllReduceSum``.

    ``partial_fn(mod, lo, hi) -> 0-d tensor`` overrides the shard kernel (tests use it to
    exercise the sharding logic on CPU/gloo).  ``stream``: the (current) side stream the call is
    made under; partial sums that come from the forward's pre-pass are then waited for through
    their own event instead of the whole forward, so the collective overlaps the GEMM."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if partial_fn is None:
        partial_fn = lambda mod, lo, hi: _default_partial(mod, lo, hi, stream)  # noqa: E731
    names, parts = [], []
    for name, mod in module.named_modules():
        if not isinstance(mod, BaseARD) or not hasattr(mod, "log_sigma2"):
            continue
        lo, hi = row_shard(mod.log_sigma2.shape[0], rank, world)
        names.append(name)
        parts.append(partial_fn(mod, lo, hi).reshape(()))
    if not parts:
        return names, torch.zeros(0)
    vec = torch.stack(parts)
    if world > 1:
        if grad_reduce not in ("mean", "sum"):
            raise ValueError("grad_reduce must be 'mean' (DDP averaging) or 'sum'")
        if torch.is_grad_enabled() and vec.requires_grad:
            vec = _AllReduceSum.apply(vec, group, float(world) if grad_reduce == "mean" else 1.0)
        else:
            dist.all_reduce(vec, op=dist.ReduceOp.SUM, group=group)
    return names, vec
