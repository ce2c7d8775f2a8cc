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


def _default_partial(mod, lo, hi):
    w_re, w_im = _layer_planes(mod)
    if hi <= lo:
        return torch.zeros((), dtype=torch.float32, device=w_re.device)
    return ops.kl(mod._kl_kind, w_re[lo:hi], None if w_im is None else w_im[lo:hi],
                  mod.log_sigma2[lo:hi], "sum").float()


def sharded_penalties(module, group=None, partial_fn=None):
    """``sum``-reduced penalties of every variational layer, each computed on this rank's row
    shard and combined with a single all-reduce.  Returns ``(names, tensor[n_layers])``.

    ``partial_fn(mod, lo, hi) -> 0-d tensor`` overrides the shard kernel (tests use it to
    exercise the sharding logic on CPU/gloo)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    partial_fn = partial_fn or _default_partial
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
        dist.all_reduce(vec, op=dist.ReduceOp.SUM, group=group)
    return names, vec
