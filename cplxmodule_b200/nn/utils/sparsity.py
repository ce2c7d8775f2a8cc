"""Sparsity bookkeeping over a network (reference: ``cplxmodule/nn/utils/sparsity.py``): every
layer with a ``sparsity(**kwargs)`` method reports ``[(id(parameter), n_zeros), ...]``."""


def named_sparsity(module, prefix="", **kwargs):
    """Yield ``(parameter name, (n_zeros, n_total))`` for every parameter of the network;
    parameters of layers without sparsity information count as dense."""
    n_dropped = {}
    for _, mod in module.named_modules(prefix=prefix):
        fn = getattr(mod, "sparsity", None)
        if callable(fn):
            n_dropped.update(fn(**kwargs))
    ignore = set()
    for name, mod in module.named_modules(prefix=prefix):
        for par in getattr(mod, "__sparsity_ignore__", ()):
            ignore.add(name + ("." if name else "") + par)
    for name, par in module.named_parameters(prefix=prefix):
        if name in ignore:
            continue
        yield name, (n_dropped.get(id(par), 0.0), par.numel())


def sparsity(module, **kwargs):
    """Overall fraction of zeroed parameters."""
    pairs = [v for _, v in named_sparsity(module, **kwargs)]
    total = float(sum(n for _, n in pairs))
    return sum(z for z, _ in pairs) / max(total, 1.0)
