from .sparsity import named_sparsity, sparsity
