"""TEST INFRASTRUCTURE ONLY.

CPU restatement (torch + scipy, the reference's own dependencies) of the algorithms on
cplxmodule's linear / conv / variational-dropout / KL hot path.  Nothing in the shipped
package ``cplxmodule_b200`` imports this; only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s CPU-baseline / ``--impl reference`` legs may.
"""
