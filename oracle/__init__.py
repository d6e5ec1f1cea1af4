"""CPU oracle: a NumPy restatement of the reference's rollout hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product
path: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it, and only as the
checker / reported CPU baseline.  ``dgppo_b200`` never imports this package.

Parity status: the env / graph / LiDAR / GAE functions are pinned against the
reference's own source, executed in this container under ``oracle/jaxshim``
(a NumPy stand-in for the jax API the reference imports; see
``tools/gen_golden_from_reference.py`` and ``tests/golden/``).  The arithmetic
that lives in un-vendored third-party packages (flax Dense/LayerNorm/GRUCell,
jraph segment ops, tfp tanh-Normal) is restated from their published
behaviour and is **parity unpinned** against a real JAX install (none exists
in this image; the reference ships no tests or golden vectors).

Arithmetic convention: fp32 everywhere, every add/mul/div/sqrt individually
rounded (no FMA contraction) - the semantics NumPy gives and the semantics
the CUDA env kernels are compiled to (``-fmad=false``).
"""
