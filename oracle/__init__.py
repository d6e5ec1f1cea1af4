"""CPU oracle: a NumPy restatement of the reference's rollout hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product
path: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it, and only as the
checker / reported CPU baseline.  ``dgppo_b200`` never imports this package.

Parity status: every function is pinned against the reference's own source,
executed in this container under ``oracle/jaxshim`` / ``oracle/flaxshim``
(NumPy stand-ins for the jax / flax / jraph / tfp APIs the reference imports;
see ``tools/gen_golden_from_reference.py`` and ``tests/golden/``): env step,
graph, LiDAR, reset graph bit for bit; GAE and the policy / Vh / Vl forward
(the reference's gnn.py, mlp.py, rnn.py, policy.py, value.py,
distribution.py) to fp32 rounding.  What stays **parity unpinned** is the
arithmetic INSIDE the un-vendored third-party layers (flax Dense / LayerNorm /
GRUCell, jraph segment ops, tfp Normal / Tanh / log_ndtr), restated from their
published behaviour in both the oracle and the stand-ins, and real-XLA
rounding (FMA contraction, libm): no JAX install exists in this image and the
reference ships no tests or golden vectors.

Arithmetic convention: fp32 everywhere, every add/mul/div/sqrt individually
rounded (no FMA contraction) - the semantics NumPy gives and the semantics
the CUDA env kernels spell out (``__fadd_rn`` / ``__fmul_rn`` / ``__fdiv_rn`` / ``__fsqrt_rn``,
which nvcc never contracts into FMAs).
"""
