"""TEST INFRASTRUCTURE (see oracle/__init__.py): stand-ins that let the reference's own
`DGPPO.update` / `update_inner` / `update_Vl` / `update_Vh` / `update_policy`
(dgppo/algo/dgppo.py:136-321, dgppo/algo/informarl.py:357-457) EXECUTE under the NumPy
shims of oracle/jaxshim.py + oracle/flaxshim.py, so that the pre-pass intermediates and
the three loss functions of the PPO update can be pinned against the reference's source
instead of a restatement.

What is added on top of flaxshim.install():
  * `optax`: adam / apply_if_finite (inert descriptors), l2_loss, piecewise_constant_schedule,
    constant_schedule;
  * `flax.training.train_state.TrainState`: create / apply_gradients (parameters are left
    untouched: no autodiff exists here) / replace;
  * `jax.value_and_grad`, `jax.grad`: evaluate the function, RECORD (function, parameters) in
    `CAPTURED`, and return an all-zero gradient.  The recorded closures are the reference's own
    `get_loss_` functions; the generator evaluates them at perturbed parameters (finite
    differences) to pin the gradients of algo/update.py;
  * `jax.vmap` with keyword arguments (dgppo.py:232-237 maps `compute_dec_ocp_gae` by keyword);
  * `jax.random.fold_in`; `Independent.entropy` (tfd.Independent sums the event axis);
  * `trace_constants()`: a context in which `np.random.randint` returns one fixed value - what
    `jax.jit` makes of the Python-side seed in TanhTransformedDistribution.entropy
    (distribution.py:37-43: evaluated once at trace time, so every sample of every update shares
    one (n_agents, action_dim) draw).
"""
from __future__ import annotations

import contextlib
import sys

import numpy as np

from . import flaxshim
from . import jaxshim as J

F = np.float32
CAPTURED = []          # (function, params) per value_and_grad / grad evaluation, in call order


def _value_and_grad(f, has_aux=False, argnums=0, **kw):
    def run(params, *a, **k):
        out = f(params, *a, **k)
        CAPTURED.append((f, params))
        return out, J.tree_map(lambda x: np.zeros_like(np.asarray(x)), params)
    return run


def _grad(f, has_aux=False, argnums=0, **kw):
    vg = _value_and_grad(f, has_aux)

    def run(params, *a, **k):
        out, g = vg(params, *a, **k)
        return (g, out[1]) if has_aux else g
    return run


class TrainState:
    """flax.training.train_state.TrainState, minus the optimiser (informarl.py:133-137)."""

    def __init__(self, step, apply_fn, params, tx):
        self.step, self.apply_fn, self.params, self.tx = step, apply_fn, params, tx

    @classmethod
    def create(cls, *, apply_fn, params, tx, **kw):
        return cls(0, apply_fn, params, tx)

    def apply_gradients(self, *, grads, **kw):
        return TrainState(self.step + 1, self.apply_fn, self.params, self.tx)

    def replace(self, **kw):
        d = dict(step=self.step, apply_fn=self.apply_fn, params=self.params, tx=self.tx)
        d.update(kw)
        return TrainState(**d)


def _piecewise(init_value, boundaries_and_scales):
    items = sorted(boundaries_and_scales.items())

    def fn(step):
        v = init_value
        for b, s in items:
            if int(step) >= b:
                v = v * s
        return v
    return fn


def _vmap_kw(fn, in_axes=0, out_axes=0):
    base = J.vmap

    def mapped(*args, **kwargs):
        if not kwargs:
            return base(fn, in_axes, out_axes)(*args)
        names = list(kwargs)

        def g(*a):
            return fn(*a[:len(args)], **dict(zip(names, a[len(args):])))
        axes = (list(in_axes) if isinstance(in_axes, (tuple, list)) else [in_axes] * len(args)) + [0] * len(names)
        return base(g, axes, out_axes)(*args, *[kwargs[k] for k in names])
    return mapped


def _independent_entropy(self, **kw):
    e = np.asarray(self.distribution.entropy(**kw))
    return J._narrow(e.sum(axis=tuple(range(-self.nd, 0)), dtype=flaxshim.F))


@contextlib.contextmanager
def trace_constants(seed_value: int = 4242):
    """np.random.randint -> one fixed value, as under jax.jit (see the module docstring)."""
    orig = np.random.randint
    np.random.randint = lambda *a, **k: seed_value
    try:
        yield seed_value
    finally:
        np.random.randint = orig


@contextlib.contextmanager
def x64():
    """Evaluate under the stand-ins in float64 (what jax_enable_x64 + float64 parameters would do): jaxshim stops
    narrowing to 32 bits and the flax / tfp stand-ins compute in float64.  Used to take clean finite differences
    of the reference's loss closures (the closures' captured data stay the float32 arrays of the fp32 run)."""
    J.X64, flaxshim.F = True, np.float64
    try:
        yield
    finally:
        J.X64, flaxshim.F = False, np.float32


def entropy_eps(seed_value: int, n_agents: int, action_dim: int) -> np.ndarray:
    """The N(0,1) draw `Normal.sample(seed=PRNGKey(seed_value))` makes under the shim: (n_agents, action_dim)."""
    return J._gen(J.PRNGKey(seed_value)).standard_normal((n_agents, action_dim)).astype(F)


def install():
    flaxshim.install()
    jax = sys.modules["jax"]
    jax.value_and_grad, jax.grad, jax.vmap = _value_and_grad, _grad, _vmap_kw
    sys.modules["jax.random"].fold_in = \
        lambda key, data: J.PRNGKey((int(np.asarray(key).ravel()[-1]) * 1000003 + int(np.asarray(data))) % (2 ** 31))
    J._module("flax.training.train_state", TrainState=TrainState)
    J._module("optax", adam=lambda learning_rate, **k: ("adam", learning_rate), apply_if_finite=lambda o, n: o,
              l2_loss=lambda p, t: J._narrow((0.5 * (np.asarray(p, flaxshim.F) - np.asarray(t, flaxshim.F)) ** 2)
                                             .astype(flaxshim.F)),
              piecewise_constant_schedule=_piecewise, constant_schedule=lambda v: (lambda step: v))
    flaxshim.Independent.entropy = _independent_entropy
