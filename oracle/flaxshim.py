"""Minimal NumPy stand-ins for the slices of flax.linen, jraph and
tensorflow_probability that the reference's network code uses, so that
dgppo/nn/{gnn,mlp,rnn}.py and dgppo/algo/module/{policy,value,distribution}.py
can be EXECUTED VERBATIM (under oracle/jaxshim.py) to pin how the reference
COMPOSES its networks: which node feeds query / key / value, reshape and head
order, module construction order (-> parameter names), head / RNN / output
wiring, the thresholded tanh-Normal log-prob.

TEST INFRASTRUCTURE ONLY.  The layer arithmetic inside these stand-ins (Dense,
LayerNorm, GRUCell, segment ops, Normal / Tanh formulas) is a restatement of
the third-party libraries' published behaviour - the same one oracle/nn_np.py
makes - and stays PARITY UNPINNED against a real flax / jraph / tfp install.
"""
from __future__ import annotations

import sys
import types

import numpy as np
from scipy import special as sps

from . import jaxshim as J

F = np.float32
_narrow = J._narrow

# ----------------------------------------------------------------- flax.linen
_SCOPE = []          # stack of (module, params_dict, counters, mode, rng)


class _Ctx:
    def __init__(self, module, params, mode, rng):
        self.module, self.params, self.mode, self.rng = module, params, mode, rng
        self.counters = {}


def compact(fn):
    fn._compact = True
    return fn


class Module:
    """Dataclass-like module: annotated class attributes are constructor fields."""
    name = None

    def __init_subclass__(cls, **kw):
        super().__init_subclass__(**kw)
        fields = []
        for klass in reversed(cls.__mro__):
            for k in getattr(klass, "__annotations__", {}):
                if k not in ("name", "parent") and k not in fields:
                    fields.append(k)
        cls._fields_ = fields

    def __init__(self, *args, **kwargs):
        fields = type(self)._fields_
        for k, v in zip(fields, args):
            setattr(self, k, v)
        name = kwargs.pop("name", None)
        for k, v in kwargs.items():
            setattr(self, k, v)
        for k in fields:
            if not hasattr(self, k):
                raise TypeError(f"{type(self).__name__}: missing field {k}")
        # flax assigns the auto-name at construction inside a parent's compact method
        if _SCOPE:
            ctx = _SCOPE[-1]
            if name is None:
                i = ctx.counters.get(type(self).__name__, 0)
                ctx.counters[type(self).__name__] = i + 1
                name = f"{type(self).__name__}_{i}"
        self.name = name

    def param(self, name, init_fn, *shape_args):
        ctx = _SCOPE[-1]
        if ctx.mode == "init" and name not in ctx.params:
            ctx.params[name] = np.asarray(init_fn(ctx.rng, *shape_args), F)
        return _narrow(np.asarray(ctx.params[name], F))

    def _run(self, params, mode, rng, args, kwargs):
        _SCOPE.append(_Ctx(self, params, mode, rng))
        try:
            return type(self).__call__.__wrapped_call__(self, *args, **kwargs) \
                if hasattr(type(self).__call__, "__wrapped_call__") else self._call_impl(*args, **kwargs)
        finally:
            _SCOPE.pop()

    def _call_impl(self, *args, **kwargs):
        raise NotImplementedError

    def __call__(self, *args, **kwargs):       # overridden by subclasses' own __call__
        raise NotImplementedError

    def init(self, key, *args, **kwargs):
        params = {}
        rng = np.random.default_rng(int(np.asarray(key).ravel()[-1]))
        _invoke(self, params, "init", rng, args, kwargs)
        return {"params": params}

    def apply(self, variables, *args, **kwargs):
        return _invoke(self, variables["params"], "apply", None, args, kwargs)


def _invoke(module, params, mode, rng, args, kwargs):
    _SCOPE.append(_Ctx(module, params, mode, rng))
    try:
        return type(module).__dict__["__call__"](module, *args, **kwargs) \
            if "__call__" in type(module).__dict__ else _find_call(module)(module, *args, **kwargs)
    finally:
        _SCOPE.pop()


def _find_call(module):
    for klass in type(module).__mro__:
        if "__call__" in klass.__dict__ and klass is not Module:
            return klass.__dict__["__call__"]
    raise TypeError("module has no __call__")


def _child_call(module, args, kwargs):
    """A module called from inside its parent's compact method."""
    parent = _SCOPE[-1]
    sub = parent.params.setdefault(module.name, {}) if parent.mode == "init" else parent.params.get(module.name, {})
    out = _invoke(module, sub, parent.mode, parent.rng, args, kwargs)
    if parent.mode == "init" and not sub:
        parent.params.pop(module.name, None)        # modules without parameters leave no entry
    return out


class _ModuleMeta(type):
    pass


def _wrap_user_call(cls):
    """Route `module(...)` through the scope machinery."""
    user_call = cls.__dict__.get("__call__")
    if user_call is None or getattr(user_call, "_routed", False):
        return

    def routed(self, *args, **kwargs):
        if _SCOPE and _SCOPE[-1].module is not self:
            return _child_call(self, args, kwargs)
        return user_call(self, *args, **kwargs)
    routed._routed = True
    routed._inner = user_call
    cls.__call__ = routed


_orig_init_subclass = Module.__init_subclass__.__func__


def _init_subclass(cls, **kw):
    _orig_init_subclass(cls, **kw)
    _wrap_user_call(cls)


Module.__init_subclass__ = classmethod(_init_subclass)


def _invoke(module, params, mode, rng, args, kwargs):      # noqa: F811  (final definition)
    call = None
    for klass in type(module).__mro__:
        c = klass.__dict__.get("__call__")
        if c is not None and klass is not Module:
            call = getattr(c, "_inner", c)
            break
    _SCOPE.append(_Ctx(module, params, mode, rng))
    try:
        return call(module, *args, **kwargs)
    finally:
        _SCOPE.pop()


def orthogonal(scale=1.0, column_axis=-1):
    def init(rng, shape, dtype=F):
        n_rows, n_cols = int(np.prod(shape[:-1])), shape[-1]
        a = rng.standard_normal((max(n_rows, n_cols), min(n_rows, n_cols)))
        q, r = np.linalg.qr(a)
        q = q * np.sign(np.diag(r))
        if n_rows < n_cols:
            q = q.T
        return (scale * q.reshape(shape)).astype(F)
    return init


def lecun_normal():
    return lambda rng, shape, dtype=F: (rng.standard_normal(shape) / np.sqrt(shape[0])).astype(F)


def zeros_init(rng, shape, dtype=F):
    return np.zeros(shape, F)


def ones_init(rng, shape, dtype=F):
    return np.ones(shape, F)


class Dense(Module):
    features: int
    use_bias: bool = True
    kernel_init: object = None

    def __call__(self, x):
        x = np.asarray(x, F)
        kinit = self.kernel_init if self.kernel_init is not None else lecun_normal()
        k = self.param("kernel", kinit, (x.shape[-1], self.features))
        y = x @ np.asarray(k)
        if self.use_bias:
            # init draws a small random bias so every term is exercised (flax's default is zeros)
            y = y + np.asarray(self.param("bias", lambda rng, s: 0.1 * rng.standard_normal(s), (self.features,)))
        return _narrow(y.astype(F))


class LayerNorm(Module):
    epsilon: float = 1e-6

    def __call__(self, x):
        x = np.asarray(x, F)
        scale = np.asarray(self.param("scale", lambda rng, s: 1.0 + 0.1 * rng.standard_normal(s), (x.shape[-1],)))
        bias = np.asarray(self.param("bias", lambda rng, s: 0.1 * rng.standard_normal(s), (x.shape[-1],)))
        mean = x.mean(-1, keepdims=True, dtype=F)
        mean2 = (x * x).mean(-1, keepdims=True, dtype=F)
        var = np.maximum(F(0), mean2 - mean * mean)
        mul = (F(1) / np.sqrt(var + F(self.epsilon))) * scale
        return _narrow(((x - mean) * mul + bias).astype(F))


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


class GRUCell(Module):
    features: int

    def __call__(self, carry, inputs):
        h = np.asarray(carry, F)
        di = lambda n: Dense(self.features, use_bias=True, kernel_init=lecun_normal(), name=n)     # noqa: E731
        dh = lambda n, b=False: Dense(self.features, use_bias=b, kernel_init=orthogonal(), name=n)  # noqa: E731
        r = _sigmoid(np.asarray(di("ir")(inputs)) + np.asarray(dh("hr")(h))).astype(F)
        z = _sigmoid(np.asarray(di("iz")(inputs)) + np.asarray(dh("hz")(h))).astype(F)
        n = np.tanh(np.asarray(di("in")(inputs)) + r * np.asarray(dh("hn", True)(h))).astype(F)
        new_h = _narrow(((F(1) - z) * n + z * h).astype(F))
        return new_h, new_h

    def initialize_carry(self, rng, input_shape):
        return _narrow(np.zeros(tuple(input_shape[:-1]) + (self.features,), F))


class LSTMCell(Module):
    features: int

    def __call__(self, carry, inputs):
        raise NotImplementedError("LSTM is outside the hot path")


class Dropout(Module):
    rate: float = 0.0
    deterministic: bool = True

    def __call__(self, x):
        return x


def relu(x):
    return _narrow(np.maximum(np.asarray(x), 0))


# ---------------------------------------------------------------------- jraph
def segment_sum(data, segment_ids, num_segments=None, **kw):
    data, seg = np.asarray(data), np.asarray(segment_ids)
    out = np.zeros((num_segments,) + data.shape[1:], data.dtype)
    np.add.at(out, seg, data)
    return _narrow(out)


def segment_softmax(logits, segment_ids, num_segments=None, **kw):
    logits, seg = np.asarray(logits), np.asarray(segment_ids)
    mx = np.full((num_segments,) + logits.shape[1:], -np.inf, logits.dtype)
    np.maximum.at(mx, seg, logits)
    ex = np.exp(logits - mx[seg]).astype(logits.dtype)
    sm = np.zeros_like(mx)
    np.add.at(sm, seg, ex)
    return _narrow((ex / sm[seg]).astype(logits.dtype))


# ------------------------------------------------------------------------ tfp
def _ndtr(x):
    hs2 = F(0.5 * np.sqrt(2.0))
    w = x * hs2
    z = np.abs(w)
    y = np.where(z < hs2, 1 + sps.erf(w), np.where(w > 0, 2 - sps.erfc(z), sps.erfc(z)))
    return (F(0.5) * y).astype(F)


def _log_ndtr(x):
    x = np.asarray(x, F)
    with np.errstate(all="ignore"):
        xl = np.minimum(x, F(-10))
        x2 = xl * xl
        series = 1.0 - 1.0 / x2 + 3.0 / (x2 * x2) - 15.0 / (x2 * x2 * x2)
        low = (-0.5 * x2 - np.log(-xl) - 0.5 * np.log(2.0 * np.pi) + np.log(series)).astype(F)
        mid = np.log(_ndtr(np.maximum(x, F(-10)))).astype(F)
        up = (-_ndtr(-x)).astype(F)
    return np.where(x > 5, up, np.where(x > -10, mid, low)).astype(F)


class Distribution:
    def mode(self):
        return self._mode()

    def sample(self, seed=None, sample_shape=()):
        raise NotImplementedError


class Normal(Distribution):
    def __init__(self, loc, scale):
        self.loc, self.scale = np.asarray(loc, F), np.asarray(scale, F)

    def sample(self, seed=None, sample_shape=()):
        eps = J._gen(seed).standard_normal(self.loc.shape).astype(F)
        self.last_eps = eps
        return _narrow((self.loc + self.scale * eps).astype(F))

    def log_prob(self, x):
        x = np.asarray(x, F)
        d = x / self.scale - self.loc / self.scale
        return _narrow((F(-0.5) * d * d - (F(0.5 * np.log(2.0 * np.pi)) + np.log(self.scale))).astype(F))

    def log_cdf(self, x):
        return _narrow(_log_ndtr((np.asarray(x, F) - self.loc) / self.scale))

    def log_survival_function(self, x):
        return _narrow(_log_ndtr(-((np.asarray(x, F) - self.loc) / self.scale)))

    def _mode(self):
        return _narrow(self.loc)

    def entropy(self):
        return _narrow((F(0.5 + 0.5 * np.log(2 * np.pi)) + np.log(self.scale)).astype(F))


class Tanh:
    def forward(self, x):
        return _narrow(np.tanh(np.asarray(x, F)))

    def inverse(self, y):
        return _narrow(np.arctanh(np.asarray(y, F)).astype(F))

    def forward_log_det_jacobian(self, x, event_ndims=0):
        x = np.asarray(x, F)
        return _narrow((F(2.0) * (F(np.log(2.0)) - x - np.logaddexp(F(-2.0) * x, 0))).astype(F))


class TransformedDistribution(Distribution):
    def __init__(self, distribution, bijector, validate_args=False, **kw):
        self.distribution, self.bijector = distribution, bijector

    def sample(self, seed=None, sample_shape=()):
        return self.bijector.forward(self.distribution.sample(seed=seed))

    def log_prob(self, value, **kw):
        x = self.bijector.inverse(value)
        return _narrow((np.asarray(self.distribution.log_prob(x)) -
                        np.asarray(self.bijector.forward_log_det_jacobian(x))).astype(F))

    def _mode(self):
        return self.bijector.forward(self.distribution.mode())

    @classmethod
    def _parameter_properties(cls, dtype, num_classes=None):
        return {"bijector": None}


class Independent(Distribution):
    def __init__(self, distribution, reinterpreted_batch_ndims=1):
        self.distribution, self.nd = distribution, reinterpreted_batch_ndims

    def sample(self, seed=None, sample_shape=()):
        return self.distribution.sample(seed=seed)

    def log_prob(self, value):
        lp = np.asarray(self.distribution.log_prob(value))
        return _narrow(lp.sum(axis=tuple(range(-self.nd, 0)), dtype=F))

    def _mode(self):
        return self.distribution.mode()


def install():
    """jaxshim.install() + working flax.linen / jraph / tfp stand-ins."""
    J.install()
    nn = J._module("flax.linen", Module=Module, compact=compact, Dense=Dense, LayerNorm=LayerNorm, GRUCell=GRUCell,
                   LSTMCell=LSTMCell, Dropout=Dropout, relu=relu,
                   initializers=types.SimpleNamespace(orthogonal=orthogonal, lecun_normal=lecun_normal,
                                                      zeros=zeros_init, ones=ones_init, Initializer=object),
                   tanh=lambda x: _narrow(np.tanh(x)), elu=None, swish=None, silu=None, gelu=None, softplus=None)
    flax = sys.modules["flax"]
    flax.linen = nn
    J._module("flax.core", FrozenDict=dict)
    J._module("jraph", segment_sum=segment_sum, segment_softmax=segment_softmax)
    tfd = types.SimpleNamespace(Normal=Normal, Independent=Independent, TransformedDistribution=TransformedDistribution,
                                Distribution=Distribution)
    tfb = types.SimpleNamespace(Tanh=Tanh)
    sub = J._module("tensorflow_probability.substrates.jax", distributions=tfd, bijectors=tfb)
    sys.modules["tensorflow_probability.substrates"].jax = sub
    sys.modules["tensorflow_probability"].substrates = sys.modules["tensorflow_probability.substrates"]
