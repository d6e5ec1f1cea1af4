"""A NumPy stand-in for the slice of the jax API that the reference's env /
graph / LiDAR / GAE code uses, so that code can be EXECUTED VERBATIM from
/root/reference in this container (jax is not installable here) to pin the
oracle and to generate the golden fixtures under tests/golden/.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  What it models:
  * jax.numpy: NumPy functions whose results are narrowed to jax's default
    x64-disabled dtypes (float64 -> float32, int64 -> int32); arrays are an
    ndarray subclass that adds the functional `.at[idx].set/add/get` updates;
  * jax.vmap: a Python loop over the mapped axis with pytree stacking;
  * jax.lax.scan / while_loop: Python loops;
  * jax.tree_util: flatten/map over tuples, lists, dicts, NamedTuples, None and
    classes registered with register_pytree_with_keys_class;
  * jax.random: NumPy Generator draws (NOT threefry: streams differ from real
    jax, which only matters for which initial states a key produces).
Modules the reference imports but never executes on this path (matplotlib,
flax, optax, jraph, tensorflow_probability, wandb, ipdb, ...) are replaced by
inert stubs.  Arithmetic is NumPy's: fp32, individually rounded, no FMA - the
same convention as oracle/env_np.py.
"""
from __future__ import annotations

import importlib.machinery
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("DGPPO_REFERENCE_ROOT", "/root/reference")


# ------------------------------------------------------------------ arrays
class _At:
    def __init__(self, arr):
        self.arr = arr

    def __getitem__(self, idx):
        return _AtIdx(self.arr, idx)


class _AtIdx:
    def __init__(self, arr, idx):
        self.arr, self.idx = arr, idx

    def set(self, v):
        out = np.array(self.arr, copy=True).view(ShimArray)
        out[self.idx] = v
        return out

    def add(self, v):
        out = np.array(self.arr, copy=True).view(ShimArray)
        np.add.at(out, self.idx, v)
        return out

    def get(self, mode=None, fill_value=None):
        return np.asarray(self.arr)[self.idx].view(ShimArray)


class ShimArray(np.ndarray):
    @property
    def at(self):
        return _At(self)


X64 = False        # True: keep float64 / int64 results (jax_enable_x64); see oracle/algoshim.x64()


def _narrow(x):
    if X64:
        return x.view(ShimArray) if isinstance(x, np.ndarray) else x
    if isinstance(x, np.ndarray):
        if x.dtype == np.float64:
            x = x.astype(np.float32)
        elif x.dtype == np.int64:
            x = x.astype(np.int32)
        return x.view(ShimArray)
    if isinstance(x, np.float64):
        return np.float32(x)
    if isinstance(x, np.int64):
        return np.int32(x)
    if isinstance(x, tuple):
        return tuple(_narrow(v) for v in x)
    return x


def _wrap(fn):
    def inner(*a, **k):
        with np.errstate(all="ignore"):
            return _narrow(fn(*a, **k))
    inner.__name__ = getattr(fn, "__name__", "fn")
    return inner


def _argsort(a, axis=-1, **kw):
    return np.argsort(a, axis=axis, kind="stable")      # jnp.argsort is stable


def _array(x, dtype=None, **kw):
    return np.array(x, dtype=dtype)


def _clip(a, a_min=None, a_max=None, min=None, max=None):
    lo = a_min if a_min is not None else min
    hi = a_max if a_max is not None else max
    return np.clip(a, lo, hi)


# ------------------------------------------------------------------ pytrees
_REGISTERED = []


def _is_namedtuple(x):
    return isinstance(x, tuple) and hasattr(x, "_fields")


def tree_flatten(t):
    leaves = []

    def rec(x):
        if x is None:
            return ("none",)
        for cls in _REGISTERED:
            if isinstance(x, cls):
                kids, aux = x.tree_flatten_with_keys()
                return ("reg", cls, aux, [rec(v) for _, v in kids])
        if _is_namedtuple(x):
            return ("nt", type(x), [rec(v) for v in x])
        if isinstance(x, (tuple, list)):
            return ("seq", type(x), [rec(v) for v in x])
        if isinstance(x, dict):
            return ("dict", list(x.keys()), [rec(v) for v in x.values()])
        leaves.append(x)
        return ("leaf",)
    return leaves, rec(t)


def tree_unflatten(treedef, leaves):
    it = iter(leaves)

    def rec(d):
        k = d[0]
        if k == "none":
            return None
        if k == "leaf":
            return next(it)
        if k == "reg":
            return d[1].tree_unflatten(d[2], [rec(c) for c in d[3]])
        if k == "nt":
            return d[1](*[rec(c) for c in d[2]])
        if k == "seq":
            return d[1](rec(c) for c in d[2])
        if k == "dict":
            return {key: rec(c) for key, c in zip(d[1], d[2])}
        raise TypeError(k)
    return rec(treedef)


def tree_map(f, tree, *rest):
    leaves, td = tree_flatten(tree)
    others = [tree_flatten(r)[0] for r in rest]
    return tree_unflatten(td, [f(*xs) for xs in zip(leaves, *others)])


def register_pytree_with_keys_class(cls):
    _REGISTERED.append(cls)
    return cls


class GetAttrKey:
    def __init__(self, name):
        self.name = name


# ---------------------------------------------------------------- transforms
def vmap(fn, in_axes=0, out_axes=0):
    def mapped(*args, **kwargs):
        if not isinstance(in_axes, (tuple, list)):
            axes = [in_axes] * len(args)
        else:
            axes = list(in_axes)
        n = None
        for a, ax in zip(args, axes):
            if ax is None:
                continue
            lv = tree_flatten(a)[0]
            if lv:
                n = np.shape(lv[0])[ax]
                break
        outs = []
        for i in range(n):
            sl = [a if ax is None else tree_map(lambda x: np.take(x, i, axis=ax).view(ShimArray)
                                                if isinstance(x, np.ndarray) else x, a)
                  for a, ax in zip(args, axes)]
            outs.append(fn(*sl, **kwargs))
        l0, td = tree_flatten(outs[0])
        stacked = [_narrow(np.stack([np.asarray(tree_flatten(o)[0][j]) for o in outs], axis=0))
                   for j in range(len(l0))]
        return tree_unflatten(td, stacked)
    return mapped


def jit(fn=None, **kw):
    if fn is None:
        return lambda f: f
    return fn


def scan(f, init, xs=None, length=None, reverse=False, unroll=1):
    if xs is None:
        n = length
    else:
        n = np.shape(tree_flatten(xs)[0][0])[0]
    order = range(n - 1, -1, -1) if reverse else range(n)
    carry, ys = init, [None] * n
    for i in order:
        x = None if xs is None else tree_map(lambda a: _narrow(np.asarray(a)[i]), xs)
        carry, y = f(carry, x)
        ys[i] = y
    l0, td = tree_flatten(ys[0])
    if not l0:
        return carry, ys[0]
    stacked = [_narrow(np.stack([np.asarray(tree_flatten(y)[0][j]) for y in ys], axis=0)) for j in range(len(l0))]
    return carry, tree_unflatten(td, stacked)


def while_loop(cond_fun, body_fun, init_val):
    v = init_val
    while bool(cond_fun(v)):
        v = body_fun(v)
    return v


# ------------------------------------------------------------------- random
class _Key(np.ndarray):
    pass


def PRNGKey(seed):
    return np.array([0, int(seed) & 0xFFFFFFFF], dtype=np.uint32)


def _gen(key):
    k = np.asarray(key, dtype=np.uint64).ravel()
    return np.random.default_rng([int(v) for v in k])


def split(key, num=2):
    g = _gen(key)
    return g.integers(0, 2 ** 32, size=(num, 2), dtype=np.uint64).astype(np.uint32)


def uniform(key, shape=(), dtype=np.float32, minval=0.0, maxval=1.0):
    g = _gen(key)
    u = g.random(size=shape, dtype=np.float32)
    lo, hi = np.asarray(minval, np.float32), np.asarray(maxval, np.float32)
    return _narrow((u * (hi - lo) + lo).astype(np.float32))


def randint(key, shape, minval, maxval, dtype=np.int32):
    """jax.random.randint: integers in [minval, maxval) (mpe_line.py:65, lidar_line.py:61)."""
    return _narrow(_gen(key).integers(int(minval), int(maxval), size=shape).astype(np.int32))


def normal(key, shape=(), dtype=np.float32):
    return _narrow(_gen(key).standard_normal(size=shape).astype(np.float32))


# --------------------------------------------------------------- module glue
class _StubMeta(type):
    """Classes whose unknown class attributes are again stub classes."""

    def __getattr__(cls, item):
        if item.startswith("__") and item.endswith("__"):
            raise AttributeError(item)
        return _stub_class(item)

    def __call__(cls, *a, **k):
        if len(a) == 1 and callable(a[0]) and not isinstance(a[0], type) and not k and cls.__dict__.get("_pure_stub"):
            return a[0]                       # decorator use: @nn.compact, @jax.jit(...)
        return super().__call__(*a, **k)


def _stub_class(name):
    return _StubMeta(name, (), {"_pure_stub": True,
                                "__init__": lambda self, *a, **k: None,
                                "__call__": lambda self, *a, **k: (a[0] if len(a) == 1 and callable(a[0]) else None),
                                "__getattr__": lambda self, item: _stub_class(item),
                                "__class_getitem__": classmethod(lambda cls, k: cls),
                                "__iter__": lambda self: iter(()),
                                "__or__": lambda self, o: self, "__ror__": lambda self, o: self})


class _Stub:
    """Inert attribute sink used for whole stub modules."""

    def __init__(self, name="stub"):
        self._name = name

    def __getattr__(self, item):
        if item.startswith("__") and item.endswith("__"):
            raise AttributeError(item)
        full = f"{self._name}.{item}"
        if full in sys.modules:
            return sys.modules[full]
        return _stub_class(item)

    def __call__(self, *a, **k):
        if len(a) == 1 and callable(a[0]) and not k:
            return a[0]
        return _Stub(self._name)


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, None)
    m.__path__ = []
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _stub_module(name):
    m = _module(name)
    stub = _Stub(name)
    m.__getattr__ = lambda item: getattr(stub, item)
    return m


def _build_jnp():
    ns = {}
    for name in ("zeros", "ones", "full", "eye", "arange", "linspace", "concatenate", "stack", "cos", "sin",
                 "arctan2", "sqrt", "abs", "sign", "minimum", "maximum", "where", "logical_and", "logical_or",
                 "logical_not", "less", "expand_dims", "reshape", "zeros_like", "ones_like", "any", "all", "cumsum",
                 "dot", "tile", "repeat", "roll", "sum", "mean", "power", "exp", "log", "tanh", "square", "floor",
                 "isnan", "isinf", "isfinite", "array_split", "split", "take", "squeeze", "transpose", "matmul",
                 "einsum", "argmin", "argmax", "prod", "broadcast_to", "atleast_1d", "cross", "flip", "meshgrid",
                 "std", "var", "allclose", "outer", "diag", "triu", "tril", "mod", "floor_divide", "greater",
                 "greater_equal", "less_equal", "equal", "not_equal", "nan_to_num", "asarray", "max", "min",
                 "full_like", "arccos", "arcsin", "tan", "hstack", "vstack", "round", "ceil", "log1p", "expm1",
                 "argwhere", "nonzero", "unique", "sort", "swapaxes", "moveaxis", "identity", "trace", "copy"):
        if hasattr(np, name):
            ns[name] = _wrap(getattr(np, name))
    ns["array"] = _wrap(_array)
    ns["clip"] = _wrap(_clip)
    ns["argsort"] = _wrap(_argsort)
    ns["float32"], ns["int32"], ns["bool_"], ns["uint32"] = np.float32, np.int32, np.bool_, np.uint32
    ns["float64"], ns["int64"] = np.float64, np.int64
    ns["ndarray"], ns["pi"], ns["inf"], ns["nan"], ns["newaxis"] = np.ndarray, np.pi, np.inf, np.nan, None
    lin = types.SimpleNamespace(norm=_wrap(np.linalg.norm), inv=_wrap(np.linalg.inv), det=_wrap(np.linalg.det))
    ns["linalg"] = lin
    return ns


def install(reference_root: str = REFERENCE_ROOT):
    """Install the fake `jax` (+ inert stubs) and make `dgppo.*` importable from
    the read-only reference checkout without executing its package __init__s
    (which would import every env, VMAS and the algos)."""
    if "jax" in sys.modules and getattr(sys.modules["jax"], "__shim__", False):
        return
    jnp = _module("jax.numpy", **_build_jnp())
    jr = _module("jax.random", PRNGKey=PRNGKey, split=split, uniform=uniform, normal=normal, randint=randint, key=PRNGKey)
    lax = _module("jax.lax", scan=scan, while_loop=while_loop,
                  cond=lambda p, t, f, *a: t(*a) if p else f(*a), stop_gradient=lambda x: x)
    jtu = _module("jax.tree_util", tree_map=tree_map, tree_flatten=tree_flatten, tree_unflatten=tree_unflatten,
                  register_pytree_with_keys_class=register_pytree_with_keys_class, GetAttrKey=GetAttrKey,
                  tree_leaves=lambda t: tree_flatten(t)[0], PyTreeDef=object,
                  tree_structure=lambda t: tree_flatten(t)[1])
    jnn = _module("jax.nn", softplus=_wrap(lambda x: np.logaddexp(x, 0)), relu=_wrap(lambda x: np.maximum(x, 0)))
    typing_m = _module("jax.typing", ArrayLike=np.ndarray, DTypeLike=object)
    src_tu = _module("jax._src.tree_util", GetAttrKey=GetAttrKey)
    src_lib = _module("jax._src.lib", xla_client=_Stub("xla_client"))
    _module("jax._src.lib.xla_client")
    src = _module("jax._src", tree_util=src_tu, lib=src_lib)
    _stub_module("jax.scipy")
    _stub_module("jax.scipy.spatial")
    _stub_module("jax.scipy.spatial.transform")
    _stub_module("jax.debug")
    _module("jax", numpy=jnp, random=jr, lax=lax, tree_util=jtu, nn=jnn, typing=typing_m, _src=src,
            vmap=vmap, jit=jit, Array=np.ndarray, tree_map=tree_map, __version__="0.0.0-shim", __shim__=True,
            config=types.SimpleNamespace(update=lambda *a, **k: None))
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.animation", "matplotlib.collections",
                 "matplotlib.patches", "matplotlib.colors", "matplotlib.cm", "matplotlib.lines",
                 "mpl_toolkits", "mpl_toolkits.mplot3d", "mpl_toolkits.mplot3d.art3d",
                 "flax", "flax.core", "flax.linen", "flax.training", "flax.training.train_state",
                 "optax", "jraph", "tensorflow_probability", "tensorflow_probability.substrates",
                 "tensorflow_probability.substrates.jax", "wandb", "ipdb", "equinox", "colour", "seaborn", "cv2",
                 "PIL", "imageio"):
        if name not in sys.modules:
            _stub_module(name)
    for pkg in ("dgppo", "dgppo.env", "dgppo.env.lidar_env", "dgppo.env.mpe", "dgppo.algo", "dgppo.utils",
                "dgppo.trainer", "dgppo.nn", "dgppo.algo.module"):
        m = types.ModuleType(pkg)
        m.__path__ = [reference_root + "/" + pkg.replace(".", "/")]
        m.__spec__ = importlib.machinery.ModuleSpec(pkg, None, is_package=True)
        m.__spec__.submodule_search_locations = m.__path__
        sys.modules[pkg] = m


def to_np(x):
    return np.asarray(x)
