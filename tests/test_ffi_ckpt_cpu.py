"""Boundary artefacts that cannot run here but must not rot: (1) the XLA-FFI shim source type-checks against a stub
of xla/ffi/api/ffi.h (every handler's call into the C ABI has the header's arity and types); (2) a checkpoint
shaped like the reference's (`actor.pkl` / `Vl.pkl` / `Vh.pkl`: flax FrozenDict-style mappings, array-like leaves,
`GRUCell_<k>` auto-name index ambiguity, informarl_lagr.py:311-327) converts and packs."""
import os
import pickle
import re
import subprocess
from collections.abc import Mapping

import numpy as np

from dgppo_b200 import _lib
from dgppo_b200.algo import params as P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_ffi_shim_type_checks_against_stub_header():
    src = os.path.join(ROOT, "dgppo_b200", "csrc", "dgppo_ffi.cc")
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-I", os.path.join(ROOT, "tests", "ffi_stub"),
                        "-I", "/usr/local/cuda/include", "-I", os.path.join(ROOT, "include"), src],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    text = open(src).read()
    handlers = set(re.findall(r"XLA_FFI_DEFINE_HANDLER_SYMBOL\((\w+),", text))
    assert handlers == {"DgppoReset", "DgppoEnvStep", "DgppoLidar", "DgppoBuildGraph", "DgppoGnnPolicy",
                        "DgppoGnnValue", "DgppoVlScan", "DgppoGae", "DgppoCbfAdvantage", "DgppoRollout"}
    # every compute entry point of the header is reached by a handler
    called = set(re.findall(r"\b(dgppo_[a-z_]+)\(st", text))
    assert called == {"dgppo_reset", "dgppo_env_step", "dgppo_lidar", "dgppo_build_graph", "dgppo_gnn_policy",
                      "dgppo_gnn_value", "dgppo_vl_scan", "dgppo_gae", "dgppo_cbf_advantage", "dgppo_rollout"}
    # without the jaxlib headers the file is an empty translation unit (what build_ffi.sh relies on)
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), src],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


class FrozenLike(Mapping):
    """Stands in for flax.core.FrozenDict: a Mapping that is NOT a dict."""

    def __init__(self, d):
        self._d = {k: FrozenLike(v) if isinstance(v, dict) else v for k, v in d.items()}

    def __getitem__(self, k):
        return self._d[k]

    def __iter__(self):
        return iter(self._d)

    def __len__(self):
        return len(self._d)


class ArrayLike:
    """Stands in for a jax Array leaf: only __array__ is available."""

    def __init__(self, a):
        self._a = np.asarray(a, np.float64)         # a checkpoint may even carry another dtype

    def __array__(self, dtype=None, copy=None):
        return self._a if dtype is None else self._a.astype(dtype)


def _rename_gru(tree, name):
    out = {}
    for k, v in tree.items():
        if isinstance(v, dict):
            v = _rename_gru(v, name)
        out[name if k.startswith("GRUCell_") else k] = v
    return out


def _wrap(tree):
    return FrozenLike({k: (_wrap(v)._d if isinstance(v, dict) else ArrayLike(v)) for k, v in tree.items()})


def test_reference_shaped_checkpoint_converts_and_packs(tmp_path):
    from dgppo_b200.algo.dgppo import _to_numpy_tree
    pol = P.init_policy_params(7, 4, 2, 2, seed=3, jitter=0.2)
    cfg = P.net_cfg(_lib.NET_POLICY, 7, 4, 2, 2)
    want = P.pack_params(pol, cfg)
    # flax names the GRU cell by how many cells RNN.__call__ instantiated before it (nn/rnn.py:19-22):
    # GRUCell_0 in some module layouts, GRUCell_1 in the reference's fixtures
    for gru_name in ("GRUCell_0", "GRUCell_1", "GRUCell_3"):
        ckpt = _wrap(_rename_gru(pol, gru_name))
        assert not isinstance(ckpt, dict)
        tree = _to_numpy_tree(ckpt)
        assert isinstance(tree, dict) and tree["params"]["ScaleHid"]["kernel"].dtype == np.float32
        assert list(tree["params"]["PolicyNet_0"]["RNN_0"]) == [gru_name]
        np.testing.assert_array_equal(P.pack_params(tree, cfg), want)
    # the pickle round trip of save / load uses plain nested dicts of NumPy arrays
    path = tmp_path / "actor.pkl"
    with open(path, "wb") as f:
        pickle.dump(_rename_gru(pol, "GRUCell_1"), f)
    with open(path, "rb") as f:
        np.testing.assert_array_equal(P.pack_params(_to_numpy_tree(pickle.load(f)), cfg), want)
    # value nets: same rule, n_out 1 / 2 / 3
    for n_out, layers, kind in ((1, 2, _lib.NET_VL), (2, 1, _lib.NET_VH), (3, 1, _lib.NET_VH)):
        v = P.init_value_params(7, 4, n_out, layers, seed=4, jitter=0.2)
        c = P.net_cfg(kind, 7, 4, layers, n_out)
        np.testing.assert_array_equal(P.pack_params(_to_numpy_tree(_wrap(_rename_gru(v, "GRUCell_2"))), c),
                                      P.pack_params(v, c))
