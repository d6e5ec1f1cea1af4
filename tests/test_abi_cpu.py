"""CPU-only checks of the C-ABI library: it loads, exports every symbol that
include/dgppo_abi.h declares, and its host-only entry points (no kernels) agree
with the oracle's size formulas.  No compute calls (there is no GPU here)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from dgppo_b200 import _lib
from dgppo_b200.algo import params as P
from oracle import env_np
from tests.util import CONFIGS, c_cfg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "dgppo_abi.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(dgppo_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    lib = _lib.lib()
    syms = declared_symbols()
    assert len(syms) >= 13, syms
    for s in syms:
        assert hasattr(lib, s), f"libdgppo_b200.so does not export {s}"
    assert set(syms) == set(_lib.SIGNATURES), (set(syms) ^ set(_lib.SIGNATURES))
    assert lib.dgppo_abi_version() == _lib.ABI_VERSION == 7


@pytest.mark.parametrize("name", list(CONFIGS))
def test_graph_dims_match_oracle(name):
    cfg = CONFIGS[name]
    cc, d = c_cfg(cfg), _lib.DgppoGraphDims()
    assert _lib.lib().dgppo_graph_dims(C.byref(cc), C.byref(d)) == 0
    assert (d.state_dim, d.node_dim, d.edge_dim) == (cfg.state_dim, cfg.node_dim, 4)
    assert (d.n_nodes, d.n_edges, d.n_obs_nodes, d.n_ag, d.n_ao) == \
        (cfg.n_nodes, cfg.n_edges, cfg.n_obs_nodes, cfg.n_ag, cfg.n_ao)


def test_survey_graph_sizes():
    """nodes / edges of the BASELINE configs."""
    want = {"C1": (31, 42), "C2": (20, 152), "C3": (81, 192), "C4": (161, 400), "C5": (641, 8704)}
    for k, (N, E) in want.items():
        assert (CONFIGS[k].n_nodes, CONFIGS[k].n_edges) == (N, E)


def test_bad_configs_are_rejected():
    lib = _lib.lib()
    d = _lib.DgppoGraphDims()
    bad = c_cfg(CONFIGS["C3"]); bad.kind = 10
    assert lib.dgppo_graph_dims(C.byref(bad), C.byref(d)) == _lib.lib().dgppo_graph_dims(C.byref(bad), C.byref(d)) == -2
    bad = c_cfg(CONFIGS["C3"]); bad.top_k = 64
    assert lib.dgppo_graph_dims(C.byref(bad), C.byref(d)) == -1
    assert lib.dgppo_graph_dims(None, C.byref(d)) == -1
    L = _lib.DgppoNetLayout()
    assert lib.dgppo_net_layout(C.byref(_lib.DgppoNetCfg(0, 7, 4, 3, 2)), C.byref(L)) == -2   # 3 layers
    assert lib.dgppo_net_layout(C.byref(_lib.DgppoNetCfg(0, 7, 4, 2, 3)), C.byref(L)) == -2   # action_dim 3
    with pytest.raises(NotImplementedError):
        _lib.check(-2, "x")
    with pytest.raises(ValueError):
        _lib.check(-1, "x")
    with pytest.raises(RuntimeError):
        _lib.check(700, "x")


@pytest.mark.parametrize("kind,nd,layers,n_out,count", [(0, 7, 2, 2, 62660), (2, 7, 2, 1, 58305), (1, 7, 1, 2, 39426),
                                                         (0, 8, 2, 2, 62980), (1, 8, 1, 2, 40066)])
def test_param_counts_and_packing(kind, nd, layers, n_out, count):
    """Parameter counts of SURVEY.md A.4 and a round trip through the packed layout."""
    tree = P.init_policy_params(nd, 4, n_out, layers, seed=1, jitter=0.3) if kind == 0 else \
        P.init_value_params(nd, 4, n_out, layers, seed=1, jitter=0.3)
    assert P.count_params(tree) == count
    cfg = P.net_cfg(kind, nd, 4, layers, n_out)
    L = P.net_layout(cfg)
    buf = P.pack_params(tree, cfg)
    assert buf.shape == (L.total,) and buf.dtype == np.float32
    p = tree["params"]
    base = p["PolicyNet_0"] if kind == 0 else p
    g0 = base["GraphTransformerGNN_0"]["GraphTransformer_0"]
    IN, D = L.in_dim[0], L.out_dim[0]
    INP = (IN + 4) // 4 * 4
    np.testing.assert_array_equal(buf[L.wq[0]:L.wq[0] + IN * 3 * D].reshape(IN, 3 * D), g0["Dense_0"]["kernel"])
    wkt = buf[L.wkt[0]:L.wkt[0] + 3 * D * INP].reshape(3, D, INP)
    np.testing.assert_array_equal(wkt[1, 5, :IN], g0["Dense_1"]["kernel"][:, D + 5])
    np.testing.assert_array_equal(wkt[2, :, IN], g0["Dense_1"]["bias"][2 * D:])
    wagg = buf[L.wagg[0]:L.wagg[0] + 3 * (IN + 5) * D].reshape(3, IN + 5, D)
    np.testing.assert_array_equal(wagg[0, :IN], g0["Dense_2"]["kernel"][:, :D])
    np.testing.assert_array_equal(wagg[1, IN], g0["Dense_2"]["bias"][D:2 * D])
    np.testing.assert_array_equal(wagg[2, IN + 1:], g0["Dense_3"]["kernel"][:, 2 * D:])
    (cell,) = base["RNN_0"].values()
    np.testing.assert_array_equal(buf[L.wi:L.wi + 64 * 192].reshape(64, 192)[:, 64:128], cell["iz"]["kernel"])
    np.testing.assert_array_equal(buf[L.bhn:L.bhn + 64], cell["hn"]["bias"])
    for off in (L.wq[0], L.wkt[0], L.wagg[0], L.wu[0], L.d0w, L.wi, L.wh, L.out_w):
        assert off % 4 == 0     # every block 16-byte aligned for float4 / cp.async


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.DgppoLibraryError):
        _lib.lib()


def test_abi_version_is_one_number_everywhere():
    """Header, library, ctypes mirror (and through it __graft_entry__.build()) agree on the ABI version."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "dgppo_abi.h")).read()
    ver = int(re.search(r"#define\s+DGPPO_ABI_VERSION\s+(\d+)", hdr).group(1))
    assert ver == _lib.ABI_VERSION == _lib.lib().dgppo_abi_version()
    entry = open(os.path.join(root, "__graft_entry__.py")).read()
    assert "_lib.ABI_VERSION" in entry and not re.search(r"dgppo_abi_version\(\)\s*==\s*\d", entry)
