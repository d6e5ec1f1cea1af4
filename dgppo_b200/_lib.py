"""ctypes binding of libdgppo_b200.so (include/dgppo_abi.h).

The library is the product path: there is no CPU or PyTorch fallback.  A
missing / unloadable library raises ``DgppoLibraryError`` at first use.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdgppo_b200.so")


class DgppoLibraryError(RuntimeError):
    pass


class DgppoEnvCfg(C.Structure):
    _fields_ = [("kind", C.c_int32), ("n_agents", C.c_int32), ("n_obs", C.c_int32),
                ("n_rays", C.c_int32), ("top_k", C.c_int32), ("reserved_", C.c_int32),
                ("comm_radius", C.c_double), ("car_radius", C.c_double), ("obs_radius", C.c_double),
                ("area_size", C.c_double), ("dt", C.c_double), ("dist2goal", C.c_double),
                ("connect_radius", C.c_double), ("goal_table", C.c_void_p)]


class DgppoGraphDims(C.Structure):
    _fields_ = [(k, C.c_int32) for k in ("state_dim", "node_dim", "edge_dim", "n_obs_nodes",
                                         "n_nodes", "n_edges", "n_ag", "n_ao")]


class DgppoNetCfg(C.Structure):
    _fields_ = [(k, C.c_int32) for k in ("kind", "node_dim", "edge_dim", "n_layers", "n_out")]


class DgppoNetLayout(C.Structure):
    _fields_ = ([(k, C.c_int32 * 2) for k in ("wqk", "wagg", "wu", "bu", "wq", "bq", "wkt", "in_dim", "out_dim")] +
                [(k, C.c_int32) for k in ("d0w", "d0b", "ln0s", "ln0b", "d1w", "d1b", "ln1s", "ln1b",
                                          "wi", "bi", "wh", "bhn", "out_w", "out_b",
                                          "tc_head", "total")])


ABI_VERSION = 7        # DGPPO_ABI_VERSION of include/dgppo_abi.h this mirror was written against

_fp = C.c_void_p   # device pointers travel as integers (tensor.data_ptr())


class DgppoRolloutBuffers(C.Structure):
    _fields_ = [(k, _fp) for k in ("nodes", "edges", "states", "receivers", "senders", "node_type",
                                   "n_node", "n_edge", "rnn", "eps", "actions", "log_pis", "rewards",
                                   "costs", "agent_ws", "hits_ws", "goal", "obstacles", "ray_dirs", "hits_ws2",
                                   "agent_rec", "hits_rec")]


class DgppoStateRecord(C.Structure):
    _fields_ = [(k, _fp) for k in ("agent", "obs_nodes", "goal")]


NET_POLICY, NET_VH, NET_VL = 0, 1, 2
OBS_STRIDE = 16

# name -> (restype, argtypes); mirrors include/dgppo_abi.h one to one
SIGNATURES = {
    "dgppo_abi_version": (C.c_int, []),
    "dgppo_graph_dims": (C.c_int, [C.POINTER(DgppoEnvCfg), C.POINTER(DgppoGraphDims)]),
    "dgppo_n_goals": (C.c_int, [C.POINTER(DgppoEnvCfg)]),
    "dgppo_n_cost": (C.c_int, [C.POINTER(DgppoEnvCfg)]),
    "dgppo_reset": (C.c_int, [_fp, C.POINTER(DgppoEnvCfg), _fp, C.c_double, C.c_double, C.c_double, C.c_double,
                              _fp, _fp, _fp, _fp, C.c_int32]),
    "dgppo_env_step": (C.c_int, [_fp, C.POINTER(DgppoEnvCfg), _fp, _fp, _fp, _fp, _fp, _fp, _fp,
                                 C.c_int32, C.c_int32]),
    "dgppo_lidar": (C.c_int, [_fp, C.POINTER(DgppoEnvCfg), _fp, _fp, _fp, _fp, C.c_int32]),
    "dgppo_build_graph": (C.c_int, [_fp, C.POINTER(DgppoEnvCfg), _fp, _fp, _fp, _fp, _fp, _fp, _fp, _fp,
                                    _fp, _fp, _fp, C.c_int32, C.c_int32]),
    "dgppo_net_layout": (C.c_int, [C.POINTER(DgppoNetCfg), C.POINTER(DgppoNetLayout)]),
    "dgppo_gnn_policy": (C.c_int, [_fp, C.POINTER(DgppoEnvCfg), C.POINTER(DgppoNetCfg), _fp,
                                   _fp, _fp, _fp, _fp, C.c_int32,
                                   _fp, _fp, C.c_int32, _fp, C.c_int32,
                                   _fp, _fp, C.c_int32, C.c_int32]),
    "dgppo_gnn_value": (C.c_int, [_fp, C.POINTER(DgppoEnvCfg), C.POINTER(DgppoNetCfg), _fp,
                                  _fp, _fp, _fp, _fp, C.c_int32,
                                  _fp, _fp, C.c_int32, _fp, C.c_int32, C.c_int32, C.c_int32]),
    "dgppo_vl_scan": (C.c_int, [_fp, C.POINTER(DgppoEnvCfg), C.POINTER(DgppoNetCfg), _fp,
                                _fp, _fp, _fp, _fp, C.c_int32,
                                _fp, C.c_int32, _fp, C.c_int32, C.c_int32, C.c_int32]),
    "dgppo_gnn_policy_from_state": (C.c_int, [_fp, C.POINTER(DgppoEnvCfg), C.POINTER(DgppoNetCfg), _fp,
                                              C.POINTER(DgppoStateRecord), C.c_int32,
                                              _fp, _fp, C.c_int32, _fp, C.c_int32,
                                              _fp, _fp, C.c_int32, C.c_int32]),
    "dgppo_gnn_value_from_state": (C.c_int, [_fp, C.POINTER(DgppoEnvCfg), C.POINTER(DgppoNetCfg), _fp,
                                             C.POINTER(DgppoStateRecord), C.c_int32,
                                             _fp, _fp, C.c_int32, _fp, C.c_int32, C.c_int32, C.c_int32]),
    "dgppo_vl_scan_from_state": (C.c_int, [_fp, C.POINTER(DgppoEnvCfg), C.POINTER(DgppoNetCfg), _fp,
                                           C.POINTER(DgppoStateRecord), C.c_int32,
                                           _fp, C.c_int32, _fp, C.c_int32, C.c_int32, C.c_int32]),
    "dgppo_gae": (C.c_int, [_fp, _fp, _fp, _fp, _fp, C.c_float, C.c_float, _fp, _fp,
                            C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "dgppo_cbf_advantage": (C.c_int, [_fp, _fp, _fp, _fp, C.c_float, C.c_float, C.c_float, C.c_float,
                                      _fp, _fp, _fp, _fp, C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "dgppo_rollout": (C.c_int, [_fp, C.POINTER(DgppoEnvCfg), C.POINTER(DgppoNetCfg), _fp,
                                C.POINTER(DgppoRolloutBuffers), C.c_int32, C.c_int32, _fp]),
    "dgppo_rollout_graph_create": (_fp, [C.POINTER(DgppoEnvCfg), C.POINTER(DgppoNetCfg), _fp,
                                         C.POINTER(DgppoRolloutBuffers), C.c_int32, C.c_int32,
                                         C.POINTER(C.c_int32)]),
    "dgppo_rollout_graph_launch": (C.c_int, [_fp, _fp]),
    "dgppo_rollout_graph_nodes": (C.c_int, [_fp]),
    "dgppo_rollout_graph_destroy": (None, [_fp]),
    "dgppo_prof_create": (_fp, [C.c_int32]),
    "dgppo_prof_destroy": (None, [_fp]),
    "dgppo_prof_read": (C.c_int, [_fp, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
}

_lib: Optional[C.CDLL] = None


def lib() -> C.CDLL:
    """Load (once) and return the kernel library; fail loudly if absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise DgppoLibraryError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or dgppo_b200/csrc/build.sh (there is no CPU fallback)")
        try:
            h = C.CDLL(LIB_PATH)
        except OSError as e:                                   # pragma: no cover
            raise DgppoLibraryError(f"cannot load {LIB_PATH}: {e}") from e
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(h, name)
            fn.restype, fn.argtypes = res, args
        if h.dgppo_abi_version() != ABI_VERSION:
            raise DgppoLibraryError("libdgppo_b200.so ABI version mismatch")
        _lib = h
    return _lib


def check(rc: int, what: str) -> None:
    if rc == 0:
        return
    if rc == -1:
        raise ValueError(f"{what}: invalid arguments (DGPPO_EINVAL)")
    if rc == -2:
        raise NotImplementedError(f"{what}: configuration not supported by the kernels (DGPPO_ENOTSUP)")
    raise RuntimeError(f"{what}: CUDA error {rc}")
