"""GPU: the graph-from-state forwards (dgppo_gnn_policy_from_state / _value_from_state / _vl_scan_from_state) against
the graph-record forwards on the graphs K3 builds from the same state: the staging phase uses build_graph's own
arithmetic, so every output must be IDENTICAL (bit for bit), for every env family and both GNN kernels."""
import ctypes as C

import numpy as np
import pytest
import torch

from dgppo_b200 import _lib
from dgppo_b200.algo import params as P
from oracle import env_np
from tests import util
from tests.util import CONFIGS

pytestmark = pytest.mark.gpu
F = np.float32

CASES = dict(CONFIGS)
CASES.update({
    "lidar_line": env_np.EnvCfg(env_np.LIDAR_LINE, n=4, n_obs=3),
    "mpe_line": env_np.EnvCfg(env_np.MPE_LINE, n=5, n_obs=3),
    "mpe_formation": env_np.EnvCfg(env_np.MPE_FORMATION, n=4, n_obs=3),
    "mpe_connect": env_np.EnvCfg(env_np.MPE_CONNECT_SPREAD, n=3, n_obs=1, area=1.0, obs_radius=0.25),
})


def _states(cfg, b, seed):
    """States with many pairs near the mask thresholds (agents clustered), hits from the LiDAR kernel."""
    rng = np.random.default_rng(seed)
    agent, goal, obstacles, mpe_obs = env_np.synthetic_states(cfg, b, seed)
    agent[..., :2] = (0.6 + rng.uniform(0, 0.7, (b, cfg.n, 2))).astype(F)
    goal = goal[:, :cfg.n_goal]
    if cfg.is_lidar and cfg.n_obs > 0:
        rays = env_np.ray_table(cfg.n_rays, cfg.comm_radius)
        obs_nodes = util.k_lidar(cfg, agent, obstacles, rays)
    else:
        obs_nodes = mpe_obs
    return agent, goal, obs_nodes


@pytest.mark.parametrize("name", list(CASES))
def test_from_state_equals_graph_record(name):
    cfg = CASES[name]
    b = 6 if cfg.n >= 40 else 37
    agent, goal, obs_nodes = _states(cfg, b, 11)
    graph = util.k_graph(cfg, agent, goal, obs_nodes)
    n = cfg.n
    rng = np.random.default_rng(5)
    rnn = (rng.standard_normal((b, n, 64)) * 0.3).astype(F)
    eps = rng.standard_normal((b, n, 2)).astype(F)
    cc = util.c_cfg(cfg)
    a_d, g_d, o_d = util.dev(agent), util.dev(goal), util.dev(obs_nodes)
    st = _lib.DgppoStateRecord(util.p(a_d), util.p(o_d), util.p(g_d))
    # ---- policy
    tree = P.init_policy_params(cfg.node_dim, 4, 2, 2, seed=1, jitter=0.2)
    net = P.net_cfg(_lib.NET_POLICY, cfg.node_dim, 4, 2, 2)
    pk = P.pack_params(tree, net)
    act, lp, h = util.k_policy(cfg, net, pk, graph, rnn, eps)
    pk_d, rnn_d, eps_d = util.dev(pk), util.dev(rnn), util.dev(eps)
    out_h = torch.empty_like(rnn_d)
    out_a = torch.empty((b, n, 2), device="cuda")
    out_lp = torch.empty((b, n), device="cuda")
    _lib.check(_lib.lib().dgppo_gnn_policy_from_state(util.stream(), C.byref(cc), C.byref(net), util.p(pk_d), C.byref(st), 1,
                                                       util.p(rnn_d), util.p(out_h), 1, util.p(eps_d), 1, util.p(out_a),
                                                       util.p(out_lp), 1, b), "policy_from_state")
    torch.cuda.synchronize()
    util.assert_bits_equal(out_a.cpu().numpy(), act, f"{name} action")
    util.assert_bits_equal(out_lp.cpu().numpy(), lp, f"{name} log_pi")
    util.assert_bits_equal(out_h.cpu().numpy(), h, f"{name} carry")
    # ---- Vh (1 layer) and Vl (2 layers, pooled)
    for kind, layers, n_out in ((_lib.NET_VH, 1, cfg.n_cost), (_lib.NET_VL, 2, 1)):
        vt = P.init_value_params(cfg.node_dim, 4, n_out, layers, seed=2, jitter=0.2)
        vnet = P.net_cfg(kind, cfg.node_dim, 4, layers, n_out)
        vpk = P.pack_params(vt, vnet)
        vl = kind == _lib.NET_VL
        r_in = (rng.standard_normal((b, 64) if vl else (b, n, 64)) * 0.3).astype(F)
        ref = util.k_value(cfg, vnet, vpk, graph, r_in)
        vpk_d, r_d = util.dev(vpk), util.dev(r_in)
        scratch = torch.empty_like(r_d)
        val = torch.empty((b,) if vl else (b, n, n_out), device="cuda")
        _lib.check(_lib.lib().dgppo_gnn_value_from_state(util.stream(), C.byref(cc), C.byref(vnet), util.p(vpk_d),
                                                          C.byref(st), 1, util.p(r_d), util.p(scratch), 1, util.p(val),
                                                          1, 1, b), "value_from_state")
        torch.cuda.synchronize()
        ref_v = ref[0] if isinstance(ref, tuple) else ref
        util.assert_bits_equal(val.cpu().numpy(), ref_v, f"{name} value kind {kind}")


def test_from_state_rejects_missing_pieces():
    cfg = CONFIGS["C3"]
    cc = util.c_cfg(cfg)
    net = P.net_cfg(_lib.NET_POLICY, cfg.node_dim, 4, 2, 2)
    x = torch.zeros(64 * 64, device="cuda")
    st = _lib.DgppoStateRecord(util.p(x), None, util.p(x))           # hits missing although n_obs > 0
    rc = _lib.lib().dgppo_gnn_policy_from_state(util.stream(), C.byref(cc), C.byref(net), util.p(x), C.byref(st), 1,
                                                 util.p(x), util.p(x), 1, None, 1, util.p(x), None, 1, 4)
    assert rc == -1
    assert _lib.lib().dgppo_gnn_policy_from_state(util.stream(), C.byref(cc), C.byref(net), util.p(x), None, 1,
                                                   util.p(x), util.p(x), 1, None, 1, util.p(x), None, 1, 4) == -1
