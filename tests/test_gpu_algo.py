"""GPU parity of K5 (GAE), the CBF advantage merge and the fused rollout."""
import ctypes as C

import numpy as np
import pytest
import torch

from dgppo_b200 import _lib
from dgppo_b200.algo import params as P
from oracle import algo_np, env_np, nn_np
from tests import util
from tests.util import CONFIGS, assert_bits_equal, dev, p, stream

pytestmark = pytest.mark.gpu
F = np.float32


def _gae_inputs(b, T, n, nh, seed):
    rng = np.random.default_rng(seed)
    hs = np.clip(rng.normal(-0.7, 0.3, (b, T, n, nh)), -1, 1).astype(F)
    l = rng.uniform(0, 0.02, (b, T)).astype(F)
    Vh = rng.normal(0, 0.5, (b, T + 1, n, nh)).astype(F)
    Vl = rng.normal(0, 0.5, (b, T + 1)).astype(F)
    return hs, l, Vh, Vl


@pytest.mark.parametrize("b,T,n,nh", [(6, 128, 8, 2), (3, 16, 3, 2), (2, 128, 64, 2), (5, 32, 1, 1), (9, 7, 16, 3)])
def test_gae(b, T, n, nh):
    hs, l, Vh, Vl = _gae_inputs(b, T, n, nh, 0)
    Qh = torch.empty((b, T, n, nh), device="cuda")
    Ql = torch.empty((b, T), device="cuda")
    d_hs, d_l, d_Vh, d_Vl = dev(hs), dev(l), dev(Vh), dev(Vl)      # keep the device tensors alive
    rc = _lib.lib().dgppo_gae(stream(), p(d_hs), p(d_l), p(d_Vh), p(d_Vl), 0.99, 0.95, p(Qh), p(Ql),
                              b, T, n, nh)
    assert rc == 0
    torch.cuda.synchronize()
    for i in range(b):
        rQh, rQl = algo_np.compute_dec_ocp_gae(hs[i], l[i], Vh[i], Vl[i], 0.99, 0.95)
        np.testing.assert_allclose(Qh[i].cpu().numpy(), rQh, rtol=1e-5, atol=2e-6)
        np.testing.assert_allclose(Ql[i].cpu().numpy(), rQl, rtol=1e-5, atol=2e-6)
    # size-independent property: the closed forms
    cQh, cQl = algo_np.gae_closed_form(hs[0], l[0], Vh[0], Vl[0], 0.99, 0.95)
    np.testing.assert_allclose(Qh[0].cpu().numpy(), cQh, rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(Ql[0].cpu().numpy(), cQl, rtol=1e-4, atol=1e-5)


def test_cbf_advantage():
    b, T, n, nh = 7, 128, 8, 2
    rng = np.random.default_rng(1)
    Ql = rng.normal(0, 1, (b, T)).astype(F)
    Vl = rng.normal(0, 1, (b, T + 1)).astype(F)
    Vh = rng.normal(-0.3, 0.3, (b, T + 1, n, nh)).astype(F)
    Vh[:, 1:] = Vh[:, :-1] * 0.9 + 0.01 * rng.normal(size=(b, T, n, nh)).astype(F)   # many safe entries
    A = torch.empty((b, T, n), device="cuda")
    d = torch.empty((b, T, n, nh), device="cuda")
    ac = torch.empty_like(d)
    sf = torch.empty((b, T, n), dtype=torch.uint8, device="cuda")
    d_Ql, d_Vl, d_Vh = dev(Ql), dev(Vl), dev(Vh)                    # keep the device tensors alive
    rc = _lib.lib().dgppo_cbf_advantage(stream(), p(d_Ql), p(d_Vl), p(d_Vh), 0.03, 10.0, 1e-2, 2.0,
                                        p(A), p(d), p(ac), p(sf), b, T, n, nh)
    assert rc == 0
    torch.cuda.synchronize()
    rA, rd, rac, rsf = algo_np.cbf_advantage(Ql, Vl, Vh, 0.03, 10.0, 1e-2, 2.0)
    np.testing.assert_allclose(d.cpu().numpy(), rd, rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(ac.cpu().numpy(), rac, rtol=1e-5, atol=1e-5)
    near = (np.abs(rd) < 1e-5).any(-1)
    assert (sf.cpu().numpy().astype(bool) == rsf)[~near].all()
    np.testing.assert_allclose(A.cpu().numpy()[~near], rA[~near], rtol=2e-5, atol=2e-5)
    assert rsf.any() and (~rsf).any()


def test_cbf_advantage_and_gae_propagate_nan():
    """jnp.maximum / jnp.max propagate NaN (dgppo.py:246,255, algo/utils.py:39-41): a NaN Vh must reach
    Acbf, A and Qh instead of being dropped by the max."""
    b, T, n, nh = 2, 16, 3, 2
    hs, l, Vh, Vl = _gae_inputs(b, T, n, nh, 3)
    Vh[1, 5, 2, 1] = np.nan
    Ql = np.zeros((b, T), F)
    A = torch.empty((b, T, n), device="cuda")
    d = torch.empty((b, T, n, nh), device="cuda")
    ac = torch.empty_like(d)
    sf = torch.empty((b, T, n), dtype=torch.uint8, device="cuda")
    d_Ql, d_Vl, d_Vh, d_hs, d_l = dev(Ql), dev(Vl), dev(Vh), dev(hs), dev(l)
    assert _lib.lib().dgppo_cbf_advantage(stream(), p(d_Ql), p(d_Vl), p(d_Vh), 0.03, 10.0, 1e-2, 1.0,
                                          p(A), p(d), p(ac), p(sf), b, T, n, nh) == 0
    Qh = torch.empty((b, T, n, nh), device="cuda")
    Qlo = torch.empty((b, T), device="cuda")
    assert _lib.lib().dgppo_gae(stream(), p(d_hs), p(d_l), p(d_Vh), p(d_Vl), 0.99, 0.95, p(Qh), p(Qlo),
                                b, T, n, nh) == 0
    torch.cuda.synchronize()
    rA, rd, rac, rsf = algo_np.cbf_advantage(Ql, Vl, Vh, 0.03, 10.0, 1e-2, 1.0)
    assert np.isnan(rA).any()
    np.testing.assert_array_equal(np.isnan(A.cpu().numpy()), np.isnan(rA))
    np.testing.assert_array_equal(np.isnan(ac.cpu().numpy()), np.isnan(rac))
    np.testing.assert_array_equal(sf.cpu().numpy().astype(bool)[np.isnan(rA)], rsf[np.isnan(rA)])
    rQh, _ = algo_np.compute_dec_ocp_gae(hs[1], l[1], Vh[1], Vl[1], 0.99, 0.95)
    assert np.isnan(rQh).any()
    np.testing.assert_array_equal(np.isnan(Qh[1].cpu().numpy()), np.isnan(rQh))
    assert not np.isnan(Qh[0].cpu().numpy()).any()


@pytest.mark.parametrize("name", ["C1", "C2", "C3", "C4"])
@pytest.mark.parametrize("stochastic", [True, False])
def test_rollout_per_step_parity(name, stochastic):
    """dgppo_rollout for T steps; every step is then re-derived by the oracle from
    the record's own graph[t] / rnn[t] / action[t]: env results bit-exact (exact
    for the bicycle up to libm), policy outputs within rtol 1e-5."""
    cfg = CONFIGS[name]
    b, T = 24, 12
    agent, goal, obstacles, mpe_obs = env_np.synthetic_states(cfg, b, 17)
    rays = env_np.ray_table(cfg.n_rays, cfg.comm_radius) if cfg.is_lidar else None
    g0 = env_np.reset_graph(cfg, agent, goal, obstacles, mpe_obs, rays)
    tree = P.init_policy_params(cfg.node_dim, 4, 2, 2, seed=3, jitter=0.1, scale_final=1.0)
    nc = P.net_cfg(_lib.NET_POLICY, cfg.node_dim, 4, 2, 2)
    packed = dev(P.pack_params(tree, nc))
    eps = np.random.default_rng(8).standard_normal((b, T, cfg.n, 2)).astype(F) if stochastic else None
    N, E, n, sd, nd = cfg.n_nodes, cfg.n_edges, cfg.n, cfg.state_dim, cfg.node_dim
    Pp = T + 1
    f32 = dict(dtype=torch.float32, device="cuda")
    i32 = dict(dtype=torch.int32, device="cuda")
    rec = dict(nodes=torch.zeros((b, Pp, N, nd), **f32), edges=torch.zeros((b, Pp, E, 4), **f32),
               states=torch.zeros((b, Pp, N, sd), **f32), receivers=torch.zeros((b, Pp, E), **i32),
               senders=torch.zeros((b, Pp, E), **i32), node_type=torch.zeros((b, Pp, N), **i32),
               n_node=torch.zeros((b, Pp), **i32), n_edge=torch.zeros((b, Pp), **i32))
    for k in rec:
        rec[k][:, 0] = dev(g0[k])
    rnn = torch.zeros((b, Pp, n, 64), **f32)
    actions = torch.zeros((b, T, n, 2), **f32)
    log_pis = torch.zeros((b, T, n), **f32)
    rewards = torch.zeros((b, T), **f32)
    costs = torch.zeros((b, T, n, 2), **f32)
    agent_ws = torch.zeros((2, b, n, sd), **f32)
    agent_ws[0] = dev(agent)
    hits_ws, obst, rays_d = None, None, None
    if cfg.is_lidar and cfg.n_obs > 0:
        hits_ws = dev(env_np.graph_slices(cfg, g0)[2])
        obst = dev(util.obs_record(obstacles))
        rays_d = dev(rays)
    elif cfg.n_obs > 0:
        obst = dev(mpe_obs)
    e_d, goal_d = dev(eps), dev(goal)
    buf = _lib.DgppoRolloutBuffers(p(rec["nodes"]), p(rec["edges"]), p(rec["states"]), p(rec["receivers"]),
                                   p(rec["senders"]), p(rec["node_type"]), p(rec["n_node"]), p(rec["n_edge"]),
                                   p(rnn), p(e_d), p(actions), p(log_pis) if stochastic else None, p(rewards),
                                   p(costs), p(agent_ws), p(hits_ws), p(goal_d), p(obst), p(rays_d))
    cc = util.c_cfg(cfg)
    rc = _lib.lib().dgppo_rollout(stream(), C.byref(cc), C.byref(nc), p(packed), C.byref(buf), T, b, None)
    assert rc == 0
    torch.cuda.synchronize()
    R = {k: v.cpu().numpy() for k, v in rec.items()}
    rnn_h, act_h, lp_h = rnn.cpu().numpy(), actions.cpu().numpy(), log_pis.cpu().numpy()
    rew_h, cost_h = rewards.cpu().numpy(), costs.cpu().numpy()
    assert (R["n_node"] == N).all() and (R["n_edge"] == E).all()
    for t in range(T):
        g_t = {k: v[:, t] for k, v in R.items()}
        e_t = None if eps is None else eps[:, t]
        a, lp, h, _ = nn_np.policy_forward(tree, g_t, rnn_h[:, t], n, e_t, 2, np.float32)
        np.testing.assert_allclose(act_h[:, t], a, rtol=1e-5, atol=1e-5, err_msg=f"action t={t}")
        np.testing.assert_allclose(rnn_h[:, t + 1], h, rtol=1e-5, atol=1e-5, err_msg=f"rnn t={t}")
        if stochastic:
            np.testing.assert_allclose(lp_h[:, t], lp, rtol=1e-5, atol=2e-5, err_msg=f"log_pi t={t}")
        g_n, r, c, _ = env_np.env_step(cfg, g_t, act_h[:, t], obstacles, rays)
        assert_bits_equal(rew_h[:, t], r, f"reward t={t}")
        assert_bits_equal(cost_h[:, t], c, f"cost t={t}")
        for k in ("receivers", "senders", "node_type"):
            assert_bits_equal(R[k][:, t + 1], g_n[k], f"{k} t={t + 1}")
        for k in ("nodes", "edges", "states"):
            if cfg.is_bicycle:
                np.testing.assert_allclose(R[k][:, t + 1], g_n[k], rtol=1e-5, atol=1e-6, err_msg=f"{k} t={t + 1}")
            else:
                assert_bits_equal(R[k][:, t + 1], g_n[k], f"{k} t={t + 1}")


@pytest.mark.parametrize("name,b,T", [("C3", 19, 5), ("C2", 7, 4), ("C4", 5, 3), ("C5", 3, 2), ("n24", 5, 3)])
def test_rollout_writes_stay_inside_the_record(name, b, T):
    """Every output buffer of dgppo_rollout is allocated with one guard environment in front and one
    behind, filled with sentinels; the kernels (slot / pitch addressing, tile tails, row chunks of the
    large-graph GNN kernel) must leave the guards untouched and overwrite every element in between."""
    cfg = CONFIGS[name]
    agent, goal, obstacles, mpe_obs = env_np.synthetic_states(cfg, b, 23)
    rays = env_np.ray_table(cfg.n_rays, cfg.comm_radius) if cfg.is_lidar else None
    g0 = env_np.reset_graph(cfg, agent, goal, obstacles, mpe_obs, rays)
    tree = P.init_policy_params(cfg.node_dim, 4, 2, 2, seed=3, jitter=0.1, scale_final=1.0)
    nc = P.net_cfg(_lib.NET_POLICY, cfg.node_dim, 4, 2, 2)
    packed = dev(P.pack_params(tree, nc))
    eps = np.random.default_rng(8).standard_normal((b, T, cfg.n, 2)).astype(F)
    N, E, n, sd, nd = cfg.n_nodes, cfg.n_edges, cfg.n, cfg.state_dim, cfg.node_dim
    Pp = T + 1
    SF, SI = float("nan"), -77777
    full = {}

    def guarded(key, shape, dtype):
        t = torch.full((b + 2,) + tuple(shape), SI if dtype == torch.int32 else SF, dtype=dtype, device="cuda")
        full[key] = t
        return t[1:b + 1]

    rec = dict(nodes=guarded("nodes", (Pp, N, nd), torch.float32), edges=guarded("edges", (Pp, E, 4), torch.float32),
               states=guarded("states", (Pp, N, sd), torch.float32),
               receivers=guarded("receivers", (Pp, E), torch.int32), senders=guarded("senders", (Pp, E), torch.int32),
               node_type=guarded("node_type", (Pp, N), torch.int32), n_node=guarded("n_node", (Pp,), torch.int32),
               n_edge=guarded("n_edge", (Pp,), torch.int32))
    for k in rec:
        rec[k][:, 0] = dev(g0[k])
    rnn = guarded("rnn", (Pp, n, 64), torch.float32)
    rnn[:, 0] = 0.0
    actions = guarded("actions", (T, n, 2), torch.float32)
    log_pis = guarded("log_pis", (T, n), torch.float32)
    rewards = guarded("rewards", (T,), torch.float32)
    costs = guarded("costs", (T, n, 2), torch.float32)
    agent_ws = torch.zeros((2, b, n, sd), dtype=torch.float32, device="cuda")
    agent_ws[0] = dev(agent)
    hits_ws, obst, rays_d = None, None, None
    if cfg.is_lidar and cfg.n_obs > 0:
        hits_ws = dev(env_np.graph_slices(cfg, g0)[2])
        obst = dev(util.obs_record(obstacles))
        rays_d = dev(rays)
    elif cfg.n_obs > 0:
        obst = dev(mpe_obs)
    e_d, goal_d = dev(eps), dev(goal)
    buf = _lib.DgppoRolloutBuffers(p(rec["nodes"]), p(rec["edges"]), p(rec["states"]), p(rec["receivers"]),
                                   p(rec["senders"]), p(rec["node_type"]), p(rec["n_node"]), p(rec["n_edge"]),
                                   p(rnn), p(e_d), p(actions), p(log_pis), p(rewards), p(costs), p(agent_ws),
                                   p(hits_ws), p(goal_d), p(obst), p(rays_d))
    cc = util.c_cfg(cfg)
    rc = _lib.lib().dgppo_rollout(stream(), C.byref(cc), C.byref(nc), p(packed), C.byref(buf), T, b, None)
    assert rc == 0
    torch.cuda.synchronize()
    for key, t in full.items():
        h = t.cpu().numpy()
        for guard in (h[0], h[-1]):
            if h.dtype == np.int32:
                assert (guard == SI).all(), f"{key}: guard environment overwritten"
            else:
                assert np.isnan(guard).all(), f"{key}: guard environment overwritten"
        inner = h[1:-1]
        if h.dtype == np.int32:
            assert (inner != SI).all(), f"{key}: elements left unwritten"
        else:
            assert not np.isnan(inner).any(), f"{key}: elements left unwritten (or NaN produced)"


@pytest.mark.parametrize("env_id,n,obs", [("LidarSpread", 8, 8), ("LidarBicycleTarget", 4, 3)])
def test_lidar_look_ahead_schedule_is_bit_identical(env_id, n, obs, monkeypatch):
    """DGPPO_LIDAR_AHEAD=1 casts the rays of graph t+1 from the predicted position on a side stream while the
    policy of step t runs (and DGPPO_ROLLOUT_THREADS=1 submits the env groups from threads): same record bits."""
    from dgppo_b200.algo import make_algo
    from dgppo_b200.env import make_env
    b, T = 64, 12
    env = make_env(env_id, num_agents=n, num_obs=obs, max_step=T)
    algo = make_algo("dgppo", env=env, node_dim=env.node_dim, edge_dim=env.edge_dim, state_dim=env.state_dim,
                     action_dim=env.action_dim, n_agents=n, batch_size=b * T, seed=2)
    g0 = env.reset(np.arange(b, dtype=np.uint64) + 50)
    eps = torch.randn((b, T, n, 2), device="cuda", generator=torch.Generator(device="cuda").manual_seed(4))
    ref = algo.collect(algo.params, None, eps=eps, graph0=g0)
    ref_t = {k: getattr(ref.next_graph, k).clone() for k in ("nodes", "edges", "states", "receivers", "senders")}
    ref_a, ref_c, ref_r = ref.actions.clone(), ref.costs.clone(), ref.rewards.clone()
    monkeypatch.setenv("DGPPO_LIDAR_AHEAD", "1")
    monkeypatch.setenv("DGPPO_ROLLOUT_THREADS", "1")
    for chunks in (1, 4):
        algo.rollout_chunks = chunks
        out = algo.collect(algo.params, None, eps=eps, graph0=g0)
        for k, v in ref_t.items():
            assert torch.equal(getattr(out.next_graph, k), v), f"{k} (chunks={chunks})"
        assert torch.equal(out.actions, ref_a) and torch.equal(out.costs, ref_c) and torch.equal(out.rewards, ref_r)
