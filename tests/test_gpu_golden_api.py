"""GPU: (1) the CUDA kernels against the REFERENCE-generated fixtures
(tests/golden/ref_*.npz, produced by running the reference's own source);
(2) the reference-shaped Python API (env.reset/step/get_graph/get_cost,
algo.act/step/collect/update) end to end."""
import numpy as np
import pytest
import torch

from oracle import algo_np, env_np, nn_np
from tests import golden_util as G
from tests import util
from tests.util import assert_bits_equal

pytestmark = pytest.mark.gpu
F = np.float32


@pytest.mark.parametrize("name", list(G.CASES))
def test_kernels_match_reference_fixtures(name):
    cfg, d, obstacles = G.load(name)
    T = d["action"].shape[1]
    rays = env_np.ray_table(cfg.n_rays, cfg.comm_radius) if cfg.is_lidar else None
    for t in range(T):
        g = G.graph_at(d, t)
        agent, goal, obs_nodes = env_np.graph_slices(cfg, g)
        nxt, rew, cost = util.k_env_step(cfg, np.ascontiguousarray(agent), np.ascontiguousarray(goal),
                                         None if obs_nodes is None else np.ascontiguousarray(obs_nodes),
                                         d["action"][:, t])
        if cfg.is_lidar:
            nobs = util.k_lidar(cfg, nxt, obstacles, rays) if cfg.n_obs > 0 else None
        else:
            nobs = np.ascontiguousarray(obs_nodes)
        kg = util.k_graph(cfg, nxt, np.ascontiguousarray(goal), nobs)
        ref = G.graph_at(d, t + 1)
        assert_bits_equal(cost, d["cost"][:, t], f"{name} cost t={t}")
        np.testing.assert_allclose(rew, d["reward"][:, t], rtol=2e-6, atol=1e-9)
        for k in ("receivers", "senders", "node_type", "n_node", "n_edge"):
            assert_bits_equal(kg[k], ref[k], f"{name} {k} t={t + 1}")
        for k in ("nodes", "edges", "states"):
            if cfg.is_bicycle:      # CUDA libm vs NumPy libm in atan2 / sin / cos
                np.testing.assert_allclose(kg[k], ref[k], rtol=1e-5, atol=1e-6, err_msg=f"{name} {k} t={t + 1}")
            else:
                assert_bits_equal(kg[k], ref[k], f"{name} {k} t={t + 1}")


def test_gae_kernel_matches_reference_fixture():
    from dgppo_b200 import _lib
    d = np.load(G.GOLDEN_DIR + "/ref_gae.npz")
    for i in range(3):
        hs, l, Vh, Vl = (d[f"c{i}_{k}"] for k in ("hs", "l", "Vh", "Vl"))
        T, a, nh = hs.shape
        th, tl, tVh, tVl = util.dev(hs[None]), util.dev(l[None]), util.dev(Vh[None]), util.dev(Vl[None])
        Qh = torch.empty((1, T, a, nh), device="cuda")
        Ql = torch.empty((1, T), device="cuda")
        assert _lib.lib().dgppo_gae(util.stream(), util.p(th), util.p(tl), util.p(tVh), util.p(tVl), 0.99, 0.95,
                                    util.p(Qh), util.p(Ql), 1, T, a, nh) == 0
        torch.cuda.synchronize()
        np.testing.assert_allclose(Qh[0].cpu().numpy(), d[f"c{i}_Qh"], rtol=1e-5, atol=2e-6)
        np.testing.assert_allclose(Ql[0].cpu().numpy(), d[f"c{i}_Ql"], rtol=1e-5, atol=2e-6)


def _np_graph(g):
    return {k: getattr(g, k).cpu().numpy() for k in G.GRAPH_FIELDS}


def _np_obstacles(rect):
    r = rect.record.cpu().numpy()
    return dict(center=r[..., 0:2], width=r[..., 2], height=r[..., 3], theta=r[..., 4], cos=r[..., 5], sin=r[..., 6],
                points=r[..., 8:16].reshape(r.shape[:-1] + (4, 2)))


@pytest.mark.parametrize("env_id,n,obs", [("LidarSpread", 3, 3), ("LidarTarget", 4, 2), ("LidarBicycleTarget", 4, 3),
                                          ("MPESpread", 5, 3), ("MPETarget", 6, 3), ("MPECorridor", 5, 2),
                                          ("LidarLine", 4, 3), ("MPELine", 3, 3), ("MPELine", 6, 2),
                                          ("MPEFormation", 5, 3), ("MPEConnectSpread", 3, 1)])
def test_env_api_reset_step(env_id, n, obs):
    from dgppo_b200.env import make_env
    env = make_env(env_id, num_agents=n, num_obs=obs)
    cfg = env_np.EnvCfg(env_np.KIND_BY_NAME[env_id], n=n, n_obs=obs, area=env.area_size,
                        obs_radius=env.params.get("obs_radius", 0.05))
    assert env.n_cost == cfg.n_cost and env.num_goals == cfg.n_goal
    g = env.reset(np.arange(10))
    assert g.nodes.shape == (10, cfg.n_nodes, cfg.node_dim) and g.receivers.dtype == torch.int32
    # reset honours the reference's rejection rules (env/utils.py:169-204)
    pos = g.env_states.agent[..., :2].cpu().numpy()
    dmin = np.linalg.norm(pos[:, :, None] - pos[:, None], axis=-1) + np.eye(n) * 10
    min_dist = (2.2 if env_id.startswith("Lidar") and env_id != "LidarLine" else 2.0) * 0.05
    assert (dmin > min_dist).all()
    if env_id.startswith("Lidar"):
        ob = _np_obstacles(g.env_states.obstacle)
        assert not env_np.rect_inside(pos, ob, 0.05 if env_id == "LidarLine" else min_dist / 2).any()
    else:
        ob = None
    action = torch.rand((10, n, 2), device="cuda") * 2.4 - 1.2
    res = env.step(g, action)
    ref_g, r, c, done = env_np.env_step(cfg, _np_graph(g), action.cpu().numpy(), ob)
    bic = cfg.is_bicycle
    for k in ("receivers", "senders", "node_type"):
        assert_bits_equal(getattr(res.graph, k).cpu().numpy(), ref_g[k], k)
    for k in ("nodes", "edges", "states"):
        if bic:
            np.testing.assert_allclose(getattr(res.graph, k).cpu().numpy(), ref_g[k], rtol=1e-5, atol=1e-6)
        else:
            assert_bits_equal(getattr(res.graph, k).cpu().numpy(), ref_g[k], k)
    assert_bits_equal(res.cost.cpu().numpy(), c, "cost")
    assert_bits_equal(res.reward.cpu().numpy(), r, "reward")
    assert res.done.dtype == torch.bool and not res.done.any() and res.info == {}
    assert_bits_equal(env.get_cost(g).cpu().numpy(), c, "get_cost")
    # single-env (unbatched) calls, as the reference's API is written
    g1 = env.reset(3)
    assert g1.is_single and g1.nodes.shape == (cfg.n_nodes, cfg.node_dim)
    r1 = env.step(g1, action[0])
    assert r1.graph.is_single and r1.reward.shape == () and r1.cost.shape == (n, cfg.n_cost)


def test_algo_api_act_step_collect_update():
    from dgppo_b200.algo import make_algo
    from dgppo_b200.env import make_env
    n, T, b = 3, 16, 12
    env = make_env("LidarSpread", num_agents=n, num_obs=3, max_step=T)
    cfg = env_np.EnvCfg(env_np.LIDAR_SPREAD, n=n, n_obs=3)
    algo = make_algo("dgppo", env=env, node_dim=env.node_dim, edge_dim=env.edge_dim, state_dim=env.state_dim,
                     action_dim=env.action_dim, n_agents=n, batch_size=b * T, rnn_step=16, seed=3)
    assert algo.init_rnn_state.shape == (1, n, 1, 64) and set(algo.params) == {"policy", "Vl", "Vh"}
    g = env.reset(np.arange(b))
    rnn0 = algo.init_rnn_state.expand(b, 1, n, 1, 64).contiguous()
    # act: deterministic mode
    a, rnn1 = algo.act(g, rnn0)
    ra, _, rh, _ = nn_np.policy_forward(algo.params["policy"], _np_graph(g), np.zeros((b, n, 64), F), n, None)
    np.testing.assert_allclose(a.cpu().numpy(), ra, rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(rnn1.reshape(b, n, 64).cpu().numpy(), rh, rtol=1e-5, atol=1e-5)
    assert rnn1.shape == rnn0.shape
    # step: explicit draws
    eps = torch.randn((b, n, 2), device="cuda")
    a, lp, _ = algo.step(g, rnn0, eps)
    ra, rlp, _, _ = nn_np.policy_forward(algo.params["policy"], _np_graph(g), np.zeros((b, n, 64), F), n,
                                         eps.cpu().numpy())
    np.testing.assert_allclose(a.cpu().numpy(), ra, rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(lp.cpu().numpy(), rlp, rtol=1e-5, atol=2e-5)
    a2, lp2, _ = algo.step(g, rnn0, key=7)
    a3, lp3, _ = algo.step(g, rnn0, key=7)
    assert torch.equal(a2, a3) and not torch.equal(a2, a)
    # collect -> Rollout with the reference's field layout
    ro = algo.collect(algo.params, np.arange(b))
    assert ro.actions.shape == (b, T, n, 2) and ro.rnn_states.shape == (b, T, 1, n, 1, 64)
    assert ro.rewards.shape == (b, T) and ro.costs.shape == (b, T, n, 2) and ro.log_pis.shape == (b, T, n)
    assert ro.dones.dtype == torch.bool and not ro.dones.any()
    assert ro.graph.nodes.shape == (b, T, cfg.n_nodes, 7) and ro.next_graph.nodes.shape == (b, T, cfg.n_nodes, 7)
    assert torch.equal(ro.graph.nodes[:, 1:], ro.next_graph.nodes[:, :-1])      # next_graph[t] == graph[t+1]
    assert (ro.rnn_states[:, 0] == 0).all()
    # update pre-pass: Vh, GAE, advantage against the oracle
    pp = algo.prepass(ro, step=0)          # (update() = this pre-pass + the PPO minibatch scan, which moves the weights)
    info = {"eval/safe_data": float(pp["bTa_is_safe"].float().mean())}
    Vh, Vl, Qh, Ql, A = (pp[k].cpu().numpy() for k in ("bTp1ah_Vh", "bTp1_Vl", "bTah_Qh", "bT_Ql", "bTa_A"))
    assert Vh.shape == (b, T + 1, n, 2) and Vl.shape == (b, T + 1) and A.shape == (b, T, n)
    gt = {k: getattr(ro.graph, k)[:, 5].cpu().numpy() for k in G.GRAPH_FIELDS}
    rVh = nn_np.vh_forward(algo.params["Vh"], gt, ro.rnn_states[:, 5].reshape(b, n, 64).cpu().numpy(), n)
    np.testing.assert_allclose(Vh[:, 5], rVh, rtol=1e-5, atol=1e-5)
    # Vl is recurrent: replay its scan with the oracle
    h = np.zeros((b, 64), F)
    for t in range(T + 1):
        src = ro.graph if t < T else ro.next_graph
        gt = {k: getattr(src, k)[:, min(t, T - 1) if t < T else T - 1].cpu().numpy() for k in G.GRAPH_FIELDS}
        v, h = nn_np.vl_forward(algo.params["Vl"], gt, h, n)
        np.testing.assert_allclose(Vl[:, t], v, rtol=2e-5, atol=2e-5, err_msg=f"Vl t={t}")
    for i in range(3):
        rQh, rQl = algo_np.compute_dec_ocp_gae(ro.costs[i].cpu().numpy(), -ro.rewards[i].cpu().numpy(), Vh[i], Vl[i],
                                               0.99, 0.95)
        np.testing.assert_allclose(Qh[i], rQh, rtol=1e-5, atol=5e-6)
        np.testing.assert_allclose(Ql[i], rQl, rtol=1e-5, atol=5e-6)
    rA, rd, _, rsafe = algo_np.cbf_advantage(Ql, Vl, Vh, env.dt, algo.alpha, algo.cbf_eps, algo.cbf_weight)
    near = (np.abs(rd) < 1e-4).any(-1)
    np.testing.assert_allclose(A[~near], rA[~near], rtol=1e-4, atol=1e-4)
    assert 0.0 <= info["eval/safe_data"] <= 1.0
    det = pp["det_rollout"]
    assert det.log_pis is None and det.rnn_states.shape == (b, T, 1, n, 1, 64)
    # test_rollout stores the POST-step carry (trainer/utils.py:73-77)
    _, h1 = algo.act(GraphsTupleAt(det.graph, 0), rnn0)
    np.testing.assert_allclose(det.rnn_states[:, 0].reshape(b, n, 64).cpu().numpy(),
                               h1.reshape(b, n, 64).cpu().numpy(), rtol=1e-5, atol=1e-6)
    # save / load round trip
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        algo.save(td, 0)
        algo2 = make_algo("dgppo", env=env, node_dim=7, edge_dim=4, state_dim=4, action_dim=2, n_agents=n, seed=99)
        algo2.load(td, 0)
        a4, _ = algo2.act(g, rnn0)
        a5, _ = algo.act(g, rnn0)
        assert torch.equal(a4, a5)


def GraphsTupleAt(g, t):
    from dgppo_b200.utils.graph import GraphsTuple
    return GraphsTuple(*[x[:, t].contiguous() if isinstance(x, torch.Tensor) else None for x in g])


@pytest.mark.parametrize("name", list(G.NN_CASES))
@pytest.mark.parametrize("gen", ["v2", "v1"])
def test_network_kernels_match_reference_fixtures(name, gen, monkeypatch):
    """CUDA policy / Vh / Vl forward vs outputs of the reference's own module code."""
    from dgppo_b200 import _lib
    from dgppo_b200.algo import params as P
    monkeypatch.setenv("DGPPO_FORCE_V1", "1" if gen == "v1" else "0")
    cfg, d, graph, trees = G.load_nn(name)
    nd = G.NN_CASES[name]
    pc = P.net_cfg(_lib.NET_POLICY, nd, 4, 2, 2)
    a, _, h = util.k_policy(cfg, pc, P.pack_params(trees["policy"], pc), graph, d["rnn"], None)
    np.testing.assert_allclose(a, d["act_mode"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(h, d["rnn_out"], rtol=1e-5, atol=1e-5)
    a, lp, _ = util.k_policy(cfg, pc, P.pack_params(trees["policy"], pc), graph, d["rnn"], d["eps"])
    np.testing.assert_allclose(a, d["act_sample"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(lp, d["log_pi"], rtol=1e-5, atol=2e-5)
    vc = P.net_cfg(_lib.NET_VH, nd, 4, 1, 2)
    vh, _ = util.k_value(cfg, vc, P.pack_params(trees["vh"], vc), graph, d["rnn"])
    np.testing.assert_allclose(vh, d["vh"], rtol=1e-5, atol=1e-5)
    lc = P.net_cfg(_lib.NET_VL, nd, 4, 2, 1)
    vl, vlh = util.k_value(cfg, lc, P.pack_params(trees["vl"], lc), graph, d["vl_rnn"])
    np.testing.assert_allclose(vl, d["vl"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(vlh, d["vl_rnn_out"], rtol=1e-5, atol=1e-5)
