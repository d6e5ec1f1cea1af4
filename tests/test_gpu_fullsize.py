"""GPU: the rollout at BASELINE.json's full sizes (4096 / 2048 / 1024 environments per GPU).

The oracle cannot replay half a million env-steps, so full-size runs are checked through
  * size-independent properties: run-to-run bit identity, independence of the stream chunking,
    independence of the environment sharding (what multi-GPU weak scaling relies on), graph
    invariants that hold for every slot of the record (symmetry of the radius graph,
    antisymmetry of the edge features, goal edges always on, LiDAR hits sorted by range and
    connected exactly when in range);
  * the oracle on a random SAMPLE of (environment, step) slots: the stored graph[t], rnn[t] and
    action[t] are pushed through the NumPy env step / policy forward and compared with slot
    t+1 (bit-exact; bicycle and networks rtol 1e-5).
"""
import numpy as np
import pytest
import torch

from oracle import env_np, nn_np
from tests.util import CONFIGS, assert_bits_equal

pytestmark = pytest.mark.gpu
F = np.float32
GRAPH_FIELDS = ("nodes", "edges", "states", "receivers", "senders", "node_type", "n_node", "n_edge")

#        name  env id                 n   obs  envs   T    sampled slots
SIZES = [("C2", "MPESpread", 8, 3, 4096, 128, 96),
         ("C3", "LidarSpread", 8, 8, 4096, 128, 96),
         ("C4", "LidarBicycleTarget", 16, 3, 2048, 128, 48),
         ("C5", "LidarSpread", 64, 64, 1024, 128, 6)]


def _setup(env_id, n, obs, b, T, seed):
    from dgppo_b200.algo import make_algo
    from dgppo_b200.env import make_env
    from dgppo_b200.env.envs import LidarEnvState, MPEEnvState, Rectangle
    env = make_env(env_id, num_agents=n, num_obs=obs, max_step=T)
    algo = make_algo("dgppo", env=env, node_dim=env.node_dim, edge_dim=env.edge_dim, state_dim=env.state_dim,
                     action_dim=env.action_dim, n_agents=n, batch_size=min(16384, b * T), seed=seed)
    dev = torch.device("cuda")
    if n >= 64:      # the reset sampler cannot place 64 + 64 + 64 in the default area: synthetic states
        cfg = CONFIGS["C5"]
        agent, goal, obstacles, _ = env_np.synthetic_states(cfg, b, seed)
        from tests.util import obs_record
        rect = Rectangle.from_record(obs_record(obstacles), dev)
        es = LidarEnvState(torch.as_tensor(agent, device=dev), torch.as_tensor(goal, device=dev), rect)
        g0 = env.get_graph(es, env.get_lidar_data(es.agent, rect))
    else:
        g0 = env.reset(np.arange(b, dtype=np.uint64) + 1000 * seed)
    eps = torch.randn((b, T, n, 2), device=dev, generator=torch.Generator(device=dev).manual_seed(seed))
    return env, algo, g0, eps


def _slice_graph0(g0, lo, hi):
    from dgppo_b200.env.envs import LidarEnvState, MPEEnvState, Rectangle
    from dgppo_b200.utils.graph import GraphsTuple
    es = g0.env_states
    if isinstance(es, LidarEnvState):
        ob = None if es.obstacle is None else Rectangle(*[t[lo:hi] for t in es.obstacle])
        es2 = LidarEnvState(es.agent[lo:hi], es.goal[lo:hi], ob)
    else:
        es2 = MPEEnvState(es.agent[lo:hi], es.goal[lo:hi], None if es.obs is None else es.obs[lo:hi])
    parts = [t[lo:hi].contiguous() if isinstance(t, torch.Tensor) else t for t in g0]
    return GraphsTuple(*parts)._replace(env_states=es2)


def _record_equal(a, b, what):
    for k in GRAPH_FIELDS:
        assert torch.equal(getattr(a.graph, k), getattr(b.graph, k)), f"{what}: graph.{k}"
        assert torch.equal(getattr(a.next_graph, k)[:, -1], getattr(b.next_graph, k)[:, -1]), f"{what}: last {k}"
    for k in ("actions", "rewards", "costs", "log_pis", "rnn_states"):
        assert torch.equal(getattr(a, k), getattr(b, k)), f"{what}: {k}"


@pytest.mark.parametrize("name,env_id,n,obs,b,T,n_samples", SIZES)
def test_full_size_rollout(name, env_id, n, obs, b, T, n_samples):
    from dgppo_b200.trainer.rollout import RolloutRecord
    cfg = CONFIGS[name]
    env, algo, g0, eps = _setup(env_id, n, obs, b, T, seed=5)
    dev = eps.device
    ro = algo.collect(algo.params, None, eps=eps, graph0=g0)

    # ---- determinism and independence of the stream chunking
    chunks = algo.rollout_chunks
    try:
        algo.rollout_chunks = 1
        ro1 = algo.collect(algo.params, None, eps=eps, graph0=g0)
    finally:
        algo.rollout_chunks = chunks
    _record_equal(ro, ro1, f"{name}: chunks={chunks} vs 1")
    del ro1
    # ---- independence of the sharding: a quarter of the environments alone gives the same bits
    q = b // 4
    ro_q = algo.collect(algo.params, None, eps=eps[q:2 * q].contiguous(), graph0=_slice_graph0(g0, q, 2 * q))
    for k in ("actions", "rewards", "costs", "log_pis"):
        assert torch.equal(getattr(ro_q, k), getattr(ro, k)[q:2 * q]), f"{name}: shard {k}"
    assert torch.equal(ro_q.next_graph.states[:, -1], ro.next_graph.states[q:2 * q, -1])
    del ro_q

    # ---- invariants over every slot of the record
    N, E, pad = cfg.n_nodes, cfg.n_edges, cfg.n_nodes - 1
    g = ro.graph
    assert bool((g.n_node == N).all()) and bool((g.n_edge == E).all())
    recv, send, edges = g.receivers, g.senders, g.edges
    aa_r = recv[..., :n * n].reshape(b, T, n, n)
    aa_s = send[..., :n * n].reshape(b, T, n, n)
    on = aa_r != pad
    ids = torch.arange(n, device=dev, dtype=aa_r.dtype)
    assert torch.equal(on, on.transpose(-1, -2)), "radius graph is symmetric"
    assert not bool(torch.diagonal(on, dim1=-2, dim2=-1).any()), "no self edges"
    assert bool((aa_r[on] == ids.view(1, 1, n, 1).expand_as(aa_r)[on]).all())
    assert bool((aa_s[on] == ids.view(1, 1, 1, n).expand_as(aa_s)[on]).all())
    assert bool((aa_s[~on] == pad).all())
    ef = edges[..., :n * n, :].reshape(b, T, n, n, 4)
    assert torch.equal(ef, -ef.transpose(2, 3)), "agent-agent edge features are antisymmetric"
    n_ag = cfg.n_ag
    ag_r = recv[..., n * n:n * n + n * n_ag].reshape(b, T, n, n_ag)
    assert bool((ag_r == ids.view(1, 1, n, 1)).all()), "goal edges are always on"
    if cfg.is_lidar and cfg.n_obs > 0:
        k = cfg.top_k
        pos = g.states[..., :n, :2]                                             # (b, T, n, 2)
        hits = g.states[..., 2 * n:2 * n + n * k, :2].reshape(b, T, n, k, 2)
        dist = (hits - pos.unsqueeze(3)).norm(dim=-1)
        # a ray without a hit keeps alpha = 1e6 (hit point half a million units away, env/utils.py:126-136):
        # the top-k are ordered by alpha, i.e. by range up to rounding
        assert bool((dist[..., 1:] >= dist[..., :-1] * (1 - 1e-4) - 1e-5).all()), "top-k hits sorted by range"
        ao_r = recv[..., n * n + n * n_ag:].reshape(b, T, n, k)
        thr = cfg.comm_radius - 0.1                                             # lidar_spread.py:86
        assert bool((ao_r[dist < thr - 1e-5] != pad).all()) and bool((ao_r[dist > thr + 1e-5] == pad).all())
    assert bool(torch.isfinite(ro.rewards).all()) and bool(torch.isfinite(ro.log_pis).all())
    assert bool((ro.actions.abs() <= 1).all())

    # ---- the oracle on a random sample of (environment, step) slots
    rng = np.random.default_rng(11)
    se = torch.as_tensor(rng.integers(0, b, n_samples), device=dev)
    st = torch.as_tensor(rng.integers(0, T, n_samples), device=dev)
    g_t = {k: getattr(ro.graph, k)[se, st].cpu().numpy() for k in GRAPH_FIELDS}
    g_n = {k: getattr(ro.next_graph, k)[se, st].cpu().numpy() for k in GRAPH_FIELDS}
    act = ro.actions[se, st].cpu().numpy()
    rnn_t = ro.rnn_states.reshape(b, T, n, 64)[se, st].cpu().numpy()
    tn = torch.clamp(st + 1, max=T - 1)
    rnn_n = ro.rnn_states.reshape(b, T, n, 64)[se, tn].cpu().numpy()
    e_t = eps[se, st].cpu().numpy()
    a, lp, h, _ = nn_np.policy_forward(algo.params["policy"], g_t, rnn_t, n, e_t, 2, np.float32)
    np.testing.assert_allclose(act, a, rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(ro.log_pis[se, st].cpu().numpy(), lp, rtol=1e-5, atol=2e-5)
    inner = (st + 1 < T).cpu().numpy()
    np.testing.assert_allclose(rnn_n[inner], h[inner], rtol=1e-5, atol=1e-5)
    obstacles, rays = None, None
    if cfg.is_lidar and cfg.n_obs > 0:
        rec = g0.env_states.obstacle.record[se].cpu().numpy()
        obstacles = dict(center=rec[..., 0:2], width=rec[..., 2], height=rec[..., 3], theta=rec[..., 4],
                         cos=rec[..., 5], sin=rec[..., 6], points=rec[..., 8:16].reshape(rec.shape[0], -1, 4, 2))
        rays = env_np.ray_table(cfg.n_rays, cfg.comm_radius)
    r_next, r_rew, r_cost, _ = env_np.env_step(cfg, g_t, act, obstacles, rays)
    assert_bits_equal(ro.rewards[se, st].cpu().numpy(), r_rew, "reward")
    assert_bits_equal(ro.costs[se, st].cpu().numpy(), r_cost, "cost")
    for k in ("receivers", "senders", "node_type"):
        assert_bits_equal(g_n[k], r_next[k], k)
    for k in ("nodes", "edges", "states"):
        if cfg.is_bicycle:
            np.testing.assert_allclose(g_n[k], r_next[k], rtol=1e-5, atol=1e-6, err_msg=k)
        else:
            assert_bits_equal(g_n[k], r_next[k], k)


def test_full_size_update_prepass():
    """algo.update() at the headline size (C3: 4096 envs x T=128): Vl scan, Vh over all (b, T+1) graphs of both
    records, both GAE passes and the CBF advantage merge.  Checked by properties over all entries and by the
    oracle on sampled environments / slots (dgppo.py:204-273)."""
    from oracle import algo_np
    name, env_id, n, obs, b, T, _ = SIZES[1]
    cfg = CONFIGS[name]
    env, algo, g0, eps = _setup(env_id, n, obs, b, T, seed=9)
    ro = algo.collect(algo.params, None, eps=eps, graph0=g0)
    pp = algo.prepass(ro, step=0)
    info = {"eval/safe_data": float(pp["bTa_is_safe"].float().mean())}
    Vl, Vh, Qh, Ql, A = pp["bTp1_Vl"], pp["bTp1ah_Vh"], pp["bTah_Qh"], pp["bT_Ql"], pp["bTa_A"]
    assert Vl.shape == (b, T + 1) and Vh.shape == (b, T + 1, n, 2) and Qh.shape == (b, T, n, 2) and A.shape == (b, T, n)
    for t in (Vl, Vh, Qh, Ql, A, pp["bTp1ah_Vh_det"], pp["bTah_Qh_det"]):
        assert bool(torch.isfinite(t).all())
    assert 0.0 <= info["eval/safe_data"] <= 1.0
    # Qh is a lambda-mix of running maxima over the costs (of any component: the recursion takes max_h h_t) and
    # the bootstrap values: bounded by their extremes per (environment, agent)
    hi = torch.maximum(ro.costs.amax(dim=(1, 3)), Vh.amax(dim=(1, 3)))            # (b, n)
    lo = torch.minimum(ro.costs.amin(dim=(1, 3)), Vh.amin(dim=(1, 3)))
    assert bool((Qh <= hi[:, None, :, None] + 1e-4).all()) and bool((Qh >= lo[:, None, :, None] - 1e-4).all())

    rng = np.random.default_rng(3)
    envs = rng.integers(0, b, 4)
    # GAE and advantage merge: full trajectories of a few environments through the oracle
    for e in envs:
        rQh, rQl = algo_np.compute_dec_ocp_gae(ro.costs[e].cpu().numpy(), -ro.rewards[e].cpu().numpy(),
                                               Vh[e].cpu().numpy(), Vl[e].cpu().numpy(), 0.99, 0.95)
        np.testing.assert_allclose(Qh[e].cpu().numpy(), rQh, rtol=1e-5, atol=5e-6)
        np.testing.assert_allclose(Ql[e].cpu().numpy(), rQl, rtol=1e-5, atol=5e-6)
    sel = torch.as_tensor(envs, device=Vl.device)
    rA, rd, _, rsafe = algo_np.cbf_advantage(Ql[sel].cpu().numpy(), Vl[sel].cpu().numpy(), Vh[sel].cpu().numpy(),
                                             env.dt, algo.alpha, algo.cbf_eps,
                                             algo.cbf_schedule_fn(0) if algo.cbf_schedule else algo.cbf_weight)
    near = (np.abs(rd) < 1e-4).any(-1)                       # is_safe flips on a sign: skip entries at the threshold
    np.testing.assert_allclose(A[sel].cpu().numpy()[~near], rA[~near], rtol=1e-4, atol=1e-4)
    # Vh on sampled (environment, slot) pairs; Vl by replaying the recurrent scan of two environments
    se = torch.as_tensor(rng.integers(0, b, 48), device=Vl.device)
    st = torch.as_tensor(rng.integers(0, T, 48), device=Vl.device)
    g_t = {k: getattr(ro.graph, k)[se, st].cpu().numpy() for k in GRAPH_FIELDS}
    rnn_t = ro.rnn_states.reshape(b, T, n, 64)[se, st].cpu().numpy()
    rVh = nn_np.vh_forward(algo.params["Vh"], g_t, rnn_t, n)
    np.testing.assert_allclose(Vh[se, st].cpu().numpy(), rVh, rtol=1e-5, atol=1e-5)
    two = torch.as_tensor(envs[:2], device=Vl.device)
    h = np.zeros((2, 64), F)
    for t in range(T + 1):
        src, tt = (ro.graph, t) if t < T else (ro.next_graph, T - 1)
        gt = {k: getattr(src, k)[two, tt].cpu().numpy() for k in GRAPH_FIELDS}
        v, h = nn_np.vl_forward(algo.params["Vl"], gt, h, n)
        np.testing.assert_allclose(Vl[two, t].cpu().numpy(), v, rtol=5e-5, atol=5e-5, err_msg=f"Vl t={t}")
