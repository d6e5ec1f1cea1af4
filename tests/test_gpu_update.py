"""DGPPO.update end to end on the GPU: the torch forward of algo/update.py against the CUDA kernels' outputs
stored in the rollout, and one full update step (parameters move, losses finite, kernels pick the new weights up)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _setup(n_env=32, T=32, batch=1024):
    from dgppo_b200.algo import make_algo
    from dgppo_b200.env import make_env
    env = make_env("LidarSpread", num_agents=3, num_obs=3, max_step=T)
    algo = make_algo("dgppo", env=env, node_dim=env.node_dim, edge_dim=env.edge_dim, state_dim=env.state_dim,
                     action_dim=env.action_dim, n_agents=env.num_agents, batch_size=batch, rnn_step=16, seed=3)
    return env, algo


def test_torch_forward_matches_kernels_on_rollout():
    """Chunk 0 of the BPTT starts from the zero carry the rollout started from, so for t < rnn_step the
    recomputed log_pi must equal the kernels' (ratio == 1 before the first update), and Vh likewise."""
    from dgppo_b200.algo import update as U
    env, algo = _setup()
    ro = algo.collect(algo.params, np.arange(32) + 11)
    b, T = ro.rewards.shape
    n = env.num_agents
    d = env.graph_dims()
    gi = U.GraphIndex(n, d.n_ag, d.n_ao, d.n_nodes, algo.device)
    arrays = algo._record_arrays(ro)
    ix = torch.arange(b, device=algo.device)
    g = U.chunk_graphs(arrays, ix, T, gi, torch.float32)
    tree = U.to_torch_tree(algo.params["policy"], algo.device, requires_grad=False)
    emb = U.gnn(tree["params"]["PolicyNet_0"]["GraphTransformerGNN_0"], g, gi, 2).reshape(b, T, n, -1)
    h = torch.zeros((b, n, 64), device=algo.device)
    for t in range(16):
        mean, std, h = U.policy_step(tree, emb[:, t], h)
        lp = U.tanh_normal_log_prob(ro.actions[:, t], mean, std)
        torch.testing.assert_close(lp, ro.log_pis[:, t], rtol=2e-4, atol=2e-4)
        if t + 1 < T:       # the carry the kernels stored for the next step
            torch.testing.assert_close(h, ro.rnn_states[:, t + 1].reshape(b, n, 64), rtol=1e-4, atol=1e-5)
    # Vh over the record vs the torch value net
    Vh = algo._value_record("Vh", ro, None)
    tv = U.to_torch_tree(algo.params["Vh"], algo.device, requires_grad=False)
    embv = U.gnn(tv["params"]["GraphTransformerGNN_0"], g, gi, 1)
    out, _ = U.value_step(tv, embv, ro.rnn_states.reshape(b * T, n, 64))
    torch.testing.assert_close(out.reshape(b, T, n, -1), Vh[:, :T], rtol=1e-4, atol=1e-5)


def test_update_step_moves_parameters_and_kernels_follow():
    env, algo = _setup()
    keys = np.arange(32) + 5
    ro = algo.collect(algo.params, keys)
    before = {k: np.concatenate([np.ravel(v) for _, v in __import__("dgppo_b200.algo.update", fromlist=["x"]).tree_leaves(t)])
              for k, t in algo.params.items()}
    a0 = ro.actions.clone()
    info = algo.update(ro, 0)
    for k in ("policy/loss", "Vl/loss", "Vh/loss_Vh", "policy/grad_norm", "Vl/grad_norm", "Vh/grad_Vh_norm",
              "policy/entropy", "policy/clip_frac", "eval/safe_data"):
        assert k in info and np.isfinite(info[k]), (k, info.get(k))
    assert info["policy/has_nan"] == 0.0 and info["policy/clip_frac"] <= 1.0
    from dgppo_b200.algo import update as U
    for k, t in algo.params.items():
        after = np.concatenate([np.ravel(v) for _, v in U.tree_leaves(t)])
        delta = np.abs(after - before[k]).max()
        assert 0 < delta < 0.05, (k, delta)                    # Adam: |step| <= lr per parameter and update
    # the kernels run with the updated weights: same keys and noise, different actions
    ro2 = algo.collect(algo.params, keys)
    assert not torch.equal(ro2.actions, a0)
    # a second update reuses the optimiser state (count advances)
    algo.update(ro2, 1)
    assert float(algo._train["policy"].count) == 2 * algo.epoch_ppo * max(1, 32 // (1024 // 32))
    assert float(algo._train["policy"].notfinite_count) == 0


def test_update_rejects_inconsistent_sizes():
    env, algo = _setup(batch=1024 * 64)
    ro = algo.collect(algo.params, np.arange(32))
    with pytest.raises(ValueError):
        algo.update(ro, 0)


@pytest.mark.parametrize("env_id,n,obs", [("MPEConnectSpread", 3, 1), ("LidarLine", 4, 3), ("MPEFormation", 4, 2)])
def test_collect_and_update_on_widened_families(env_id, n, obs):
    """Rollout + update through a three-cost env (Vh with 3 outputs, GAE with nh = 3) and through the landmark
    families (goal nodes != agents): shapes, finiteness, the record's graph equals a fresh get_graph."""
    from dgppo_b200.algo import make_algo
    from dgppo_b200.env import make_env
    T = 32
    env = make_env(env_id, num_agents=n, num_obs=obs, max_step=T)
    algo = make_algo("dgppo", env=env, node_dim=env.node_dim, edge_dim=env.edge_dim, state_dim=env.state_dim,
                     action_dim=env.action_dim, n_agents=n, batch_size=512, rnn_step=16, seed=1)
    ro = algo.collect(algo.params, np.arange(16) + 3)
    assert ro.costs.shape == (16, T, n, env.n_cost)
    d = env.graph_dims()
    assert ro.graph.nodes.shape == (16, T, d.n_nodes, d.node_dim)
    assert ro.graph.env_states.goal.shape[-2] == env.num_goals
    # slot t of the record == get_graph of the state stored there
    es0 = ro.graph.env_states                      # agent / goal are per slot, the obstacles static per env
    es = type(es0)(es0.agent[:, 7], es0.goal[:, 7], es0[2])
    if env_id.startswith("Lidar"):
        hits = ro.graph.states[:, 7, n + env.num_goals:n + env.num_goals + 8 * n, :2].reshape(16, n, 8, 2)
        g7 = env.get_graph(es, hits)
    else:
        g7 = env.get_graph(es)
    for k in ("nodes", "edges", "receivers", "senders"):
        assert torch.equal(getattr(g7, k), getattr(ro.graph, k)[:, 7]), k
    info = algo.update(ro, 0)
    assert all(np.isfinite(v) for v in info.values()), info
    assert algo.last_prepass["bTp1ah_Vh"].shape == (16, T + 1, n, env.n_cost)


def test_graphed_update_equals_eager(monkeypatch):
    """The CUDA-graph replay of the minibatch step takes the same step as the eager one (same data, same draws)."""
    keys = np.arange(32) + 5
    outs = []
    for flag in ("1", "0"):
        monkeypatch.setenv("DGPPO_UPDATE_GRAPH", flag)
        env, algo = _setup()
        ro = algo.collect(algo.params, keys)
        info = algo.update(ro, 0)
        info2 = algo.update(algo.collect(algo.params, keys), 1)
        outs.append((info, info2, {k: v.flat.detach().clone() for k, v in algo._train.items()}))
    (a1, a2, pa), (b1, b2, pb) = outs
    for k in a1:
        np.testing.assert_allclose(a1[k], b1[k], rtol=2e-4, atol=1e-6, err_msg=k)
    for k in pa:      # parameters after two updates (fp32 reductions may be ordered differently under capture)
        torch.testing.assert_close(pa[k], pb[k], rtol=1e-3, atol=2e-5)
