"""The kernels' update pre-pass and the GPU update step against the REFERENCE'S OWN `DGPPO.update`
(dgppo/algo/dgppo.py:136-321) executed under the NumPy stand-ins (tests/golden/ref_update_*.npz, made by
tools/gen_golden_update_from_reference.py; the CPU side of the same fixture: tests/test_update_reference.py).

The reference's rollouts (stochastic + deterministic) and its three parameter pytrees are loaded into
`DGPPO`; `scan_Vl`, the Vh record, both GAE passes and the CBF advantage merge run on the CUDA kernels and
must reproduce the reference's intermediates; the fp32 CUDA losses of algo/update.py must equal the values
of the reference's `get_loss_` closures."""
import numpy as np
import pytest
import torch

from tests.test_update_reference import load

pytestmark = pytest.mark.gpu


def _algo_and_rollouts():
    from dgppo_b200.algo import make_algo
    from dgppo_b200.env import make_env
    from dgppo_b200.trainer.data import Rollout
    from dgppo_b200.utils.graph import GraphsTuple
    d, trees, hp, (n, n_obs, b, T, rnn_step) = load()
    step, train_steps = int(d["meta"][5]), int(d["meta"][6])
    env = make_env("LidarSpread", num_agents=n, num_obs=n_obs, max_step=T)
    algo = make_algo("dgppo", env=env, node_dim=env.node_dim, edge_dim=env.edge_dim, state_dim=env.state_dim,
                     action_dim=env.action_dim, n_agents=n, batch_size=b * T, rnn_step=rnn_step, seed=0,
                     train_steps=train_steps, compact_record=False)
    for name in ("policy", "Vl", "Vh"):
        algo.set_params(name, trees[name])
    dev = algo.device

    def rollout(tag):
        t = lambda k, dt=torch.float32: torch.tensor(d[f"{tag}:{k}"], dtype=dt, device=dev)      # noqa: E731
        full = {k: t(k, torch.int32 if k in ("n_node", "n_edge", "receivers", "senders", "node_type") else torch.float32)
                for k in ("n_node", "n_edge", "nodes", "edges", "states", "receivers", "senders", "node_type")}

        def view(lo, hi):
            return GraphsTuple(*[full[k][:, lo:hi].contiguous() for k in
                                 ("n_node", "n_edge", "nodes", "edges", "states", "receivers", "senders", "node_type")],
                               None)
        return Rollout(graph=view(0, T), actions=t("actions"),
                       rnn_states=t("rnn_states").reshape(b, T, 1, n, 1, 64), rewards=t("rewards"), costs=t("costs"),
                       dones=torch.zeros((b, T), dtype=torch.bool, device=dev),
                       log_pis=t("log_pis") if tag == "ro" else None, next_graph=view(1, T + 1))
    return d, hp, (n, n_obs, b, T, rnn_step, step), algo, rollout("ro"), rollout("det")


def test_kernel_prepass_reproduces_the_reference_update():
    d, hp, (n, n_obs, b, T, rnn_step, step), algo, ro, det = _algo_and_rollouts()
    Vl, _ = algo.scan_Vl(ro)
    np.testing.assert_allclose(Vl.cpu().numpy(), d["Vl"], rtol=2e-5, atol=5e-6)
    Vh = algo._value_record("Vh", ro, None)
    np.testing.assert_allclose(Vh.cpu().numpy(), d["Vh"], rtol=2e-5, atol=5e-6)
    Vh_det = algo._value_record("Vh", det, None)
    np.testing.assert_allclose(Vh_det.cpu().numpy(), d["Vh_det"], rtol=2e-5, atol=5e-6)
    Qh, Ql = algo.gae(ro.costs, -ro.rewards, Vh, Vl)
    np.testing.assert_allclose(Qh.cpu().numpy(), d["Qh"], rtol=2e-5, atol=5e-6)
    np.testing.assert_allclose(Ql.cpu().numpy(), d["Ql"], rtol=2e-5, atol=5e-6)
    Qh_det, _ = algo.gae(det.costs, -det.rewards, Vh_det, Vl)
    np.testing.assert_allclose(Qh_det.cpu().numpy(), d["Qh_det"], rtol=2e-5, atol=5e-6)
    assert algo.cbf_schedule_fn(step) == hp["cbf_weight"]
    A, deriv, _, safe = algo.cbf_advantage(Ql, Vl, Vh, step)
    near = (deriv.abs() < 1e-3).any(-1).cpu().numpy()          # residual within rounding of 0: `is_safe` may flip
    # the residual divides a Vh difference by dt = 0.03: fp32 differences of the 1e-5-accurate Vh, x 33
    np.testing.assert_allclose(A.cpu().numpy()[~near], d["A"][~near], rtol=5e-4, atol=5e-4)
    np.testing.assert_allclose(float(safe.float().mean()), float(d["safe_data"]), atol=near.mean() + 1e-7)


def test_cuda_losses_equal_the_reference_closures():
    from dgppo_b200.algo import update as U
    d, hp, (n, n_obs, b, T, rnn_step, step), algo, ro, det = _algo_and_rollouts()
    dev = algo.device
    dims = algo._env.graph_dims()
    gi = U.GraphIndex(n, dims.n_ag, dims.n_ao, dims.n_nodes, dev)
    ix = torch.arange(b, device=dev)
    g = U.chunk_graphs(algo._record_arrays(ro), ix, T, gi, torch.float32)
    gd = U.chunk_graphs(algo._record_arrays(det), ix, T, gi, torch.float32)
    t = lambda a: torch.tensor(np.asarray(a), dtype=torch.float32, device=dev)      # noqa: E731
    tp = {k: U.to_torch_tree(algo.params[k], dev, requires_grad=False) for k in ("policy", "Vl", "Vh")}
    eps = t(d["entropy_eps"]).expand(b, T, n, 2)
    lv = U.loss_Vl(tp["Vl"], g, t(d["Ql"]), gi, 2, rnn_step)
    lh = U.loss_Vh(tp["Vh"], gd, det.rnn_states.reshape(b, T, n, 64), t(d["Qh_det"]), gi, 1)
    lp, info = U.loss_policy(tp["policy"], g, ro.actions, ro.log_pis, t(d["A"]), eps, gi, 2, rnn_step,
                             hp["clip_eps"], hp["coef_ent"])
    np.testing.assert_allclose(float(lv), float(d["loss:Vl"]), rtol=1e-4)
    np.testing.assert_allclose(float(lh), float(d["loss:Vh"]), rtol=1e-4)
    np.testing.assert_allclose(float(lp), float(d["loss:policy"]), rtol=1e-4)
    np.testing.assert_allclose(float(info["policy/entropy"]), float(d["aux:policy/entropy"]), rtol=1e-4)


def test_deterministic_rollout_replays_the_reference_trajectory():
    """`DGPPO.det_rollout_fn` (the captured CUDA rollout) from the reference's initial state reproduces the
    reference's own test_rollout (trainer/utils.py:60-86 through its policy, env.step, LiDAR and get_graph):
    actions, carries (post-step convention), rewards, costs and every graph of the T + 1 slots."""
    from dgppo_b200.env.envs import LidarEnvState, Rectangle, rect_record
    d, hp, (n, n_obs, b, T, rnn_step, step), algo, ro, det_ref = _algo_and_rollouts()
    env, dev = algo._env, algo.device
    rec = rect_record(d["det:obs_center"], d["det:obs_width"], d["det:obs_height"], d["det:obs_theta"])
    rec[..., 8:16] = d["det:obs_points"].reshape(b, n_obs, 8)          # the reference's corners (its jnp.dot: 1 ulp)
    st0 = d["det:states"][:, 0]
    t = lambda a: torch.tensor(np.ascontiguousarray(a), dtype=torch.float32, device=dev)      # noqa: E731
    es = LidarEnvState(t(st0[:, :n]), t(st0[:, n:2 * n]), Rectangle.from_record(rec, dev))
    g0 = env.get_graph(es, t(st0[:, 2 * n:2 * n + 8 * n, :2]))
    for k in ("nodes", "edges", "states", "receivers", "senders"):
        assert np.array_equal(getattr(g0, k).cpu().numpy(), d[f"det:{k}"][:, 0]), k
    out = algo.det_rollout_fn(algo.params, None, graph0=g0)
    np.testing.assert_allclose(out.actions.cpu().numpy(), d["det:actions"], rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(out.rnn_states.reshape(b, T, n, 64).cpu().numpy(), d["det:rnn_states"], rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(out.rewards.cpu().numpy(), d["det:rewards"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(out.costs.cpu().numpy(), d["det:costs"], rtol=1e-4, atol=2e-5)
    nodes = torch.cat([out.graph.nodes, out.next_graph.nodes[:, -1:]], dim=1).cpu().numpy()
    finite = np.abs(d["det:nodes"]) < 1e3                          # LiDAR misses sit ~5e5 away: compare relatively
    np.testing.assert_allclose(nodes[finite], d["det:nodes"][finite], rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(nodes[~finite], d["det:nodes"][~finite], rtol=1e-4)
    for k in ("receivers", "senders"):
        got = torch.cat([getattr(out.graph, k), getattr(out.next_graph, k)[:, -1:]], dim=1).cpu().numpy()
        assert (got != d[f"det:{k}"]).mean() < 0.01, k            # a mask may flip where a distance sits on its threshold


def test_stochastic_rollout_replays_the_reference_trajectory():
    """`DGPPO.collect` against the reference's own `collect` (informarl.py:254-256 -> trainer/utils.py:22-57): the
    N(0,1) draw behind every sampled action is recovered from the reference's actions with the oracle (which
    reproduces that trajectory: tests/test_update_reference.py), then the CUDA rollout runs from the reference's
    initial state with those draws: actions, log_pis, pre-step carries, rewards, costs."""
    from dgppo_b200.env.envs import LidarEnvState, Rectangle, rect_record
    from oracle import env_np, nn_np
    from tests.test_update_reference import _obstacles
    d, hp, (n, n_obs, b, T, rnn_step, step), algo, ro_ref, det_ref = _algo_and_rollouts()
    trees = {k: algo.params[k] for k in ("policy",)}
    cfg = env_np.EnvCfg(env_np.LIDAR_SPREAD, n=n, n_obs=n_obs, max_step=T)
    fields = ("n_node", "n_edge", "nodes", "edges", "states", "receivers", "senders", "node_type")
    g = {k: d[f"ro:{k}"][:, 0] for k in fields}
    obst, h = _obstacles(d, "ro"), np.zeros((b, n, 64), np.float32)
    eps = np.zeros((b, T, n, 2), np.float32)
    for t in range(T):
        _, _, _, (mean, std) = nn_np.policy_forward(trees["policy"], g, h, n, eps=None)
        eps[:, t] = ((np.arctanh(d["ro:actions"][:, t].astype(np.float64)) - mean) / std).astype(np.float32)
        _, _, h, _ = nn_np.policy_forward(trees["policy"], g, h, n, eps=eps[:, t])
        g, _, _, _ = env_np.env_step(cfg, g, d["ro:actions"][:, t], obst)
    env, dev = algo._env, algo.device
    rec = rect_record(d["ro:obs_center"], d["ro:obs_width"], d["ro:obs_height"], d["ro:obs_theta"])
    rec[..., 8:16] = d["ro:obs_points"].reshape(b, n_obs, 8)
    st0 = d["ro:states"][:, 0]
    t_ = lambda a: torch.tensor(np.ascontiguousarray(a), dtype=torch.float32, device=dev)      # noqa: E731
    es = LidarEnvState(t_(st0[:, :n]), t_(st0[:, n:2 * n]), Rectangle.from_record(rec, dev))
    g0 = env.get_graph(es, t_(st0[:, 2 * n:2 * n + 8 * n, :2]))
    out = algo.collect(algo.params, None, eps=t_(eps), graph0=g0)
    np.testing.assert_allclose(out.actions.cpu().numpy(), d["ro:actions"], rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(out.log_pis.cpu().numpy(), d["ro:log_pis"], rtol=2e-4, atol=2e-4)
    np.testing.assert_allclose(out.rnn_states.reshape(b, T, n, 64).cpu().numpy(), d["ro:rnn_states"], rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(out.rewards.cpu().numpy(), d["ro:rewards"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(out.costs.cpu().numpy(), d["ro:costs"], rtol=1e-4, atol=2e-5)
