"""CPU-only invariants of the oracle: structure of the
padded graph, LiDAR geometry, GAE closed forms, tanh-Normal log-prob."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st
from scipy import stats

from oracle import algo_np, env_np, nn_np
from tests.util import CONFIGS, threshold_states

F = np.float32


@pytest.mark.parametrize("name", [k for k in CONFIGS if k != "C5"])
def test_graph_structure(name):
    cfg = CONFIGS[name]
    agent, goal, obst, mo = threshold_states(cfg, 16, 0)
    g = env_np.reset_graph(cfg, agent, goal, obst, mo)
    n, N, E = cfg.n, cfg.n_nodes, cfg.n_edges
    assert E == n * n + n * cfg.n_ag + n * cfg.n_ao
    recv, send = g["receivers"], g["senders"]
    assert ((recv < n) | (recv == N - 1)).all() and ((recv == N - 1) == (send == N - 1)).all()
    assert (recv[:, ::n + 1][:, :n] == N - 1).all()                # no self edges (diagonal of the a-a block)
    assert (recv[:, n * n:n * n + n * cfg.n_ag] < n).all()           # agent-goal edges are always on
    assert (g["nodes"][:, N - 1] == 0).all() and (g["states"][:, N - 1] == -1).all()
    assert (g["node_type"][:, :n] == 0).all() and (g["node_type"][:, N - 1] == -1).all()
    sd = cfg.state_dim
    assert (g["nodes"][:, :n, sd + 2] == 1).all() and (g["nodes"][:, n:2 * n, sd + 1] == 1).all()
    # a-a edge features are antisymmetric: f(i) - f(j) = -(f(j) - f(i))
    aa = g["edges"][:, :n * n].reshape(-1, n, n, 4)
    np.testing.assert_array_equal(aa, -aa.transpose(0, 2, 1, 3))


def test_lidar_geometry():
    cfg = CONFIGS["C3"]
    agent, goal, obst, _ = env_np.synthetic_states(cfg, 64, 3)
    rays = env_np.ray_table(cfg.n_rays, cfg.comm_radius)
    al = env_np.lidar_alphas(agent[..., :2], obst, rays)
    hits = env_np.lidar_hits(cfg, agent[..., :2], obst, rays)
    inside = env_np.rect_inside(agent[..., :2], obst, 0.0)
    assert (al[inside] == 0).all() and ((al[~inside] >= 0) | np.isnan(al[~inside])).all()
    d = np.linalg.norm(hits - agent[:, :, None, :2], axis=-1)
    real = d < 1.0                                              # misses land ~5e5 away
    assert (d[real] <= cfg.comm_radius * (1 + 1e-5)).all()
    assert real.any() and (~real).any()
    # the k-th returned hit has the k-th smallest alpha (stable order)
    srt = np.sort(al, axis=-1)[..., :cfg.top_k]
    np.testing.assert_allclose(d[real], (srt * cfg.comm_radius)[real], rtol=2e-5, atol=1e-6)


def test_ray_table_is_unit_circle_scaled():
    t = env_np.ray_table(32, 0.5)
    np.testing.assert_allclose(np.hypot(t[:, 0], t[:, 1]), 0.5, rtol=1e-6)
    assert t[0, 0] == F(-0.5) and abs(t[16, 1]) < 1e-7


@settings(max_examples=25, deadline=None)
@given(T=st.integers(1, 24), a=st.integers(1, 4), nh=st.integers(1, 3), seed=st.integers(0, 10 ** 6),
       gamma=st.floats(0.5, 0.999), lam=st.floats(0.0, 1.0))
def test_gae_closed_forms(T, a, nh, seed, gamma, lam):
    rng = np.random.default_rng(seed)
    hs, l = rng.normal(-0.5, 0.5, (T, a, nh)), rng.uniform(0, 0.1, T)
    Vh, Vl = rng.normal(0, 0.5, (T + 1, a, nh)), rng.normal(0, 0.5, T + 1)
    Qh, Ql = algo_np.compute_dec_ocp_gae(hs, l, Vh, Vl, gamma, lam, np.float64)
    cQh, cQl = algo_np.gae_closed_form(hs, l, Vh, Vl, gamma, lam)
    np.testing.assert_allclose(Qh, cQh, rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(Ql, cQl, rtol=1e-9, atol=1e-12)
    # last step: Qh[T-1] = max(h, (1-gamma) max_h h + gamma Vh_T)
    np.testing.assert_allclose(Qh[T - 1], np.maximum(hs[T - 1], (1 - gamma) * hs[T - 1].max(-1, keepdims=True)
                                                     + gamma * Vh[T]), rtol=1e-9, atol=1e-12)


def test_tanh_normal_log_prob_is_a_density():
    """log_prob integrates to ~1 over (-0.999, 0.999) plus the two tail masses."""
    loc, scale = np.array([[0.3]]), np.array([[0.7]])
    ys = np.linspace(-0.998, 0.998, 200001)
    lp = nn_np.tanh_normal_log_prob(ys[:, None, None], loc, scale, np.float64)[:, 0]
    mass = np.trapezoid(np.exp(lp), ys)
    tails = stats.norm.cdf((-np.arctanh(0.999) - 0.3) / 0.7) + stats.norm.sf((np.arctanh(0.999) - 0.3) / 0.7)
    assert abs(mass + tails - 1.0) < 2e-3
    # thresholded branches equal log(tail mass / 1e-3)
    left = nn_np.tanh_normal_log_prob(np.array([[-1.0]]), loc, scale, np.float64)
    np.testing.assert_allclose(left, np.log(stats.norm.cdf((-np.arctanh(0.999) - 0.3) / 0.7) / 1e-3), rtol=1e-3)


def test_log_ndtr_segments():
    x = np.array([-30., -12., -10.5, -9., -1., 0., 2., 6., 12.])
    np.testing.assert_allclose(nn_np.log_ndtr(x, np.float64), stats.norm.logcdf(x), rtol=2e-4, atol=1e-12)
    np.testing.assert_allclose(nn_np.log_ndtr(x.astype(F), F), stats.norm.logcdf(x), rtol=2e-3, atol=1e-7)


def test_gnn_masked_edges_do_not_reach_agents():
    """Edges with recv = send = pad only feed the pad node (gnn.py:101,114): perturbing
    their features must not change any agent's output - the property the CUDA kernels
    rely on to skip them."""
    from dgppo_b200.algo import params as P
    cfg = CONFIGS["C3"]
    agent, goal, obst, mo = threshold_states(cfg, 6, 4)
    g = env_np.reset_graph(cfg, agent, goal, obst, mo)
    tree = P.init_policy_params(7, 4, 2, 2, seed=0, jitter=0.1, scale_final=1.0)
    rnn = np.zeros((6, cfg.n, 64), F)
    a0, _, h0, _ = nn_np.policy_forward(tree, g, rnn, cfg.n, None)
    g2 = dict(g)
    masked = g["receivers"] == cfg.n_nodes - 1
    assert masked.any()
    g2["edges"] = np.where(masked[..., None], F(123.0), g["edges"]).astype(F)
    a1, _, h1, _ = nn_np.policy_forward(tree, g2, rnn, cfg.n, None)
    np.testing.assert_array_equal(a0, a1)
    np.testing.assert_array_equal(h0, h1)
