"""The oracle against the REFERENCE: fixtures produced by executing the
reference's own env / GAE source under oracle/jaxshim.py
(tools/gen_golden_from_reference.py).  CPU only."""
import numpy as np
import pytest

from oracle import algo_np, env_np
from tests import golden_util as G
from tests.util import assert_bits_equal


@pytest.mark.parametrize("name", list(G.CASES))
def test_env_step_matches_reference(name):
    cfg, d, obstacles = G.load(name)
    T = d["action"].shape[1]
    rays = env_np.ray_table(cfg.n_rays, cfg.comm_radius) if cfg.is_lidar else None
    for t in range(T):
        g = G.graph_at(d, t)
        g_next, r, c, done = env_np.env_step(cfg, g, d["action"][:, t], obstacles, rays)
        ref = G.graph_at(d, t + 1)
        for k in ("receivers", "senders", "node_type", "n_node", "n_edge", "nodes", "edges", "states"):
            assert_bits_equal(g_next[k], ref[k], f"{name} graph.{k} t={t + 1}")
        assert_bits_equal(c, d["cost"][:, t], f"{name} cost t={t}")
        # the reward's means are summed in an unspecified order by NumPy/XLA
        np.testing.assert_allclose(r, d["reward"][:, t], rtol=2e-6, atol=1e-9)
        assert not done.any()


@pytest.mark.parametrize("name", list(G.CASES))
def test_reset_graph_matches_reference(name):
    """get_graph (+ LiDAR) on the reference's reset states reproduces its graph."""
    cfg, d, obstacles = G.load(name)
    g0 = G.graph_at(d, 0)
    agent, goal, obs_nodes = env_np.graph_slices(cfg, g0)
    mpe_obs = d["mpe_obs"] if "mpe_obs" in d.files else None
    g = env_np.reset_graph(cfg, agent, goal, obstacles, mpe_obs)
    for k in G.GRAPH_FIELDS:
        assert_bits_equal(g[k], g0[k], f"{name} reset graph.{k}")


def test_obstacle_points_match_reference():
    cfg, d, obstacles = G.load("LidarSpread_n8_obs8")
    rc = env_np.rect_create(d["obs_center"], d["obs_width"], d["obs_height"], d["obs_theta"])
    # Rectangle.create multiplies rot @ bbox with jnp.dot (obstacle.py:53): the BLAS / XLA dot may
    # contract to FMA, the oracle rounds every product -> agreement to 1 ulp.  The corner points are
    # reset-time DATA for the hot path (carried in the obstacle record), not recomputed per step.
    np.testing.assert_allclose(rc["points"], d["obs_points"], rtol=2.5e-7, atol=1e-7)


def test_gae_matches_reference():
    d = np.load(G.GOLDEN_DIR + "/ref_gae.npz")
    for i in range(3):
        Qh, Ql = algo_np.compute_dec_ocp_gae(d[f"c{i}_hs"], d[f"c{i}_l"], d[f"c{i}_Vh"], d[f"c{i}_Vl"], 0.99, 0.95)
        np.testing.assert_allclose(Qh, d[f"c{i}_Qh"], rtol=1e-5, atol=2e-6)
        np.testing.assert_allclose(Ql, d[f"c{i}_Ql"], rtol=1e-5, atol=2e-6)


@pytest.mark.parametrize("name", list(G.NN_CASES))
def test_networks_match_reference_modules(name):
    """Oracle vs the reference's own nn/gnn.py, nn/mlp.py, nn/rnn.py, policy.py, value.py and
    distribution.py executed under oracle/flaxshim.py: pins the COMPOSITION (which node feeds
    q/k/v, head order, module naming, carry wiring, thresholded log-prob)."""
    from oracle import nn_np
    cfg, d, graph, trees = G.load_nn(name)
    n = cfg.n
    # flax auto-naming as the reference builds it: the GRU sits under GRUCell_1
    assert list(trees["policy"]["params"]["PolicyNet_0"]["RNN_0"]) == ["GRUCell_1"]
    assert set(trees["policy"]["params"]) == {"PolicyNet_0", "ScaleHid", "OutputDenseMean", "OutputDenseStdTrans"}
    assert set(trees["vh"]["params"]) == {"GraphTransformerGNN_0", "ValueGNNHead", "RNN_0", "Dense_0"}
    a, _, h, _ = nn_np.policy_forward(trees["policy"], graph, d["rnn"], n, None, 2)
    np.testing.assert_allclose(a, d["act_mode"], rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(h, d["rnn_out"], rtol=1e-5, atol=2e-6)
    a, lp, _, _ = nn_np.policy_forward(trees["policy"], graph, d["rnn"], n, d["eps"], 2)
    np.testing.assert_allclose(a, d["act_sample"], rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(lp, d["log_pi"], rtol=1e-5, atol=5e-6)
    vh = nn_np.vh_forward(trees["vh"], graph, d["rnn"], n, 1)
    np.testing.assert_allclose(vh, d["vh"], rtol=1e-5, atol=2e-6)
    vl, vlh = nn_np.vl_forward(trees["vl"], graph, d["vl_rnn"], n, 2)
    np.testing.assert_allclose(vl, d["vl"], rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(vlh, d["vl_rnn_out"], rtol=1e-5, atol=2e-6)
