"""GPU parity of K1 (step), K2 (LiDAR), K3 (graph) against the oracle, through
the C ABI.  Integer / mask / state results must be BIT-EXACT; the bicycle
dynamics (atan2 / sin / cos from different libms) use rtol 1e-5."""
import numpy as np
import pytest

from oracle import env_np
from tests import util
from tests.util import CONFIGS, assert_bits_equal

pytestmark = pytest.mark.gpu
F = np.float32


def _states(cfg, b, seed, threshold):
    fn = util.threshold_states if threshold else env_np.synthetic_states
    return fn(cfg, b, seed)


@pytest.mark.parametrize("name", list(CONFIGS))
@pytest.mark.parametrize("threshold", [False, True])
def test_lidar_bit_exact(name, threshold):
    cfg = CONFIGS[name]
    if not cfg.is_lidar or cfg.n_obs == 0:
        pytest.skip("no LiDAR in this config")
    b = 8 if cfg.n >= 64 else 64
    agent, goal, obstacles, _ = _states(cfg, b, 1, threshold)
    rays = env_np.ray_table(cfg.n_rays, cfg.comm_radius)
    ref = env_np.lidar_hits(cfg, agent[..., :2], obstacles, rays)
    got = util.k_lidar(cfg, agent, obstacles, rays)
    assert_bits_equal(got, ref, f"lidar hits {name}")


def test_lidar_parallel_ray_nan():
    """A ray exactly parallel to an obstacle edge gives det == 0 -> alpha NaN for
    that obstacle (env/obstacle.py:89-104: sign(0) = 0); NaN must propagate through
    the min and sort last, exactly as in the reference arithmetic."""
    cfg = env_np.EnvCfg(env_np.LIDAR_SPREAD, n=2, n_obs=1)
    obstacles = env_np.rect_create(np.array([[[1.0, 0.75]]], F), np.array([[0.2]], F), np.array([[0.2]], F),
                                   np.array([[0.0]], F))
    agent = np.zeros((1, 2, 4), F)
    agent[0, 0, :2] = [0.5, 0.75]     # ray 16 (theta = 0) is parallel to the horizontal edges
    agent[0, 1, :2] = [1.0, 0.3]
    rays = env_np.ray_table(32, 0.5)
    rays[16] = [0.5, 0.0]             # the table is data: make beam 16 exactly horizontal
    ref = env_np.lidar_hits(cfg, agent[..., :2], obstacles, rays)
    got = util.k_lidar(cfg, agent, obstacles, rays)
    al = env_np.lidar_alphas(agent[..., :2], obstacles, rays)
    assert np.isnan(al[0, 0, 16]), "test premise: beam 16 of agent 0 has a NaN alpha"
    assert not np.isnan(ref).any()    # NaN sorts last, so the head-on beam drops out of the top-k
    assert_bits_equal(got, ref, "lidar hits with NaN")


@pytest.mark.parametrize("name", list(CONFIGS))
@pytest.mark.parametrize("threshold", [False, True])
def test_graph_bit_exact(name, threshold):
    cfg = CONFIGS[name]
    b = 8 if cfg.n >= 64 else 64
    agent, goal, obstacles, mpe_obs = _states(cfg, b, 2, threshold)
    if cfg.is_lidar:
        obs_nodes = env_np.lidar_hits(cfg, agent[..., :2], obstacles) if cfg.n_obs > 0 else None
    else:
        obs_nodes = mpe_obs
    ref = env_np.get_graph(cfg, agent, goal, obs_nodes)
    got = util.k_graph(cfg, agent, goal, obs_nodes)
    for k in ("receivers", "senders", "node_type", "n_node", "n_edge", "nodes", "edges", "states"):
        assert_bits_equal(got[k], ref[k], f"graph.{k} {name}")
    # structural invariants
    N = cfg.n_nodes
    assert got["receivers"].shape[1] == cfg.n * cfg.n + cfg.n * cfg.n_ag + cfg.n * cfg.n_ao
    assert ((got["receivers"] < cfg.n) | (got["receivers"] == N - 1)).all()
    assert ((got["receivers"] == N - 1) == (got["senders"] == N - 1)).all()
    if threshold and cfg.n >= 2:
        assert (got["receivers"] == N - 1).any() and (got["receivers"] < cfg.n).any()


@pytest.mark.parametrize("name", list(CONFIGS))
def test_step_cost_reward(name):
    cfg = CONFIGS[name]
    b = 8 if cfg.n >= 64 else 64
    agent, goal, obstacles, mpe_obs = util.threshold_states(cfg, b, 3)
    rng = np.random.default_rng(7)
    action = rng.uniform(-1.5, 1.5, (b, cfg.n, 2)).astype(F)       # exercises clip_action
    if cfg.is_lidar:
        obs_nodes = env_np.lidar_hits(cfg, agent[..., :2], obstacles) if cfg.n_obs > 0 else None
    else:
        obs_nodes = mpe_obs
    a = env_np.clip_action(action)
    ref_next = env_np.agent_step_euler(cfg, agent, a)
    ref_rew = env_np.get_reward(cfg, agent, goal, a)
    ref_cost = env_np.get_cost(cfg, agent, obs_nodes)
    nxt, rew, cost = util.k_env_step(cfg, agent, goal, obs_nodes, action)
    assert_bits_equal(cost, ref_cost, f"cost {name}")
    assert_bits_equal(rew, ref_rew, f"reward {name}")
    if cfg.is_bicycle:
        np.testing.assert_allclose(nxt, ref_next, rtol=1e-5, atol=1e-6)   # libm atan2/sin/cos
    else:
        assert_bits_equal(nxt, ref_next, f"next state {name}")


def test_empty_batch_and_bad_args():
    import ctypes as C
    from dgppo_b200 import _lib
    cfg = util.c_cfg(CONFIGS["C3"])
    lib = _lib.lib()
    assert lib.dgppo_env_step(None, C.byref(cfg), None, None, None, None, None, None, None, 1, 0) == 0
    assert lib.dgppo_env_step(None, C.byref(cfg), None, None, None, None, None, None, None, 1, 4) == -1
    bad = util.c_cfg(CONFIGS["C3"])
    bad.kind = 10
    assert lib.dgppo_lidar(None, C.byref(bad), None, None, None, None, 4) == -2


def test_env_step_multi_step_trajectory():
    """Full env.step chain (step -> lidar -> graph) over 20 steps, bit-exact at
    every step when both sides are fed the same actions."""
    cfg = CONFIGS["C3"]
    b, T = 32, 20
    agent, goal, obstacles, _ = env_np.synthetic_states(cfg, b, 11)
    rays = env_np.ray_table(cfg.n_rays, cfg.comm_radius)
    g = env_np.reset_graph(cfg, agent, goal, obstacles, None, rays)
    rng = np.random.default_rng(5)
    cur_agent, cur_hits = agent, env_np.graph_slices(cfg, g)[2]
    for t in range(T):
        action = rng.uniform(-1, 1, (b, cfg.n, 2)).astype(F)
        g, r, c, _ = env_np.env_step(cfg, g, action, obstacles, rays)
        nxt, rew, cost = util.k_env_step(cfg, cur_agent, goal, cur_hits, action)
        hits = util.k_lidar(cfg, nxt, obstacles, rays)
        kg = util.k_graph(cfg, nxt, goal, hits)
        assert_bits_equal(rew, r, f"reward t={t}")
        assert_bits_equal(cost, c, f"cost t={t}")
        for k in ("nodes", "edges", "states", "receivers", "senders"):
            assert_bits_equal(kg[k], g[k], f"{k} t={t}")
        cur_agent, cur_hits = nxt, hits
