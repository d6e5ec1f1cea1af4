"""GPU: K0 dgppo_reset against the oracle restatement of the reset sampler (same counter-based
random stream), and the invariants get_node_goal_rng guarantees."""
import ctypes as C

import numpy as np
import pytest
import torch

from dgppo_b200 import _lib
from oracle import env_np, reset_np
from tests import util

pytestmark = pytest.mark.gpu
F = np.float32

CASES = {
    "LidarSpread": env_np.EnvCfg(env_np.LIDAR_SPREAD, n=8, n_obs=8),
    "LidarTarget": env_np.EnvCfg(env_np.LIDAR_TARGET, n=5, n_obs=2),
    "LidarBicycleTarget": env_np.EnvCfg(env_np.LIDAR_BICYCLE_TARGET, n=4, n_obs=3),
    "MPESpread": env_np.EnvCfg(env_np.MPE_SPREAD, n=8, n_obs=3),
    "MPETarget": env_np.EnvCfg(env_np.MPE_TARGET, n=6, n_obs=3),
    "MPECorridor": env_np.EnvCfg(env_np.MPE_CORRIDOR, n=5, n_obs=2, area=1.0, obs_radius=0.2),
    "LidarSpread_noobs": env_np.EnvCfg(env_np.LIDAR_SPREAD, n=4, n_obs=0),
    "crowded": env_np.EnvCfg(env_np.LIDAR_SPREAD, n=24, n_obs=12),
}


def _run(cfg, keys):
    b = len(keys)
    k = torch.from_numpy(np.asarray(keys, np.uint64).astype(np.int64)).cuda()
    agent = torch.empty((b, cfg.n, cfg.state_dim), device="cuda")
    goal = torch.empty_like(agent)
    w = _lib.OBS_STRIDE if cfg.is_lidar else 4
    obst = torch.empty((b, max(cfg.n_obs, 1), w), device="cuda")
    nd = torch.empty(b, dtype=torch.int32, device="cuda")
    th = (-np.pi, np.pi) if cfg.is_bicycle else (0.0, 2 * np.pi)
    cc = util.c_cfg(cfg)
    rc = _lib.lib().dgppo_reset(util.stream(), C.byref(cc), util.p(k), 0.1, 0.3, th[0], th[1], util.p(agent),
                                util.p(goal), util.p(obst) if cfg.n_obs > 0 else None, util.p(nd), b)
    assert rc == 0
    torch.cuda.synchronize()
    return agent.cpu().numpy(), goal.cpu().numpy(), obst.cpu().numpy()[:, :cfg.n_obs], nd.cpu().numpy()


@pytest.mark.parametrize("name", list(CASES))
def test_reset_matches_oracle(name):
    cfg = CASES[name]
    keys = np.array([0, 1, 2, 12345, 2 ** 40 + 7, 2 ** 63 + 11], np.uint64)
    agent, goal, obst, nd = _run(cfg, keys)
    for i, key in enumerate(keys):
        ra, rg, ro, rn = reset_np.reset_states(cfg, int(key))
        assert nd[i] == rn, f"{name} key {key}: number of draws"
        np.testing.assert_array_equal(agent[i][:, :2], ra[:, :2])
        np.testing.assert_array_equal(goal[i], rg)
        if cfg.is_bicycle:
            np.testing.assert_allclose(agent[i][:, 2:], ra[:, 2:], rtol=1e-6, atol=1e-6)   # libm cos/sin
        if cfg.n_obs > 0:
            if cfg.is_lidar:
                np.testing.assert_array_equal(obst[i][:, :5], ro[:, :5])
                np.testing.assert_allclose(obst[i][:, 5:], ro[:, 5:], rtol=1e-6, atol=1e-6)
            else:
                np.testing.assert_array_equal(obst[i], ro)


@pytest.mark.parametrize("name", ["LidarSpread", "MPESpread", "crowded"])
def test_reset_invariants_at_scale(name):
    cfg = CASES[name]
    b = 2048
    agent, goal, obst, nd = _run(cfg, np.arange(b, dtype=np.uint64) * 977 + 5)
    min_dist = F((2.2 if cfg.is_lidar else 2.0) * cfg.car_radius)
    for pts in (agent[..., :2], goal[..., :2]):
        d = np.linalg.norm(pts[:, :, None] - pts[:, None], axis=-1) + np.eye(cfg.n) * 10
        assert (d > min_dist).all()
        assert (pts >= 0).all() and (pts <= cfg.area).all()
        assert (np.linalg.norm(pts, axis=-1) > min_dist).all()       # the origin repels (zero-initialised slots)
    if cfg.is_lidar:
        ob = dict(center=obst[..., 0:2], width=obst[..., 2], height=obst[..., 3], cos=obst[..., 5], sin=obst[..., 6])
        assert not env_np.rect_inside(agent[..., :2], ob, float(min_dist) / 2).any()
        assert not env_np.rect_inside(goal[..., :2], ob, float(min_dist) / 2).any()
        assert (obst[..., 2:4] >= 0.1).all() and (obst[..., 2:4] <= 0.3).all()
    else:
        da = np.linalg.norm(obst[:, :, None, :2] - agent[:, None, :, :2], axis=-1)
        assert (da > cfg.car_radius + cfg.obs_radius).all()
        assert (obst[..., :2] >= 3 * cfg.car_radius - 1e-7).all()
    assert (nd >= 2 * cfg.n).all()
    # different keys give different layouts; the same key is reproducible
    a2, _, _, _ = _run(cfg, np.arange(b, dtype=np.uint64) * 977 + 5)
    np.testing.assert_array_equal(agent, a2)
    assert len({tuple(x) for x in agent[:, 0, :2].round(6)}) > b * 0.99


def test_reset_infeasible_area_terminates_and_is_flagged():
    """64 agents + 64 goals + 64 obstacles do not fit the default 1.5 x 1.5 area: the reference's sampler
    restarts for ever (env/utils.py:229-232); the kernel gives up after 16 restarts, flags the env through
    n_draws = -1, and the host API raises."""
    cfg = env_np.EnvCfg(env_np.LIDAR_SPREAD, n=64, n_obs=64)
    _, _, _, nd = _run(cfg, np.arange(8, dtype=np.uint64))
    assert (nd == -1).all()
    from dgppo_b200.env import make_env
    env = make_env("LidarSpread", num_agents=64, num_obs=64)
    with pytest.raises(RuntimeError, match="cannot hold"):
        env.reset(np.arange(4, dtype=np.uint64))


def test_collect_raises_for_infeasible_area():
    """algo.collect resets with the feasibility check deferred past the rollout launches; it must still raise."""
    from dgppo_b200.algo import make_algo
    from dgppo_b200.env import make_env
    env = make_env("LidarSpread", num_agents=64, num_obs=64, max_step=2)
    algo = make_algo("dgppo", env=env, node_dim=env.node_dim, edge_dim=env.edge_dim, state_dim=env.state_dim,
                     action_dim=env.action_dim, n_agents=64, batch_size=4)
    with pytest.raises(RuntimeError, match="cannot hold"):
        algo.collect(algo.params, np.arange(4, dtype=np.uint64))
    # and a feasible environment leaves no stale flag behind
    env2 = make_env("LidarSpread", num_agents=3, num_obs=3, max_step=2)
    algo2 = make_algo("dgppo", env=env2, node_dim=7, edge_dim=4, state_dim=4, action_dim=2, n_agents=3, batch_size=4)
    ro = algo2.collect(algo2.params, np.arange(4, dtype=np.uint64))
    assert ro.actions.shape == (4, 2, 3, 2)


LANDMARK_CASES = {
    "LidarLine": env_np.EnvCfg(env_np.LIDAR_LINE, n=4, n_obs=3),
    "LidarLine_n7": env_np.EnvCfg(env_np.LIDAR_LINE, n=7, n_obs=4),
    "MPELine_n3": env_np.EnvCfg(env_np.MPE_LINE, n=3, n_obs=3),
    "MPELine_n5": env_np.EnvCfg(env_np.MPE_LINE, n=5, n_obs=3),
    "MPEFormation": env_np.EnvCfg(env_np.MPE_FORMATION, n=4, n_obs=3),
    "MPEConnectSpread": env_np.EnvCfg(env_np.MPE_CONNECT_SPREAD, n=3, n_obs=1, area=1.0, obs_radius=0.25),
}


def _run_landmark(cfg, keys):
    b = len(keys)
    k = torch.from_numpy(np.asarray(keys, np.uint64).astype(np.int64)).cuda()
    agent = torch.empty((b, cfg.n, 4), device="cuda")
    goal = torch.empty((b, cfg.n_goal, 4), device="cuda")
    w = _lib.OBS_STRIDE if cfg.is_lidar else 4
    obst = torch.empty((b, cfg.n_obs, w), device="cuda")
    nd = torch.empty(b, dtype=torch.int32, device="cuda")
    cc = util.c_cfg(cfg)
    rc = _lib.lib().dgppo_reset(util.stream(), C.byref(cc), util.p(k), 0.1, 0.3, 0.0, np.pi, util.p(agent),
                                util.p(goal), util.p(obst), util.p(nd), b)
    assert rc == 0
    torch.cuda.synchronize()
    return agent.cpu().numpy(), goal.cpu().numpy(), obst.cpu().numpy(), nd.cpu().numpy()


@pytest.mark.parametrize("name", list(LANDMARK_CASES))
def test_landmark_reset_matches_oracle(name):
    """Line / Formation / ConnectSpread samplers (lidar_line.py:38-126, mpe_line.py:38-117,
    mpe_formation.py:38-91, mpe_connect_spread.py:52-103) against their restatement, same counter stream."""
    cfg = LANDMARK_CASES[name]
    keys = np.array([0, 1, 2, 12345, 2 ** 40 + 7, 2 ** 63 + 11], np.uint64)
    agent, goal, obst, nd = _run_landmark(cfg, keys)
    for i, key in enumerate(keys):
        ra, rg, ro, rn = reset_np.reset_landmark_states(cfg, int(key))
        assert nd[i] == rn, f"{name} key {key}: number of draws"
        np.testing.assert_array_equal(agent[i], ra)
        np.testing.assert_array_equal(goal[i], rg)
        if cfg.is_lidar:
            np.testing.assert_array_equal(obst[i][:, :5], ro[:, :5])
            np.testing.assert_allclose(obst[i][:, 5:], ro[:, 5:], rtol=1e-6, atol=1e-6)     # libm cos / sin
        else:
            np.testing.assert_array_equal(obst[i], ro)


@pytest.mark.parametrize("name", list(LANDMARK_CASES))
def test_landmark_reset_invariants(name):
    cfg = LANDMARK_CASES[name]
    keys = np.arange(256, dtype=np.uint64) * 7919 + 3
    agent, goal, obst, nd = _run_landmark(cfg, keys)
    assert (nd > 0).all()
    pos = agent[..., :2]
    n = cfg.n
    dmin = (np.linalg.norm(pos[:, :, None] - pos[:, None], axis=-1) + np.eye(n) * 10).min(-1)
    assert (dmin > (2.3 if cfg.kind == env_np.MPE_CONNECT_SPREAD else 2.0) * cfg.car_radius).all()
    assert (pos >= 0).all() and (pos[..., 0] <= cfg.area).all()
    if cfg.kind == env_np.MPE_CONNECT_SPREAD:
        assert (dmin <= cfg.connect_radius).all()                                    # agents connected
        g = goal[..., :2]
        gmin = (np.linalg.norm(g[:, :, None] - g[:, None], axis=-1) + np.eye(n) * 10).min(-1)
        assert (gmin <= cfg.connect_radius).all()
        assert (obst[:, 0, 1] == np.float32(cfg.area / 2)).all()
        assert (g[..., 1] > pos[..., 1].max()).all() or (g[..., 1].min() > cfg.area / 2)   # goals past the obstacle
    else:
        eg = env_np.landmark2goal(cfg, goal[..., :2])
        if cfg.kind in (env_np.LIDAR_LINE, env_np.MPE_LINE):
            short = cfg.kind == env_np.MPE_LINE and n <= 3
            lm_min = n * 5 * cfg.car_radius if short else (n - 2) * 6 * cfg.car_radius
            assert (np.linalg.norm(goal[:, 1, :2] - goal[:, 0, :2], axis=-1) >= np.float32(lm_min) - 1e-6).all()
        if cfg.is_lidar:
            th = obst[..., 4]
            ob = dict(center=obst[..., 0:2], width=obst[..., 2], height=obst[..., 3], cos=obst[..., 5], sin=obst[..., 6])
            pts = np.concatenate([pos, eg], axis=1)
            assert not env_np.rect_inside(pts, ob, np.float32(cfg.car_radius) * np.float32(1.1)).any()
            assert (th >= 0).all() and (th <= np.pi + 1e-6).all()
        else:
            d_a = np.linalg.norm(pos[:, :, None] - obst[:, None, :, :2], axis=-1)
            d_g = np.linalg.norm(eg[:, :, None] - obst[:, None, :, :2], axis=-1)
            assert (d_a > cfg.car_radius + cfg.obs_radius).all() and (d_g > 2 * cfg.car_radius + cfg.obs_radius).all()
