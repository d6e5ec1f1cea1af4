"""Test helpers: oracle <-> C-ABI plumbing (the oracle is the checker only)."""
import ctypes as C

import numpy as np
import torch

from dgppo_b200 import _lib
from oracle import env_np

F = np.float32

CONFIGS = {   # BASELINE.json configs at test sizes (b small), + edge cases
    "C1": env_np.EnvCfg(env_np.LIDAR_SPREAD, n=3, n_obs=3),
    "C2": env_np.EnvCfg(env_np.MPE_SPREAD, n=8, n_obs=3),
    "C3": env_np.EnvCfg(env_np.LIDAR_SPREAD, n=8, n_obs=8),
    "C4": env_np.EnvCfg(env_np.LIDAR_BICYCLE_TARGET, n=16, n_obs=3),
    "C5": env_np.EnvCfg(env_np.LIDAR_SPREAD, n=64, n_obs=64),
    "target": env_np.EnvCfg(env_np.LIDAR_TARGET, n=5, n_obs=2),
    "noobs_lidar": env_np.EnvCfg(env_np.LIDAR_SPREAD, n=4, n_obs=0),
    "noobs_mpe": env_np.EnvCfg(env_np.MPE_SPREAD, n=4, n_obs=0),
    "fullobs": env_np.EnvCfg(env_np.LIDAR_SPREAD, n=6, n_obs=4, comm_radius=15.0),
    "one_agent": env_np.EnvCfg(env_np.LIDAR_TARGET, n=1, n_obs=1),
    "rays48": env_np.EnvCfg(env_np.LIDAR_SPREAD, n=5, n_obs=3, n_rays=48, top_k=8),
    "mpe_target": env_np.EnvCfg(env_np.MPE_TARGET, n=6, n_obs=3),
    "mpe_corridor": env_np.EnvCfg(env_np.MPE_CORRIDOR, n=5, n_obs=2, area=1.0, obs_radius=0.2),
    # n > 16: the one-graph-per-tile GNN kernel (row chunks of 16, ragged last chunk)
    "n24": env_np.EnvCfg(env_np.LIDAR_SPREAD, n=24, n_obs=5),
    "mpe20": env_np.EnvCfg(env_np.MPE_SPREAD, n=20, n_obs=3),
    "target40": env_np.EnvCfg(env_np.LIDAR_TARGET, n=40, n_obs=4),
    "dense20": env_np.EnvCfg(env_np.LIDAR_SPREAD, n=20, n_obs=4, comm_radius=15.0),
}


_TABLES = {}      # device goal tables stay alive for the test session


def c_cfg(cfg: env_np.EnvCfg) -> _lib.DgppoEnvCfg:
    table = None
    if cfg.kind == env_np.MPE_FORMATION:
        key = (cfg.n, cfg.comm_radius)
        if key not in _TABLES:
            _TABLES[key] = torch.as_tensor(env_np.formation_offsets(cfg)).cuda().contiguous()
        table = _TABLES[key].data_ptr()
    return _lib.DgppoEnvCfg(cfg.kind, cfg.n, cfg.n_obs, cfg.n_rays, cfg.top_k, 0, cfg.comm_radius,
                            cfg.car_radius, cfg.obs_radius, cfg.area, cfg.dt, cfg.dist2goal,
                            cfg.connect_radius, table)


def obs_record(obs: dict) -> np.ndarray:
    b, o = obs["center"].shape[:2]
    rec = np.zeros((b, o, _lib.OBS_STRIDE), F)
    rec[..., 0:2] = obs["center"]
    rec[..., 2], rec[..., 3], rec[..., 4] = obs["width"], obs["height"], obs["theta"]
    rec[..., 5], rec[..., 6] = obs["cos"], obs["sin"]
    rec[..., 8:16] = obs["points"].reshape(b, o, 8)
    return rec


def dev(x, dtype=None):
    if x is None:
        return None
    t = torch.as_tensor(np.ascontiguousarray(x))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda().contiguous()


def p(t):
    return None if t is None else t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def k_env_step(cfg, agent, goal, obs_nodes, action):
    b, n = agent.shape[:2]
    a, g, o, u = dev(agent), dev(goal), dev(obs_nodes), dev(action)
    nxt = torch.empty_like(a)
    rew = torch.empty(b, device="cuda")
    cost = torch.empty((b, n, cfg.n_cost), device="cuda")
    cc = c_cfg(cfg)
    _lib.check(_lib.lib().dgppo_env_step(stream(), C.byref(cc), p(a), p(g), p(o), p(u), p(nxt), p(rew), p(cost), 1, b),
               "env_step")
    torch.cuda.synchronize()
    return nxt.cpu().numpy(), rew.cpu().numpy(), cost.cpu().numpy()


def k_lidar(cfg, agent, obstacles, rays):
    b, n = agent.shape[:2]
    a, rec, r = dev(agent), dev(obs_record(obstacles)), dev(rays)
    hits = torch.full((b, n, cfg.top_k, 2), float("nan"), device="cuda")
    cc = c_cfg(cfg)
    _lib.check(_lib.lib().dgppo_lidar(stream(), C.byref(cc), p(a), p(rec), p(r), p(hits), b), "lidar")
    torch.cuda.synchronize()
    return hits.cpu().numpy()


def k_graph(cfg, agent, goal, obs_nodes):
    b = agent.shape[0]
    a, g, o = dev(agent), dev(goal), dev(obs_nodes)
    N, E = cfg.n_nodes, cfg.n_edges
    out = dict(nodes=torch.empty((b, N, cfg.node_dim), device="cuda"), edges=torch.empty((b, E, 4), device="cuda"),
               states=torch.empty((b, N, cfg.state_dim), device="cuda"),
               receivers=torch.empty((b, E), dtype=torch.int32, device="cuda"),
               senders=torch.empty((b, E), dtype=torch.int32, device="cuda"),
               node_type=torch.empty((b, N), dtype=torch.int32, device="cuda"),
               n_node=torch.empty(b, dtype=torch.int32, device="cuda"),
               n_edge=torch.empty(b, dtype=torch.int32, device="cuda"))
    cc = c_cfg(cfg)
    _lib.check(_lib.lib().dgppo_build_graph(stream(), C.byref(cc), p(a), p(g), p(o), p(out["nodes"]), p(out["edges"]),
                                             p(out["states"]), p(out["receivers"]), p(out["senders"]),
                                             p(out["node_type"]), p(out["n_node"]), p(out["n_edge"]), 1, b), "graph")
    torch.cuda.synchronize()
    return {k: v.cpu().numpy() for k, v in out.items()}


def k_policy(cfg, net_cfg, packed, graph, rnn, eps):
    b = graph["nodes"].shape[0]
    n = cfg.n
    nodes, edges = dev(graph["nodes"]), dev(graph["edges"])
    recv, send = dev(graph["receivers"], torch.int32), dev(graph["senders"], torch.int32)
    rnn_in, e = dev(rnn), dev(eps)
    rnn_out = torch.empty_like(rnn_in)
    action = torch.empty((b, n, 2), device="cuda")
    log_pi = torch.empty((b, n), device="cuda")
    pk = dev(packed)
    cc = c_cfg(cfg)
    _lib.check(_lib.lib().dgppo_gnn_policy(stream(), C.byref(cc), C.byref(net_cfg), p(pk), p(nodes), p(edges), p(recv),
                                            p(send), 1, p(rnn_in), p(rnn_out), 1, p(e), 1, p(action),
                                            p(log_pi) if eps is not None else None, 1, b), "policy")
    torch.cuda.synchronize()
    return action.cpu().numpy(), (log_pi.cpu().numpy() if eps is not None else None), rnn_out.cpu().numpy()


def k_value(cfg, net_cfg, packed, graph, rnn):
    b = graph["nodes"].shape[0]
    n = cfg.n
    nodes, edges = dev(graph["nodes"]), dev(graph["edges"])
    recv, send = dev(graph["receivers"], torch.int32), dev(graph["senders"], torch.int32)
    rnn_in = dev(rnn)
    vl = net_cfg.kind == _lib.NET_VL
    rnn_out = torch.empty_like(rnn_in)
    val = torch.empty((b,) if vl else (b, n, net_cfg.n_out), device="cuda")
    pk = dev(packed)
    cc = c_cfg(cfg)
    _lib.check(_lib.lib().dgppo_gnn_value(stream(), C.byref(cc), C.byref(net_cfg), p(pk), p(nodes), p(edges), p(recv),
                                           p(send), 1, p(rnn_in), p(rnn_out), 1, p(val), 1, 1, b), "value")
    torch.cuda.synchronize()
    return val.cpu().numpy(), rnn_out.cpu().numpy()


def assert_bits_equal(a, b, what=""):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    assert a.shape == b.shape and a.dtype == b.dtype, (what, a.shape, b.shape, a.dtype, b.dtype)
    if a.dtype.kind == "f":
        ai, bi = a.view(np.int32), b.view(np.int32)
        # NaN payloads may differ; -0.0 vs 0.0 never arise from the same op sequence
        bad = (ai != bi) & ~(np.isnan(a) & np.isnan(b))
    else:
        bad = a != b
    if bad.any():
        idx = np.argwhere(bad)[:5]
        raise AssertionError(f"{what}: {bad.sum()} of {bad.size} elements differ, first at {idx.tolist()}: "
                             f"{a[tuple(idx[0])]!r} vs {b[tuple(idx[0])]!r}")


def threshold_states(cfg: env_np.EnvCfg, b: int, seed: int):
    """Synthetic states where agent pairs sit within a few ulps of the comm
    radius and agents sit on obstacle boundaries, to stress the masks."""
    agent, goal, obstacles, mpe_obs = env_np.synthetic_states(cfg, b, seed)
    rng = np.random.default_rng(seed + 1000)
    n = cfg.n
    if n >= 2:
        ang = rng.uniform(0, 2 * np.pi, b).astype(F)
        r = np.nextafter(np.full(b, cfg.comm_radius, F), rng.choice([0, 1], b).astype(F))
        r = np.where(rng.random(b) < 0.3, F(cfg.comm_radius), r).astype(F)
        base = rng.uniform(0.5, 1.0, (b, 2)).astype(F)
        agent[:, 0, :2] = base
        agent[:, 1, 0] = base[:, 0] + r * np.cos(ang)
        agent[:, 1, 1] = base[:, 1] + r * np.sin(ang)
    if n >= 3:
        agent[:, 2, :2] = agent[:, 0, :2]          # coincident agents (distance exactly 0)
    if obstacles is not None and n >= 4:
        agent[:, 3, :2] = obstacles["center"][:, 0]  # inside an obstacle -> all alphas 0
    return agent, goal, obstacles, mpe_obs
