"""NumPy restatement of the reference environments on the rollout hot path.

TEST INFRASTRUCTURE (see oracle/__init__.py).  All citations are relative to
/root/reference/.  Every function is batched over a leading env axis ``b``
(the reference batches single-env code with jax.vmap,
dgppo/algo/informarl.py:183-184); the arithmetic per env is the literal
sequence of fp32 operations the reference source spells out.

Reductions over small axes (means in the reward) are evaluated left to right
(``seq_sum``): XLA's order is not observable from the reference source, and a
defined order lets the CUDA kernels be compared bit for bit.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import numpy as np

F = np.float32

LIDAR_SPREAD = 0          # dgppo/env/lidar_env/lidar_spread.py
LIDAR_TARGET = 1          # dgppo/env/lidar_env/lidar_target.py
LIDAR_BICYCLE_TARGET = 2  # dgppo/env/lidar_env/lidar_bicycle_target.py
MPE_SPREAD = 3            # dgppo/env/mpe/mpe_spread.py
MPE_TARGET = 4            # dgppo/env/mpe/mpe_target.py
MPE_CORRIDOR = 5          # dgppo/env/mpe/mpe_corridor.py
LIDAR_LINE = 6            # dgppo/env/lidar_env/lidar_line.py     (2 landmarks -> n goals on the segment)
MPE_LINE = 7              # dgppo/env/mpe/mpe_line.py
MPE_FORMATION = 8         # dgppo/env/mpe/mpe_formation.py        (1 landmark -> n goals on a circle)
MPE_CONNECT_SPREAD = 9    # dgppo/env/mpe/mpe_connect_spread.py   (third cost: connectivity)

KIND_BY_NAME = {
    "LidarSpread": LIDAR_SPREAD,
    "LidarTarget": LIDAR_TARGET,
    "LidarBicycleTarget": LIDAR_BICYCLE_TARGET,
    "MPESpread": MPE_SPREAD,
    "MPETarget": MPE_TARGET,
    "MPECorridor": MPE_CORRIDOR,
    "LidarLine": LIDAR_LINE,
    "MPELine": MPE_LINE,
    "MPEFormation": MPE_FORMATION,
    "MPEConnectSpread": MPE_CONNECT_SPREAD,
}


@dataclass(frozen=True)
class EnvCfg:
    """Static env description (PARAMS dicts: lidar_spread.py:13-22,
    mpe_spread.py:12-19; dt/max_step: env/__init__.py:26,51)."""
    kind: int
    n: int
    n_obs: int = 3
    n_rays: int = 32
    top_k: int = 8
    comm_radius: float = 0.5
    car_radius: float = 0.05
    obs_radius: float = 0.05
    area: float = 1.5
    dt: float = 0.03
    dist2goal: float = 0.01
    max_step: int = 128
    connect_radius: float = 0.45

    @property
    def is_lidar(self) -> bool:
        return self.kind in (LIDAR_SPREAD, LIDAR_TARGET, LIDAR_BICYCLE_TARGET, LIDAR_LINE)

    @property
    def n_cost(self) -> int:             # mpe_connect_spread.py:45-46
        return 3 if self.kind == MPE_CONNECT_SPREAD else 2

    @property
    def is_bicycle(self) -> bool:
        return self.kind == LIDAR_BICYCLE_TARGET

    @property
    def state_dim(self) -> int:
        return 5 if self.is_bicycle else 4

    @property
    def node_dim(self) -> int:
        return self.state_dim + 3

    @property
    def edge_dim(self) -> int:
        return 4

    @property
    def n_goal(self) -> int:             # goal NODES: landmarks for the Line / Formation families
        if self.kind in (LIDAR_LINE, MPE_LINE):
            return 2                     # lidar_line.py:36, mpe_line.py:36
        if self.kind == MPE_FORMATION:
            return 1                     # mpe_formation.py:36
        return self.n

    @property
    def n_hits(self) -> int:
        return self.top_k * self.n if (self.is_lidar and self.n_obs > 0) else 0

    @property
    def n_obs_nodes(self) -> int:
        return self.n_hits if self.is_lidar else self.n_obs

    @property
    def n_nodes(self) -> int:            # including the pad node
        return self.n + self.n_goal + self.n_obs_nodes + 1

    @property
    def n_ag(self) -> int:               # goal senders per agent
        return 1 if self.kind in (LIDAR_TARGET, LIDAR_BICYCLE_TARGET, MPE_TARGET) else self.n_goal

    @property
    def n_ao(self) -> int:               # obstacle senders per agent
        if self.n_obs == 0:
            return 0
        return self.top_k if self.is_lidar else self.n_obs

    @property
    def n_edges(self) -> int:
        return self.n * self.n + self.n * self.n_ag + self.n * self.n_ao


def seq_sum(x: np.ndarray, axis: int = -1) -> np.ndarray:
    """fp32 left-to-right sum over ``axis``."""
    x = np.moveaxis(x, axis, -1)
    acc = x[..., 0].astype(F).copy()
    for k in range(1, x.shape[-1]):
        acc = (acc + x[..., k]).astype(F)
    return acc


def seq_mean(x: np.ndarray, axis: int = -1) -> np.ndarray:
    return (seq_sum(x, axis) / F(x.shape[axis])).astype(F)


def norm2(dx: np.ndarray, dy: np.ndarray) -> np.ndarray:
    """jnp.linalg.norm over a 2-vector: sqrt(dx*dx + dy*dy), each op rounded."""
    return np.sqrt((dx * dx).astype(F) + (dy * dy).astype(F)).astype(F)


# ---------------------------------------------------------------- obstacles
def rect_create(center, width, height, theta) -> Dict[str, np.ndarray]:
    """Rectangle.create (dgppo/env/obstacle.py:39-56), batched over leading
    axes.  Also stores cos/sin(theta): ``Rectangle.inside`` re-evaluates them
    on every call (obstacle.py:65-66); the obstacle is static over an episode,
    so the values are carried as data to keep libm differences out of the
    inside test."""
    center = np.asarray(center, F)
    width = np.asarray(width, F)
    height = np.asarray(height, F)
    theta = np.asarray(theta, F)
    c = np.cos(theta).astype(F)
    s = np.sin(theta).astype(F)
    hw = (width / F(2)).astype(F)
    hh = (height / F(2)).astype(F)
    bx = np.stack([hw, -hw, -hw, hw], axis=-1)          # (+,+),(-,+),(-,-),(+,-)
    by = np.stack([hh, hh, -hh, -hh], axis=-1)
    px = ((c[..., None] * bx).astype(F) + ((-s)[..., None] * by).astype(F)).astype(F)
    py = ((s[..., None] * bx).astype(F) + (c[..., None] * by).astype(F)).astype(F)
    px = (px + center[..., 0:1]).astype(F)
    py = (py + center[..., 1:2]).astype(F)
    return dict(center=center, width=width, height=height, theta=theta,
                cos=c, sin=s, points=np.stack([px, py], axis=-1).astype(F))


def rect_inside(pos: np.ndarray, obs: Dict[str, np.ndarray], r: float = 0.0) -> np.ndarray:
    """inside_obstacles / Rectangle.inside (env/utils.py:82-112,
    obstacle.py:62-72).  pos (b,n,2), obstacle fields (b,o,...) -> bool (b,n)."""
    rel_x = (pos[:, :, None, 0] - obs["center"][:, None, :, 0]).astype(F)
    rel_y = (pos[:, :, None, 1] - obs["center"][:, None, :, 1]).astype(F)
    c = obs["cos"][:, None, :]
    s = obs["sin"][:, None, :]
    hw = (obs["width"] / F(2)).astype(F)[:, None, :]
    hh = (obs["height"] / F(2)).astype(F)[:, None, :]
    rel_xx = (np.abs(((rel_x * c).astype(F) + (rel_y * s).astype(F)).astype(F)) - hw).astype(F)
    rel_yy = (np.abs(((rel_x * s).astype(F) - (rel_y * c).astype(F)).astype(F)) - hh).astype(F)
    r = F(r)
    in_down = (rel_xx < r) & (rel_yy < 0)
    in_up = (rel_xx < 0) & (rel_yy < r)
    out_corner = (rel_xx > 0) & (rel_yy > 0)
    in_circle = np.sqrt((rel_xx * rel_xx).astype(F) + (rel_yy * rel_yy).astype(F)).astype(F) < r
    is_in = in_down | in_up | (out_corner & in_circle)
    return is_in.any(axis=-1)


def ray_table(n_rays: int, sense_range: float) -> np.ndarray:
    """Ray end offsets (cos(theta)*range, sin(theta)*range), get_lidar
    (env/utils.py:51-55).  The table is DATA for both the oracle and the CUDA
    kernel (cos/sin differ by an ulp between libms)."""
    thetas = np.linspace(-np.pi, np.pi - 2 * np.pi / n_rays, n_rays).astype(F)
    return np.stack([(np.cos(thetas).astype(F) * F(sense_range)).astype(F),
                     (np.sin(thetas).astype(F) * F(sense_range)).astype(F)], axis=-1)


def lidar_alphas(pos: np.ndarray, obs: Dict[str, np.ndarray], rays: np.ndarray) -> np.ndarray:
    """Per-ray hit parameter alpha (b,n,R): raytracing + Rectangle.raytracing
    (env/utils.py:115-129, obstacle.py:74-105)."""
    with np.errstate(all="ignore"):
        x1 = pos[:, :, 0][:, :, None, None, None]          # (b,n,1,1,1)
        y1 = pos[:, :, 1][:, :, None, None, None]
        x2 = (pos[:, :, 0:1] + rays[None, None, :, 0]).astype(F)[:, :, :, None, None]   # (b,n,R,1,1)
        y2 = (pos[:, :, 1:2] + rays[None, None, :, 1]).astype(F)[:, :, :, None, None]
        P = obs["points"]                                   # (b,o,4,2)
        x3 = P[:, None, None, :, :, 0]                      # (b,1,1,o,4)
        y3 = P[:, None, None, :, :, 1]
        P4 = P[:, :, [3, 0, 1, 2], :]
        x4 = P4[:, None, None, :, :, 0]
        y4 = P4[:, None, None, :, :, 1]
        dx12 = (x1 - x2).astype(F)
        dy12 = (y1 - y2).astype(F)
        dx43 = (x4 - x3).astype(F)
        dy43 = (y4 - y3).astype(F)
        dx13 = (x1 - x3).astype(F)
        dy13 = (y1 - y3).astype(F)
        det = ((dx12 * dy43).astype(F) - (dy12 * dx43).astype(F)).astype(F)
        det = (np.sign(det) * np.clip(np.abs(det), F(1e-7), F(1e7))).astype(F)
        alphas = (((dy43 * dx13).astype(F) - (dx43 * dy13).astype(F)).astype(F) / det).astype(F)
        betas = ((((-dy12) * dx13).astype(F) + (dx12 * dy13).astype(F)).astype(F) / det).astype(F)
        valids = (alphas <= 1) & (alphas >= 0) & (betas <= 1) & (betas >= 0)
        vf = valids.astype(F)
        alphas = ((vf * alphas).astype(F) + ((F(1) - vf) * F(1e6)).astype(F)).astype(F)
        alphas = alphas.min(axis=-1).min(axis=-1)           # edges, then obstacles (NaN propagates)
        is_in = rect_inside(pos, obs, 0.0)
        alphas = (alphas * (F(1) - is_in.astype(F))[:, :, None]).astype(F)
    return alphas


def lidar_hits(cfg: EnvCfg, pos: np.ndarray, obs: Dict[str, np.ndarray],
               rays: Optional[np.ndarray] = None, chunk: int = 256) -> np.ndarray:
    """get_lidar_data (lidar_env/base.py:126-140) -> (b,n,top_k,2): the top_k
    hit points with the smallest alpha, stable order (env/utils.py:132-136)."""
    if rays is None:
        rays = ray_table(cfg.n_rays, cfg.comm_radius)
    b, n, _ = pos.shape
    out = np.empty((b, n, cfg.top_k, 2), F)
    for s in range(0, b, chunk):
        e = min(b, s + chunk)
        p = pos[s:e]
        o = {k: v[s:e] for k, v in obs.items()}
        with np.errstate(all="ignore"):
            al = lidar_alphas(p, o, rays)                   # (c,n,R)
            idx = np.argsort(al, axis=-1, kind="stable")[..., :cfg.top_k]
            x2 = (p[:, :, 0:1] + rays[None, None, :, 0]).astype(F)
            y2 = (p[:, :, 1:2] + rays[None, None, :, 1]).astype(F)
            hx = (p[:, :, 0:1] + ((x2 - p[:, :, 0:1]).astype(F) * al).astype(F)).astype(F)
            hy = (p[:, :, 1:2] + ((y2 - p[:, :, 1:2]).astype(F) * al).astype(F)).astype(F)
        out[s:e, :, :, 0] = np.take_along_axis(hx, idx, axis=-1)
        out[s:e, :, :, 1] = np.take_along_axis(hy, idx, axis=-1)
    return out


# ----------------------------------------------------------------- dynamics
def state_lim(cfg: EnvCfg) -> Tuple[np.ndarray, np.ndarray]:
    """state_lim (lidar_env/base.py:273-276, mpe/base.py:243-246,
    lidar_bicycle_target.py:120-123)."""
    A = cfg.area
    if cfg.is_bicycle:
        return np.array([0, 0, -1, -1, -0.5], F), np.array([A, A, 1, 1, 0.5], F)
    if cfg.kind in (MPE_CORRIDOR, MPE_CONNECT_SPREAD):  # mpe_corridor.py:64-67, mpe_connect_spread.py:140-143
        return np.array([0, 0, -1, -1], F), np.array([A, A * 2, 1, 1], F)
    if not cfg.is_lidar:
        return np.array([0, 0, -1, -1], F), np.array([A, A, 1, 1], F)
    return np.array([0, 0, -0.5, -0.5], F), np.array([A, A, 0.5, 0.5], F)


def clip_action(action: np.ndarray) -> np.ndarray:
    """clip_action with action_lim = [-1,1]^2 (env/base.py:84-86)."""
    return np.clip(action.astype(F), F(-1), F(1))


def agent_step_euler(cfg: EnvCfg, agent: np.ndarray, action: np.ndarray) -> np.ndarray:
    """agent_step_euler: double integrator (lidar_env/base.py:142-149 ==
    mpe/base.py:129-135) or bicycle (lidar_bicycle_target.py:92-111)."""
    dt = F(cfg.dt)
    lo, hi = state_lim(cfg)
    if cfg.is_bicycle:
        x = agent
        theta = np.arctan2(x[..., 3], x[..., 2]).astype(F)
        theta_next = (theta + (((x[..., 4] * action[..., 0]).astype(F) * dt).astype(F) * F(10)).astype(F)).astype(F)
        nx = np.stack([
            (x[..., 0] + ((x[..., 4] * np.cos(theta).astype(F)).astype(F) * dt).astype(F)).astype(F),
            (x[..., 1] + ((x[..., 4] * np.sin(theta).astype(F)).astype(F) * dt).astype(F)).astype(F),
            np.cos(theta_next).astype(F),
            np.sin(theta_next).astype(F),
            (x[..., 4] + ((action[..., 1] * dt).astype(F) * F(10.)).astype(F)).astype(F),
        ], axis=-1)
    else:
        x_dot = np.concatenate([agent[..., 2:], (action * F(10.)).astype(F)], axis=-1)
        nx = ((x_dot * dt).astype(F) + agent).astype(F)
    return np.clip(nx, lo, hi).astype(F)


# ------------------------------------------------------------- cost, reward
def get_cost(cfg: EnvCfg, agent: np.ndarray, obs_nodes: Optional[np.ndarray]) -> np.ndarray:
    """get_cost: Lidar (lidar_env/base.py:180-207; obstacle term from the hit
    nodes stored in the graph) / MPE (mpe/base.py:164-191; obstacle centres).
    agent (b,n,sd); obs_nodes (b,n,top_k,2) hits or (b,n_obs,>=2) centres."""
    b, n, _ = agent.shape
    px, py = agent[..., 0], agent[..., 1]
    dist = norm2((px[:, :, None] - px[:, None, :]).astype(F), (py[:, :, None] - py[:, None, :]).astype(F))
    dist = (dist + (np.eye(n, dtype=F) * F(1e6)).astype(F)[None]).astype(F)
    min_dist = dist.min(axis=2)
    agent_cost = (F(cfg.car_radius * 2) - min_dist).astype(F)
    if cfg.n_obs == 0:
        obs_cost = np.zeros((b, n), F)
    elif cfg.is_lidar:
        d = norm2((obs_nodes[..., 0] - px[:, :, None]).astype(F), (obs_nodes[..., 1] - py[:, :, None]).astype(F))
        obs_cost = (F(cfg.car_radius) - d.min(axis=2)).astype(F)
    else:
        d = norm2((px[:, :, None] - obs_nodes[:, None, :, 0]).astype(F),
                  (py[:, :, None] - obs_nodes[:, None, :, 1]).astype(F))
        obs_cost = (F(cfg.car_radius + cfg.obs_radius) - d.min(axis=2)).astype(F)
    cols = [agent_cost, obs_cost]
    if cfg.kind == MPE_CONNECT_SPREAD:                  # connectivity cost (mpe_connect_spread.py:116-118)
        connect = (min_dist - F(cfg.connect_radius)).astype(F).max(axis=1)
        cols.append(np.broadcast_to(connect[:, None], (b, n)))
    cost = np.stack(cols, axis=-1)
    eps = F(0.5)
    cost = np.where(cost <= 0.0, (cost - eps).astype(F), (cost + eps).astype(F)).astype(F)
    if cfg.is_lidar or cfg.kind == MPE_CONNECT_SPREAD:  # clip to [-1, 1] (lidar_env/base.py:205, mpe_connect_spread.py:135)
        return np.clip(cost, F(-1.0), F(1.0))
    return np.maximum(cost, F(-1.0))                    # a_min only (mpe/base.py:189)


def get_reward(cfg: EnvCfg, agent: np.ndarray, goal: np.ndarray, action: np.ndarray) -> np.ndarray:
    """get_reward: Spread (lidar_spread.py:35-52 == mpe_spread.py:32-49),
    Target (lidar_target.py:35-52).  ``action`` is the CLIPPED action
    (lidar_env/base.py:160,170).  -> (b,)"""
    ax, ay = agent[..., 0], agent[..., 1]
    if cfg.kind in (LIDAR_LINE, MPE_LINE, MPE_FORMATION):
        goal = landmark2goal(cfg, goal[..., :2])
    gx, gy = goal[..., 0], goal[..., 1]
    if cfg.kind not in (LIDAR_TARGET, LIDAR_BICYCLE_TARGET, MPE_TARGET):
        d = norm2((gx[:, :, None] - ax[:, None, :]).astype(F), (gy[:, :, None] - ay[:, None, :]).astype(F))
        dist2goal = d.min(axis=2)
    else:
        dist2goal = norm2((gx - ax).astype(F), (gy - ay).astype(F))
    reward = np.zeros(agent.shape[0], F)
    reward = (reward - (seq_mean(dist2goal) * F(0.01)).astype(F)).astype(F)
    far = np.where(dist2goal > F(cfg.dist2goal), F(1.0), F(0.0)).astype(F)
    reward = (reward - (seq_mean(far) * F(0.001)).astype(F)).astype(F)
    an = norm2(action[..., 0], action[..., 1])
    reward = (reward - (seq_mean((an * an).astype(F)) * F(0.0001)).astype(F)).astype(F)
    return reward


def formation_offsets(cfg: EnvCfg) -> np.ndarray:
    """R * [cos, sin](linspace(0, 2 pi, n + 1)[:-1]) (mpe_formation.py:93-96), fp32: a per-env-config TABLE
    (the kernels take it as data, like the LiDAR ray table, so no device libm enters the goal positions)."""
    th = np.linspace(0, 2 * np.pi, cfg.n + 1).astype(F)[:-1]
    return (F(cfg.comm_radius) * np.stack([np.cos(th).astype(F), np.sin(th).astype(F)], axis=-1)).astype(F)


def landmark2goal(cfg: EnvCfg, lm: np.ndarray) -> np.ndarray:
    """landmark2goal: the n goal positions the reward uses, from the landmark nodes lm (b, n_goal, 2).
    Line (lidar_line.py:128-133, mpe_line.py:119-128): l0 + k * (l1 - l0) / n_interval; Formation
    (mpe_formation.py:93-96): landmark + R [cos, sin](theta_k).  Each op rounded to fp32."""
    n = cfg.n
    if cfg.kind == MPE_FORMATION:
        return (lm[:, 0:1, :] + formation_offsets(cfg)[None]).astype(F)
    direction = (lm[:, 1] - lm[:, 0]).astype(F)
    if cfg.kind == MPE_LINE and n <= 3:
        n_int, ks = n + 1, np.arange(1, n + 1)
    else:
        n_int, ks = n - 1, np.arange(0, n)
    step = ((ks.astype(F)[None, :, None] * direction[:, None, :]).astype(F) / F(n_int)).astype(F)
    return (lm[:, 0:1, :] + step).astype(F)


# -------------------------------------------------------------------- graph
def state2feat(cfg: EnvCfg, state: np.ndarray) -> np.ndarray:
    """state2feat: identity (lidar_spread.py:54-55) or [x,y,v cos,v sin]
    (lidar_bicycle_target.py:113-118)."""
    if cfg.is_bicycle:
        return np.stack([state[..., 0], state[..., 1],
                         (state[..., 4] * state[..., 2]).astype(F),
                         (state[..., 4] * state[..., 3]).astype(F)], axis=-1)
    return state


def get_graph(cfg: EnvCfg, agent: np.ndarray, goal: np.ndarray,
              obs_nodes: Optional[np.ndarray]) -> Dict[str, np.ndarray]:
    """get_graph + edge_blocks + GetGraph.to_padded
    (lidar_env/base.py:227-271, mpe/base.py:211-241, lidar_spread.py:57-96,
    lidar_target.py:57-96, mpe_spread.py:51-81, utils/graph.py:35-44,212-247).

    obs_nodes: Lidar (b,n,top_k,2) hit points; MPE (b,n_obs,4) obstacle
    states; None when n_obs == 0.  Returns the GraphsTuple array fields."""
    b, n, sd = agent.shape
    g = cfg.n_goal
    nd, N, E = cfg.node_dim, cfg.n_nodes, cfg.n_edges
    pad = N - 1
    n_on = cfg.n_obs_nodes

    nodes = np.zeros((b, N, nd), F)
    states = np.zeros((b, N, sd), F)
    node_type = np.full((b, N), -1, np.int32)
    nodes[:, :n, :sd] = agent
    nodes[:, n:n + g, :sd] = goal
    nodes[:, :n, sd + 2] = 1.0
    nodes[:, n:n + g, sd + 1] = 1.0
    states[:, :n] = agent
    states[:, n:n + g] = goal
    node_type[:, :n] = 0
    node_type[:, n:n + g] = 1
    if n_on > 0:
        if cfg.is_lidar:
            flat = obs_nodes.reshape(b, n_on, 2)
            nodes[:, n + g:n + g + n_on, :2] = flat
            states[:, n + g:n + g + n_on, :2] = flat
        else:
            nodes[:, n + g:n + g + n_on, :sd] = obs_nodes
            states[:, n + g:n + g + n_on] = obs_nodes
        nodes[:, n + g:n + g + n_on, sd] = 1.0
        node_type[:, n + g:n + g + n_on] = 2
    states[:, pad] = -1.0

    edges = np.zeros((b, E, 4), F)
    recv = np.full((b, E), pad, np.int32)
    send = np.full((b, E), pad, np.int32)

    fa = state2feat(cfg, agent)
    fg = state2feat(cfg, goal)
    px, py = agent[..., 0], agent[..., 1]
    R = F(cfg.comm_radius)
    ids = np.arange(n, dtype=np.int32)

    # agent-agent block, receiver-major (EdgeBlock.make_edges, graph.py:35-44)
    aa = (fa[:, :, None, :] - fa[:, None, :, :]).astype(F)
    dist = norm2((px[:, :, None] - px[:, None, :]).astype(F), (py[:, :, None] - py[:, None, :]).astype(F))
    dist = (dist + (np.eye(n, dtype=F) * F(cfg.comm_radius + 1)).astype(F)[None]).astype(F)
    m = dist < R
    edges[:, :n * n] = aa.reshape(b, n * n, 4)
    recv[:, :n * n] = np.where(m, ids[None, :, None], pad).reshape(b, n * n)
    send[:, :n * n] = np.where(m, ids[None, None, :], pad).reshape(b, n * n)
    off = n * n

    # agent-goal block(s): mask is all ones
    if cfg.n_ag == g:
        ag = (fa[:, :, None, :] - fg[:, None, :, :]).astype(F)
        edges[:, off:off + n * g] = ag.reshape(b, n * g, 4)
        recv[:, off:off + n * g] = np.broadcast_to(ids[None, :, None], (b, n, g)).reshape(b, n * g)
        send[:, off:off + n * g] = np.broadcast_to((n + np.arange(g, dtype=np.int32))[None, None, :],
                                                   (b, n, g)).reshape(b, n * g)
        off += n * g
    else:
        edges[:, off:off + n] = (fa - fg).astype(F)
        recv[:, off:off + n] = ids[None]
        send[:, off:off + n] = (n + ids)[None]
        off += n

    # agent-obstacle blocks
    if n_on > 0:
        if cfg.is_lidar:
            k = cfg.top_k
            fx = (px[:, :, None] - obs_nodes[..., 0]).astype(F)
            fy = (py[:, :, None] - obs_nodes[..., 1]).astype(F)
            active = norm2(fx, fy) < F(cfg.comm_radius - 1e-1)
            ao = np.zeros((b, n, k, 4), F)
            ao[..., 0] = fx
            ao[..., 1] = fy
            edges[:, off:off + n * k] = ao.reshape(b, n * k, 4)
            sid = (n + g + np.arange(n * k, dtype=np.int32)).reshape(n, k)
            recv[:, off:off + n * k] = np.where(active, ids[None, :, None], pad).reshape(b, n * k)
            send[:, off:off + n * k] = np.where(active, sid[None], pad).reshape(b, n * k)
        else:
            o = cfg.n_obs
            ao = (agent[:, :, None, :] - obs_nodes[:, None, :, :]).astype(F)
            d = norm2((px[:, :, None] - obs_nodes[:, None, :, 0]).astype(F),
                      (py[:, :, None] - obs_nodes[:, None, :, 1]).astype(F))
            # within comm_radius (mpe_spread.py:73-75); always on in the corridor (x100: mpe_corridor.py:93)
            # (also x100 in the connect-spread env: mpe_connect_spread.py:168)
            active = d < (F(cfg.comm_radius * 100) if cfg.kind in (MPE_CORRIDOR, MPE_CONNECT_SPREAD) else R)
            edges[:, off:off + n * o] = ao.reshape(b, n * o, 4)
            sid = n + g + np.arange(o, dtype=np.int32)
            recv[:, off:off + n * o] = np.where(active, ids[None, :, None], pad).reshape(b, n * o)
            send[:, off:off + n * o] = np.where(active, sid[None, None, :], pad).reshape(b, n * o)

    return dict(n_node=np.full((b,), N, np.int32), n_edge=np.full((b,), E, np.int32),
                nodes=nodes, edges=edges, states=states, receivers=recv, senders=send,
                node_type=node_type)


def graph_slices(cfg: EnvCfg, graph: Dict[str, np.ndarray]):
    """type_states(0/1/2) (utils/graph.py:129-141) == static row slices for
    these envs.  -> agent, goal, obs_nodes (in get_graph's input format)."""
    n, g = cfg.n, cfg.n_goal
    st = graph["states"]
    agent = st[:, :n]
    goal = st[:, n:n + g]
    obs_nodes = None
    if cfg.n_obs_nodes > 0:
        o = st[:, n + g:n + g + cfg.n_obs_nodes]
        obs_nodes = o[..., :2].reshape(st.shape[0], n, cfg.top_k, 2) if cfg.is_lidar else o
    return agent, goal, obs_nodes


def env_step(cfg: EnvCfg, graph: Dict[str, np.ndarray], action: np.ndarray,
             obstacles: Optional[Dict[str, np.ndarray]] = None,
             rays: Optional[np.ndarray] = None):
    """LidarEnv.step / MPE.step (lidar_env/base.py:151-174, mpe/base.py:137-158).
    Reward and cost are evaluated on the PRE-step graph; LiDAR on the next
    agent states.  -> next_graph, reward (b,), cost (b,n,2), done (b,) False."""
    agent, goal, obs_nodes = graph_slices(cfg, graph)
    a = clip_action(action)
    nxt = agent_step_euler(cfg, agent, a)
    if cfg.is_lidar:
        nxt_obs = lidar_hits(cfg, nxt[..., :2], obstacles, rays) if cfg.n_obs > 0 else None
    else:
        nxt_obs = obs_nodes
    reward = get_reward(cfg, agent, goal, a)
    cost = get_cost(cfg, agent, obs_nodes)
    done = np.zeros(agent.shape[0], bool)
    return get_graph(cfg, nxt, goal, nxt_obs), reward, cost, done


# -------------------------------------------------------- synthetic states
def synthetic_states(cfg: EnvCfg, b: int, seed: int = 0):
    """Synthetic batched env states of the shapes the reference's reset
    produces (lidar_env/base.py:89-124, mpe/base.py:81-127,
    lidar_bicycle_target.py:60-90) WITHOUT the rejection sampling
    (SURVEY.md 8d).  -> agent, goal, obstacles(dict | None), mpe_obs | None"""
    rng = np.random.default_rng(seed)
    n, A = cfg.n, cfg.area
    pos = rng.uniform(0, A, (b, n, 2)).astype(F)
    if cfg.is_bicycle:
        th = rng.uniform(0, 2 * np.pi, (b, n)).astype(F)
        v = rng.uniform(-0.5, 0.5, (b, n, 1)).astype(F)
        agent = np.concatenate([pos, np.cos(th)[..., None], np.sin(th)[..., None], v], axis=-1).astype(F)
    else:
        vmax = 1.0 if not cfg.is_lidar else 0.5
        vel = rng.uniform(-vmax, vmax, (b, n, 2)).astype(F)
        agent = np.concatenate([pos, vel], axis=-1).astype(F)
    goal = np.zeros((b, n, cfg.state_dim), F)
    goal[..., :2] = rng.uniform(0, A, (b, n, 2)).astype(F)
    obstacles, mpe_obs = None, None
    if cfg.n_obs > 0:
        if cfg.is_lidar:
            c = rng.uniform(0, A, (b, cfg.n_obs, 2)).astype(F)
            wh = rng.uniform(0.1, 0.3, (b, cfg.n_obs, 2)).astype(F)
            th = rng.uniform(0, 2 * np.pi, (b, cfg.n_obs)).astype(F)
            obstacles = rect_create(c, wh[..., 0], wh[..., 1], th)
        else:
            mpe_obs = np.zeros((b, cfg.n_obs, 4), F)
            mpe_obs[..., :2] = rng.uniform(0.15, A - 0.15, (b, cfg.n_obs, 2)).astype(F)
    return agent, goal, obstacles, mpe_obs


def reset_graph(cfg: EnvCfg, agent, goal, obstacles, mpe_obs, rays=None):
    """Tail of reset: lidar on the initial states, then get_graph
    (lidar_env/base.py:121-124, mpe/base.py:125-127)."""
    if cfg.is_lidar:
        hits = lidar_hits(cfg, agent[..., :2], obstacles, rays) if cfg.n_obs > 0 else None
        return get_graph(cfg, agent, goal, hits)
    return get_graph(cfg, agent, goal, mpe_obs)
