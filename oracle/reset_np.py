"""NumPy restatement of the reset sampler: obstacle sampling + get_node_goal_rng
(dgppo/env/utils.py:139-244, lidar_env/base.py:89-119, mpe/base.py:81-125,
lidar_bicycle_target.py:60-85) on the counter-based hash stream that
dgppo_reset uses (include/dgppo_abi.h, K0).  TEST INFRASTRUCTURE: a plain
per-environment Python loop, for small batches.
"""
from __future__ import annotations

import numpy as np

from . import env_np

F = np.float32
MASK = (1 << 64) - 1
MAX_ITER = 1024
MAX_RESTARTS = 16          # kernel guard (csrc/reset_kernels.cu): the reference itself would loop for ever


def splitmix64(x: int) -> int:
    x &= MASK
    x ^= x >> 30; x = (x * 0xBF58476D1CE4E5B9) & MASK
    x ^= x >> 27; x = (x * 0x94D049BB133111EB) & MASK
    x ^= x >> 31
    return x


class Rng:
    def __init__(self, key: int):
        self.key, self.ctr = int(key) & MASK, 0

    def next2(self):
        x = splitmix64(self.key + self.ctr * 0x9E3779B97F4A7C15)
        self.ctr += 1
        s = F(1.0 / 16777216.0)
        return F(x >> 40) * s, F((x >> 8) & 0xFFFFFF) * s


def _inside_any(p, rec, r):
    if rec is None or len(rec) == 0:
        return False
    obs = dict(center=rec[None, :, 0:2], width=rec[None, :, 2], height=rec[None, :, 3], cos=rec[None, :, 5],
               sin=rec[None, :, 6])
    return bool(env_np.rect_inside(np.asarray(p, F)[None, None, :], obs, r)[0, 0])


def _collides(c, pts, min_dist):
    d = env_np.norm2((pts[:, 0] - c[0]).astype(F), (pts[:, 1] - c[1]).astype(F))
    return bool((d <= F(min_dist)).any())


def reset_states(cfg: env_np.EnvCfg, key: int, obs_len=(0.1, 0.3), theta_range=None):
    """-> agent (n, sd), goal (n, sd), obstacle record (n_obs, 16) | mpe obs (n_obs, 4) | None, n_draws"""
    rng = Rng(key)
    n, A = cfg.n, F(cfg.area)
    lid = cfg.is_lidar
    rec = None
    if lid and cfg.n_obs > 0:
        th_lo, th_hi = theta_range if theta_range is not None else \
            ((-np.pi, np.pi) if cfg.is_bicycle else (0.0, 2 * np.pi))
        lo, hi, tl, thh = F(obs_len[0]), F(obs_len[1]), F(th_lo), F(th_hi)
        rec = np.zeros((cfg.n_obs, 16), F)
        for o in range(cfg.n_obs):
            c, l, t = rng.next2(), rng.next2(), rng.next2()
            cx, cy = F(c[0] * A), F(c[1] * A)
            w = F(lo + F(l[0] * F(hi - lo)))
            h = F(lo + F(l[1] * F(hi - lo)))
            th = F(tl + F(t[0] * F(thh - tl)))
            r = env_np.rect_create(np.array([cx, cy], F), w, h, th)
            rec[o, 0:2] = [cx, cy]
            rec[o, 2:7] = [w, h, th, r["cos"], r["sin"]]
            rec[o, 8:16] = r["points"].reshape(8)
    min_dist = F((2.2 if lid else 2.0) * cfg.car_radius)
    half = F(min_dist / F(2))
    st, gl = np.zeros((n, 2), F), np.zeros((n, 2), F)
    corridor = cfg.kind == env_np.MPE_CORRIDOR
    Ay = A
    if corridor:     # mpe_corridor.py:39-50: side_length_y of the sampler, Python doubles rounded once
        Ay = F((cfg.area - cfg.obs_radius * 2) / 2 - 1.5 * cfg.car_radius)
    agent_id = restarts = 0
    while agent_id < n:
        u = rng.next2()
        c = (F(u[0] * A), F(u[1] * Ay))
        it_a = 0
        while it_a < MAX_ITER and (_collides(c, st, min_dist) or _inside_any(c, rec, half)):
            it_a += 1
            u = rng.next2()
            c = (F(u[0] * A), F(u[1] * Ay))
        st[agent_id] = c
        u = rng.next2()
        g = (F(u[0] * A), F(u[1] * Ay))
        it_g = 0
        while it_g < MAX_ITER and (_collides(g, gl, min_dist) or _inside_any(g, rec, half)
                                   or g[0] < 0 or g[1] < 0 or g[0] > A or g[1] > A):
            it_g += 1
            u = rng.next2()
            g = (F(u[0] * A), F(u[1] * Ay))
        gl[agent_id] = g
        agent_id += 1
        if it_a >= MAX_ITER or it_g >= MAX_ITER:
            restarts += 1
            if restarts > MAX_RESTARTS:
                continue
            agent_id = 0
            st[:] = 0
            gl[:] = 0
    obst = rec
    if corridor:     # goals shifted past the corridor, two fixed obstacles (mpe_corridor.py:50-54)
        gl[:, 1] = (gl[:, 1] + F(cfg.area - (cfg.area - cfg.obs_radius * 2) / 2 + 1.5 * cfg.car_radius)).astype(F)
        obst = np.zeros((2, 4), F)
        obst[0, :2] = (F(cfg.obs_radius), F(cfg.area / 2))
        obst[1, :2] = (F(cfg.area - cfg.obs_radius), F(cfg.area / 2))
    elif not lid and cfg.n_obs > 0:
        obst = np.zeros((cfg.n_obs, 4), F)
        car, obr = F(cfg.car_radius), F(cfg.obs_radius)
        lo, hi = F(car * F(3)), F(A - F(car * F(3)))
        for o in range(cfg.n_obs):
            u = rng.next2()
            p = (F(u[0] * A), F(u[1] * A))
            while (_collides(p, st, F(car + obr)) or _collides(p, gl, F(F(car * F(2)) + obr))
                   or p[0] < lo or p[1] < lo or p[0] > hi or p[1] > hi):
                u = rng.next2()
                p = (F(lo + F(u[0] * F(hi - lo))), F(lo + F(u[1] * F(hi - lo))))
            obst[o, 0:2] = p
    sd = cfg.state_dim
    agent, goal = np.zeros((n, sd), F), np.zeros((n, sd), F)
    agent[:, :2], goal[:, :2] = st, gl
    if cfg.is_bicycle:
        for i in range(n):
            u = rng.next2()
            th = F(u[0] * F(6.283185307179586))
            agent[i, 2], agent[i, 3] = np.cos(th), np.sin(th)
    return agent, goal, obst, rng.ctr


# --------------------------------------------------------------------------------------------------
# Landmark families + connected spread (env kinds 6-9): LidarLine.reset (lidar_line.py:38-126), MPELine.reset
# (mpe_line.py:38-117), MPEFormation.reset (mpe_formation.py:38-91), MPEConnectSpread.reset
# (mpe_connect_spread.py:52-103), on the counter stream and in the draw order of reset_landmark_kernel.
def _sample_nodes(rng, n, min_dist, ax, ay, area):
    """get_node_goal_rng without obstacles -> st, gl, ok"""
    st, gl = np.zeros((n, 2), F), np.zeros((n, 2), F)
    agent_id = restarts = 0
    while agent_id < n:
        u = rng.next2()
        c = (F(u[0] * ax), F(u[1] * ay))
        it_a = 0
        while it_a < MAX_ITER and _collides(c, st, min_dist):
            it_a += 1
            u = rng.next2()
            c = (F(u[0] * ax), F(u[1] * ay))
        st[agent_id] = c
        u = rng.next2()
        g = (F(u[0] * ax), F(u[1] * ay))
        it_g = 0
        while it_g < MAX_ITER and (_collides(g, gl, min_dist) or g[0] < 0 or g[1] < 0 or g[0] > area or g[1] > area):
            it_g += 1
            u = rng.next2()
            g = (F(u[0] * ax), F(u[1] * ay))
        gl[agent_id] = g
        agent_id += 1
        if it_a >= MAX_ITER or it_g >= MAX_ITER:
            restarts += 1
            if restarts > MAX_RESTARTS:
                continue
            agent_id = 0
            st[:] = 0
            gl[:] = 0
    return st, gl, restarts <= MAX_RESTARTS


def _badly_spaced(p, lo, hi):
    n = len(p)
    d = env_np.norm2((p[:, None, 0] - p[None, :, 0]).astype(F), (p[:, None, 1] - p[None, :, 1]).astype(F))
    d = (d + (np.eye(n, dtype=F) * F(1e6)).astype(F)).astype(F)
    m = d.min(axis=1)
    return bool(((m > F(hi)) | (m < F(lo))).any())


def reset_landmark_states(cfg: env_np.EnvCfg, key: int, obs_len=(0.1, 0.3)):
    """-> agent (n, 4), goal nodes (n_goal, 4), obstacle record (n_obs, 16) | mpe obs (n_obs, 4), n_draws (-1: gave up)"""
    rng = Rng(key)
    n, A = cfg.n, F(cfg.area)
    car, obr = F(cfg.car_radius), F(cfg.obs_radius)
    ok = True
    agent = np.zeros((n, 4), F)
    if cfg.kind == env_np.MPE_CONNECT_SPREAD:
        side_y = F((cfg.area - cfg.obs_radius * 2) / 2 - 1.5 * cfg.car_radius)
        shift = F(cfg.area - (cfg.area - cfg.obs_radius * 2) / 2 + 1.5 * cfg.car_radius)
        tries = 0
        while True:
            st, gl, ok = _sample_nodes(rng, n, F(2.3 * cfg.car_radius), A, side_y, A)
            gl[:, 1] = (gl[:, 1] + shift).astype(F)
            tries += 1
            if not (ok and tries < 4096 and (_badly_spaced(st, F(car * F(2)), cfg.connect_radius)
                                             or _badly_spaced(gl, 0.0, cfg.connect_radius))):
                break
        if tries >= 4096:
            ok = False
        u = rng.next2()
        obst = np.zeros((1, 4), F)
        obst[0, 0] = F(F(u[0] * F(F(A - obr) - obr)) + obr)
        obst[0, 1] = F(A / F(2))
        goal = np.zeros((n, 4), F)
        agent[:, :2], goal[:, :2] = st, gl
        return agent, goal, obst, (rng.ctr if ok else -1)

    st, _, ok = _sample_nodes(rng, n, F(2.0 * cfg.car_radius), A, A, A)
    short_line = cfg.kind == env_np.MPE_LINE and n <= 3
    if cfg.kind == env_np.MPE_FORMATION:
        lo = F(cfg.comm_radius + 2 * cfg.car_radius)
        hi = F(cfg.area - cfg.comm_radius - 2 * cfg.car_radius)
        u = rng.next2()
        lm = np.array([[F(F(u[0] * F(hi - lo)) + lo), F(F(u[1] * F(hi - lo)) + lo)]], F)
    else:
        lm_min = (n * 5 * cfg.car_radius) if short_line else ((n - 2) * 6 * cfg.car_radius)
        u = rng.next2()
        if short_line:
            l0 = (F(u[0] * A), F(u[1] * A))
        else:
            side = F(cfg.area - lm_min)
            half = F(A / F(2))
            cx = F(F(F(u[0] * F(A - side)) - half) + F(0))
            cy = F(F(F(u[1] * side) - F(0)) + F(half - side))
            r = rng.next2()
            region = min(3, int(F(r[0] * F(4))))
            rx, ry = [(cx, cy), (-cy, cx), (-cx, -cy), (cy, -cx)][region]
            l0 = (F(rx + half), F(ry + half))
        u = rng.next2()
        l1 = (F(u[0] * A), F(u[1] * A))
        guard = 0
        while guard < (1 << 16) and env_np.norm2(F(l1[0] - l0[0]), F(l1[1] - l0[1])) < F(lm_min):
            guard += 1
            u = rng.next2()
            l1 = (F(u[0] * A), F(u[1] * A))
        if guard >= (1 << 16):
            ok = False
        lm = np.array([l0, l1], F)
    eg = env_np.landmark2goal(cfg, lm[None])[0]
    if cfg.kind == env_np.LIDAR_LINE:
        r = F(car * F(1.1))
        lo_, hi_ = F(obs_len[0]), F(obs_len[1])
        obst = np.zeros((cfg.n_obs, 16), F)
        pts = np.concatenate([st, eg], axis=0)
        for o in range(cfg.n_obs):
            guard = 0
            while True:
                c, l, t = rng.next2(), rng.next2(), rng.next2()
                cx, cy = F(c[0] * A), F(c[1] * A)
                w = F(lo_ + F(l[0] * F(hi_ - lo_)))
                h = F(lo_ + F(l[1] * F(hi_ - lo_)))
                th = F(t[0] * F(3.14159274101257324))
                rc = env_np.rect_create(np.array([cx, cy], F), w, h, th)
                rec = np.zeros(16, F)
                rec[0:2] = [cx, cy]
                rec[2:7] = [w, h, th, rc["cos"], rc["sin"]]
                rec[8:16] = rc["points"].reshape(8)
                bad = any(_inside_any(p, rec[None], r) for p in pts)
                guard += 1
                if not (bad and guard < (1 << 16)):
                    break
            if guard >= (1 << 16):
                ok = False
            obst[o] = rec
    else:
        obst = np.zeros((cfg.n_obs, 4), F)
        lo, hi = F(car * F(3)), F(A - F(car * F(3)))
        for o in range(cfg.n_obs):
            u = rng.next2()
            p = (F(u[0] * A), F(u[1] * A))
            guard = 0
            while guard < (1 << 20) and (_collides(p, st, F(car + obr)) or _collides(p, eg, F(F(car * F(2)) + obr))
                                         or p[0] < lo or p[1] < lo or p[0] > hi or p[1] > hi):
                guard += 1
                u = rng.next2()
                p = (F(lo + F(u[0] * F(hi - lo))), F(lo + F(u[1] * F(hi - lo))))
            obst[o, 0:2] = p
    goal = np.zeros((cfg.n_goal, 4), F)
    agent[:, :2], goal[:, :2] = st, lm
    return agent, goal, obst, (rng.ctr if ok else -1)
