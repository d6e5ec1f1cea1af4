"""NumPy restatement of the DGPPO rollout / GAE / CBF-advantage pre-pass.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Citations are relative to
/root/reference/.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np

from . import env_np, nn_np

F = np.float32


def compute_dec_ocp_gae(Tah_hs, T_l, Tp1ah_Vh, Tp1_Vl, disc_gamma, gae_lambda, dt=F):
    """compute_dec_ocp_gae (dgppo/algo/utils.py:11-79), literal O(T^2) DP for
    ONE trajectory.  hs (T,a,nh), l (T,), Vh (T+1,a,nh), Vl (T+1,)
    -> Qh (T,a,nh), Ql (T,)"""
    Tah_hs = np.asarray(Tah_hs, dt)
    T_l = np.asarray(T_l, dt)
    Tp1ah_Vh = np.asarray(Tp1ah_Vh, dt)
    Tp1_Vl = np.asarray(Tp1_Vl, dt)
    T, a, nh = Tah_hs.shape
    g, lam = dt(disc_gamma), dt(gae_lambda)
    Vhs_row = np.zeros((T + 1, a, nh), dt)
    Vhs_row[0] = Tp1ah_Vh[-1]
    Vl_row = np.zeros((T + 1, a), dt)
    Vl_row[0] = Tp1_Vl[-1]
    coeffs = np.zeros(T + 1, dt)
    coeffs[0] = 1.0
    Qh = np.zeros((T, a, nh), dt)
    Ql = np.zeros((T,), dt)
    for ii in range(T):                       # reverse scan: ts[t] = T-1-t = ii
        t = T - 1 - ii
        hs, l = Tah_hs[t], T_l[t]
        mask = (np.arange(T + 1) < ii + 1).astype(dt)
        h_disc = hs.max(-1)
        disc_to_h = ((dt(1) - g) * h_disc[None, :, None]).astype(dt) + (g * Vhs_row).astype(dt)
        Vhs_row = (mask[:, None, None] * np.maximum(hs[None], disc_to_h)).astype(dt)
        Vl_row = (mask[:, None] * (l + (g * Vl_row).astype(dt))).astype(dt)
        cat = np.concatenate([Vhs_row, Vl_row[:, :, None]], axis=-1)
        Q = np.einsum("kah,k->ah", cat, coeffs).astype(dt)
        Qh[t] = Q[:, :nh]
        Ql[t] = Q[0, nh]
        Vhs_row[ii + 1] = Tp1ah_Vh[t]
        Vl_row[ii + 1] = Tp1_Vl[t]
        coeffs = np.roll(coeffs, 1)
        coeffs[0] = np.power(lam, dt(ii + 1))
        coeffs[1] = np.power(lam, dt(ii)) * (dt(1) - lam)
    return Qh, Ql


def gae_closed_form(Tah_hs, T_l, Tp1ah_Vh, Tp1_Vl, gamma, lam):
    """fp64 closed forms of the same quantities:
    Ql = lambda-return, Qh = lambda-mix of n-step max-returns."""
    hs = np.asarray(Tah_hs, np.float64)
    l = np.asarray(T_l, np.float64)
    Vh = np.asarray(Tp1ah_Vh, np.float64)
    Vl = np.asarray(Tp1_Vl, np.float64)
    T = hs.shape[0]
    Ql = np.zeros(T)
    G = None
    for t in range(T - 1, -1, -1):
        G = l[t] + gamma * Vl[t + 1] if t == T - 1 else l[t] + gamma * ((1 - lam) * Vl[t + 1] + lam * G)
        Ql[t] = G
    Qh = np.zeros_like(hs)
    hmax = hs.max(-1, keepdims=True)
    # M[t][n]: n-step max-return from t bootstrapped with Vh[t+n]
    M_next = {0: Vh[T]}
    for t in range(T - 1, -1, -1):
        M = {0: Vh[t]}
        for n_ in range(1, T - t + 1):
            M[n_] = np.maximum(hs[t], (1 - gamma) * hmax[t] + gamma * M_next[n_ - 1])
        K = T - t
        q = lam ** (K - 1) * M[K]
        for n_ in range(1, K):
            q = q + (1 - lam) * lam ** (n_ - 1) * M[n_]
        Qh[t] = q
        M_next = M
    return Qh, Ql


def cbf_advantage(bT_Ql, bTp1_Vl, bTp1ah_Vh, dt_env, alpha, cbf_eps, cbf_weight, dt=F):
    """Advantage merge of DGPPO.update_inner (dgppo/algo/dgppo.py:239-259).
    -> A (b,T,a), cbf_deriv (b,T,a,nh), Acbf (b,T,a,nh), is_safe (b,T,a)"""
    Ql = np.asarray(bT_Ql, dt)
    Vl = np.asarray(bTp1_Vl, dt)[:, :-1]
    Vh = np.asarray(bTp1ah_Vh, dt)
    Al = (Ql - Vl).astype(dt)
    mean = Al.mean(axis=1, keepdims=True, dtype=dt)
    std = np.sqrt(((Al - mean) ** 2).mean(axis=1, keepdims=True, dtype=dt)).astype(dt)
    Al = ((Al - mean) / (std + dt(1e-8))).astype(dt)
    Vh_t = Vh[:, :-1]
    deriv = ((Vh[:, 1:] - Vh_t) / dt(dt_env) + dt(alpha) * Vh_t).astype(dt)
    Acbf = np.maximum(deriv + dt(cbf_eps), dt(0)).astype(dt)
    is_safe = (deriv <= 0).min(axis=-1)
    A = np.where(is_safe, Al[:, :, None], dt(0)).astype(dt)
    A = (A + Acbf.max(axis=-1) * dt(cbf_weight)).astype(dt)
    return (-A).astype(dt), deriv, Acbf, is_safe


def rollout(cfg: env_np.EnvCfg, policy_params, graph0: Dict[str, np.ndarray], obstacles,
            eps: Optional[np.ndarray], T: int, n_layers: int = 2, rays=None, dt=F):
    """rollout / test_rollout scan body (dgppo/trainer/utils.py:45-57,70-86)
    from an already-reset batch of graphs.  eps (b,T,n,nu) N(0,1) draws for the
    stochastic rollout, None for the deterministic one (actor = algo.act).
    Returns per-field arrays stacked over T+1 graphs (graph = [:, :T],
    next_graph = [:, 1:]), rnn (b,T+1,n,64) (rollout: [:, :T]; test_rollout
    stores the post-step carry: [:, 1:]), actions, rewards, costs, log_pis."""
    b = graph0["nodes"].shape[0]
    n = cfg.n
    graphs = {k: [v] for k, v in graph0.items()}
    rnn = [np.zeros((b, n, 64), F)]
    actions, rewards, costs, log_pis = [], [], [], []
    g = graph0
    for t in range(T):
        e = None if eps is None else eps[:, t]
        a, lp, h, _ = nn_np.policy_forward(policy_params, g, rnn[-1], n, e, n_layers, dt)
        a = a.astype(F)
        g, r, c, _ = env_np.env_step(cfg, g, a, obstacles, rays)
        for k, v in g.items():
            graphs[k].append(v)
        rnn.append(h.astype(F))
        actions.append(a)
        rewards.append(r)
        costs.append(c)
        if lp is not None:
            log_pis.append(lp.astype(F))
    out = {k: np.stack(v, axis=1) for k, v in graphs.items()}
    out.update(rnn_states=np.stack(rnn, axis=1), actions=np.stack(actions, axis=1),
               rewards=np.stack(rewards, axis=1), costs=np.stack(costs, axis=1),
               dones=np.zeros((b, T), bool),
               log_pis=np.stack(log_pis, axis=1) if log_pis else None)
    return out
