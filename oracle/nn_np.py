"""NumPy restatement of the reference networks on the rollout hot path.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Citations are relative to
/root/reference/.  Parameters are the reference's flax pytrees (nested dicts
of arrays, names as flax auto-generates them - SURVEY.md appendix A.4); the
third-party layer arithmetic (flax Dense / LayerNorm / GRUCell, jraph segment
ops, tfp Normal / Tanh) is restated from the libraries' published behaviour:
PARITY UNPINNED against a real JAX install.

``dtype`` selects fp32 (the reference's arithmetic) or fp64 (an error-free
yardstick used by the tests to bound fp32 rounding noise).
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import numpy as np
from scipy import special as sps

F = np.float32
N_HEADS = 3


def _c(x, dt):
    return np.asarray(x, dt)


def dense(x, p, dt=F):
    """flax nn.Dense: y = x @ kernel (+ bias); kernel is (in, out)."""
    y = x @ _c(p["kernel"], dt)
    if "bias" in p:
        y = y + _c(p["bias"], dt)
    return y.astype(dt)


def layer_norm(x, p, dt=F, eps=1e-6):
    """flax nn.LayerNorm defaults: eps 1e-6, use_fast_variance (E[x^2]-E[x]^2,
    clipped at 0), scale and bias (dgppo/nn/mlp.py:28)."""
    mean = x.mean(-1, keepdims=True, dtype=dt)
    mean2 = (x * x).mean(-1, keepdims=True, dtype=dt)
    var = np.maximum(dt(0), mean2 - mean * mean)
    mul = (dt(1) / np.sqrt(var + dt(eps))) * _c(p["scale"], dt)
    return ((x - mean) * mul + _c(p["bias"], dt)).astype(dt)


def mlp_head(x, p, dt=F):
    """MLP(hid_sizes=(64,64), act_final=True, layernorm) (dgppo/nn/mlp.py:14-30)."""
    for i in range(2):
        x = dense(x, p[f"Dense_{i}"], dt)
        x = layer_norm(x, p[f"LayerNorm_{i}"], dt)
        x = np.maximum(x, dt(0))
    return x


def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def gru_cell(p, h, x, dt=F):
    """flax nn.GRUCell (dgppo/nn/rnn.py:19-21): biases on ir/iz/in and hn only."""
    r = sigmoid(dense(x, p["ir"], dt) + dense(h, p["hr"], dt)).astype(dt)
    z = sigmoid(dense(x, p["iz"], dt) + dense(h, p["hz"], dt)).astype(dt)
    n = np.tanh(dense(x, p["in"], dt) + r * dense(h, p["hn"], dt)).astype(dt)
    return ((dt(1) - z) * n + z * h).astype(dt)


def _gru_params(p_rnn):
    """RNN_0 holds exactly one GRUCell_* child (its auto-name index depends on
    how many cells RNN.__call__ instantiated: dgppo/nn/rnn.py:19-22)."""
    (cell,) = list(p_rnn.values())
    return cell


def graph_transformer(p, nodes, edges, recv, send, out_dim, dt=F):
    """GraphTransformer + GNNUpdate (dgppo/nn/gnn.py:22-41,78-117), batched.
    nodes (b,N,in), edges (b,E,ed), recv/send (b,E) -> new nodes (b,N,out)."""
    b, N, _ = nodes.shape
    E = edges.shape[1]
    H, d = N_HEADS, out_dim
    bi = np.arange(b)[:, None]
    x_s = nodes[bi, send]
    x_r = nodes[bi, recv]
    q = dense(x_r, p["Dense_0"], dt).reshape(b, E, H, d)
    k = dense(x_s, p["Dense_1"], dt).reshape(b, E, H, d)
    v = dense(x_s, p["Dense_2"], dt).reshape(b, E, H, d)
    e = dense(edges.astype(dt), p["Dense_3"], dt).reshape(b, E, H, d)
    attn = ((q * k).sum(-1, dtype=dt) / np.sqrt(dt(d))).astype(dt)      # (b,E,H)
    # jraph.segment_softmax over receivers, num_segments = N (pad is a segment)
    seg = (recv + bi * N).reshape(-1)
    a2 = attn.reshape(b * E, H)
    smax = np.full((b * N, H), -np.inf, dt)
    np.maximum.at(smax, seg, a2)
    ex = np.exp(a2 - smax[seg]).astype(dt)
    ssum = np.zeros((b * N, H), dt)
    np.add.at(ssum, seg, ex)
    sm = (ex / ssum[seg]).astype(dt).reshape(b, E, H, 1)
    msgs = ((sm * (v + e)).astype(dt)).mean(axis=2, dtype=dt)            # mean over heads
    agg = np.zeros((b * N, d), dt)
    np.add.at(agg, seg, msgs.reshape(b * E, d))
    feats = dense(nodes, p["Dense_4"], dt)
    return np.maximum(feats + agg.reshape(b, N, d), dt(0)).astype(dt)


def gnn(p, graph, n_layers, dt=F, msg_dim=32, out_dim=64):
    """GraphTransformerGNN (dgppo/nn/gnn.py:127-142): edges stay the env's."""
    x = graph["nodes"].astype(dt)
    for i in range(n_layers):
        od = out_dim if i == n_layers - 1 else msg_dim
        x = graph_transformer(p[f"GraphTransformer_{i}"], x, graph["edges"],
                              graph["receivers"], graph["senders"], od, dt)
    return x


def softplus(x):
    return np.logaddexp(x, 0)


# ----------------------------------------------------------- tanh-Normal
def _ndtr(x, dt):
    hs2 = dt(0.5 * np.sqrt(2.0))
    w = x * hs2
    z = np.abs(w)
    y = np.where(z < hs2, 1 + sps.erf(w), np.where(w > 0, 2 - sps.erfc(z), sps.erfc(z)))
    return (dt(0.5) * y).astype(dt)


def log_ndtr(x, dt=F):
    """tfp special_math.log_ndtr with the dtype's segment bounds and the
    3-term asymptotic series for the lower tail."""
    lower, upper = (-10.0, 5.0) if dt == F else (-20.0, 8.0)
    x = np.asarray(x, dt)
    with np.errstate(all="ignore"):
        xl = np.minimum(x, dt(lower))
        x2 = xl * xl
        series = 1.0 - 1.0 / x2 + 3.0 / (x2 * x2) - 15.0 / (x2 * x2 * x2)
        low = (-0.5 * x2 - np.log(-xl) - 0.5 * np.log(2.0 * np.pi) + np.log(series)).astype(dt)
        mid = np.log(_ndtr(np.maximum(x, dt(lower)), dt)).astype(dt)
        up = (-_ndtr(-x, dt)).astype(dt)
    return np.where(x > upper, up, np.where(x > lower, mid, low)).astype(dt)


def normal_log_prob(x, loc, scale, dt=F):
    """tfd.Normal._log_prob."""
    d = x / scale - loc / scale
    return (dt(-0.5) * d * d - (dt(0.5 * np.log(2.0 * np.pi)) + np.log(scale))).astype(dt)


def tanh_normal_log_prob(value, loc, scale, dt=F, threshold=0.999):
    """TanhTransformedDistribution.log_prob (dgppo/algo/module/distribution.py:25-35)
    wrapped in tfd.Independent(…, 1): sum over the action axis."""
    thr = dt(threshold)
    inv_thr = np.arctanh(thr).astype(dt)
    log_eps = dt(np.log(1.0 - threshold))
    lp_left = (log_ndtr((-inv_thr - loc) / scale, dt) - log_eps).astype(dt)
    lp_right = (log_ndtr(-((inv_thr - loc) / scale), dt) - log_eps).astype(dt)
    v = np.clip(value, -thr, thr).astype(dt)
    x = np.arctanh(v).astype(dt)
    fldj = dt(2.0) * (dt(np.log(2.0)) - x - softplus(dt(-2.0) * x))
    inner = (normal_log_prob(x, loc, scale, dt) - fldj).astype(dt)
    lp = np.where(v <= -thr, lp_left, np.where(v >= thr, lp_right, inner)).astype(dt)
    return lp.sum(-1, dtype=dt)


STD_DEV_INIT_INV = float(np.log(np.exp(0.5) - 1.0))   # TanhNormal.std_dev_init_inv (policy.py:54-59)
STD_DEV_MIN = 1e-5


def policy_forward(params, graph, rnn_state, n_agents, eps=None, n_layers=2, dt=F):
    """PPOPolicy.get_action / sample_action (dgppo/algo/module/policy.py:20-33,
    61-74,191-203).  graph fields batched (b,...); rnn_state (b,n,64);
    eps None -> mode tanh(mean) (distribution.py:45-46), else the N(0,1) draw
    (b,n,nu) used for tanh(mean + std*eps).
    -> action (b,n,nu), log_pi (b,n) | None, new rnn_state (b,n,64), (mean,std)"""
    p = params["params"]
    base = p["PolicyNet_0"]
    x = gnn(base["GraphTransformerGNN_0"], graph, n_layers, dt)[:, :n_agents]     # type_nodes(0, n)
    x = mlp_head(x, base["PolicyGNNHead"], dt)
    h = gru_cell(_gru_params(base["RNN_0"]), rnn_state.astype(dt), x, dt)
    f = dense(h, p["ScaleHid"], dt)
    mean = dense(f, p["OutputDenseMean"], dt)
    std = (softplus(dense(f, p["OutputDenseStdTrans"], dt) + dt(STD_DEV_INIT_INV)) + dt(STD_DEV_MIN)).astype(dt)
    if eps is None:
        return np.tanh(mean).astype(dt), None, h, (mean, std)
    action = np.tanh(mean + std * eps.astype(dt)).astype(dt)
    log_pi = tanh_normal_log_prob(action, mean, std, dt)
    return action, log_pi, h, (mean, std)


def vh_forward(params, graph, rnn_state, n_agents, n_layers=1, dt=F):
    """DGPPO.get_Vh -> DecRStateFn (dgppo/algo/dgppo.py:128-134,
    dgppo/algo/module/value.py:47-79).  The GRU carry is the POLICY's stored
    rnn state.  -> Vh (b,n,n_cost)"""
    p = params["params"]
    x = gnn(p["GraphTransformerGNN_0"], graph, n_layers, dt)[:, :n_agents]
    x = mlp_head(x, p["ValueGNNHead"], dt)
    h = gru_cell(_gru_params(p["RNN_0"]), rnn_state.astype(dt), x, dt)
    return dense(h, p["Dense_0"], dt)


def vl_forward(params, graph, rnn_state, n_agents, n_layers=2, dt=F):
    """ValueNet.get_value -> RStateFn (value.py:15-44): mean-pool over agents,
    one GRU row per env.  rnn_state (b,64) -> Vl (b,), new state (b,64)"""
    p = params["params"]
    x = gnn(p["GraphTransformerGNN_0"], graph, n_layers, dt)[:, :n_agents]
    x = x.mean(axis=1, keepdims=True, dtype=dt)
    x = mlp_head(x, p["ValueGNNHead"], dt)
    h = gru_cell(_gru_params(p["RNN_0"]), rnn_state.astype(dt)[:, None, :], x, dt)
    return dense(h, p["Dense_0"], dt)[:, 0, 0], h[:, 0]
