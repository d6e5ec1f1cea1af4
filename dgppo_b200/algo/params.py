"""Network parameters: reference-shaped pytrees, their initialisation, and the
packing into the device layout of ``DgppoNetLayout`` (include/dgppo_abi.h).

The pytrees mirror what flax builds for the reference modules (nested dicts,
names as flax auto-generates them: SURVEY.md appendix A.4;
dgppo/algo/module/policy.py:20-78, value.py:15-79, nn/gnn.py:78-142,
nn/mlp.py:14-30, nn/rnn.py:14-30), so a reference checkpoint (`actor.pkl`,
`Vl.pkl`, `Vh.pkl`: informarl_lagr.py:311-317) converted to NumPy can be
packed as is.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict

import numpy as np

from .. import _lib

N_HEADS, MSG_DIM, OUT_DIM, HID = 3, 32, 64, 64


def net_cfg(kind: int, node_dim: int, edge_dim: int, n_layers: int, n_out: int) -> _lib.DgppoNetCfg:
    return _lib.DgppoNetCfg(kind, node_dim, edge_dim, n_layers, n_out)


def net_layout(cfg: _lib.DgppoNetCfg) -> _lib.DgppoNetLayout:
    L = _lib.DgppoNetLayout()
    _lib.check(_lib.lib().dgppo_net_layout(C.byref(cfg), C.byref(L)), "dgppo_net_layout")
    return L


# ------------------------------------------------------------------- init
def _orthogonal(rng: np.random.Generator, shape, scale: float = 1.0) -> np.ndarray:
    """nn.initializers.orthogonal (dgppo/nn/utils.py:20): QR of a Gaussian."""
    n_rows, n_cols = shape
    a = rng.standard_normal((max(n_rows, n_cols), min(n_rows, n_cols)))
    qm, rm = np.linalg.qr(a)
    qm = qm * np.sign(np.diag(rm))
    if n_rows < n_cols:
        qm = qm.T
    return (scale * qm).astype(np.float32)


def _dense(rng, n_in, n_out, bias=True, scale=1.0, jitter=0.0):
    p = {"kernel": _orthogonal(rng, (n_in, n_out), scale)}
    if bias:
        p["bias"] = (jitter * rng.standard_normal(n_out)).astype(np.float32)
    return p


def _gnn(rng, node_dim, edge_dim, n_layers, jitter):
    out = {}
    in_dim = node_dim
    for i in range(n_layers):
        d = OUT_DIM if i == n_layers - 1 else MSG_DIM          # gnn.py:136
        hd = N_HEADS * d
        out[f"GraphTransformer_{i}"] = {
            "Dense_0": _dense(rng, in_dim, hd, jitter=jitter),  # query  (gnn.py:86)
            "Dense_1": _dense(rng, in_dim, hd, jitter=jitter),  # key    (gnn.py:89)
            "Dense_2": _dense(rng, in_dim, hd, jitter=jitter),  # value  (gnn.py:92)
            "Dense_3": _dense(rng, edge_dim, hd, bias=False),   # edge   (gnn.py:95)
            "Dense_4": _dense(rng, in_dim, d, jitter=jitter),   # update (gnn.py:110)
        }
        in_dim = d
    return out


def _head(rng, jitter):
    out = {}
    for i in range(2):
        out[f"Dense_{i}"] = _dense(rng, HID, HID, jitter=jitter)
        out[f"LayerNorm_{i}"] = {
            "scale": (1.0 + jitter * rng.standard_normal(HID)).astype(np.float32),
            "bias": (jitter * rng.standard_normal(HID)).astype(np.float32)}
    return out


def _gru(rng, jitter):
    cell = {}
    for k in ("ir", "iz", "in"):                     # lecun_normal kernels, zero biases
        cell[k] = {"kernel": (rng.standard_normal((HID, HID)) / np.sqrt(HID)).astype(np.float32),
                   "bias": (jitter * rng.standard_normal(HID)).astype(np.float32)}
    for k in ("hr", "hz"):
        cell[k] = {"kernel": _orthogonal(rng, (HID, HID))}
    cell["hn"] = {"kernel": _orthogonal(rng, (HID, HID)),
                  "bias": (jitter * rng.standard_normal(HID)).astype(np.float32)}
    return {"GRUCell_0": cell}


def init_policy_params(node_dim: int, edge_dim: int, action_dim: int, n_layers: int = 2,
                       seed: int = 0, jitter: float = 0.0, scale_final: float = 0.01) -> Dict:
    """TanhNormal(PolicyNet) parameter pytree (policy.py:61-74).  ``jitter``
    perturbs the (reference-zero) biases / LayerNorm affine so tests exercise
    every term."""
    rng = np.random.default_rng(seed)
    return {"params": {
        "PolicyNet_0": {
            "GraphTransformerGNN_0": _gnn(rng, node_dim, edge_dim, n_layers, jitter),
            "PolicyGNNHead": _head(rng, jitter),
            "RNN_0": _gru(rng, jitter)},
        "ScaleHid": _dense(rng, HID, HID, scale=scale_final, jitter=jitter),
        "OutputDenseMean": _dense(rng, HID, action_dim, jitter=jitter),
        "OutputDenseStdTrans": _dense(rng, HID, action_dim, jitter=jitter)}}


def init_value_params(node_dim: int, edge_dim: int, n_out: int, n_layers: int,
                      seed: int = 0, jitter: float = 0.0) -> Dict:
    """DecRStateFn / RStateFn parameter pytree (value.py:15-79)."""
    rng = np.random.default_rng(seed)
    return {"params": {
        "GraphTransformerGNN_0": _gnn(rng, node_dim, edge_dim, n_layers, jitter),
        "ValueGNNHead": _head(rng, jitter),
        "RNN_0": _gru(rng, jitter),
        "Dense_0": _dense(rng, HID, n_out, jitter=jitter)}}


# ------------------------------------------------------------------- pack
def _np(x) -> np.ndarray:
    return np.asarray(x, np.float32)


def pack_params(tree: Dict, cfg: _lib.DgppoNetCfg) -> np.ndarray:
    """Flatten a reference-shaped pytree into the packed fp32 device layout."""
    L = net_layout(cfg)
    buf = np.zeros(L.total, np.float32)
    p = tree["params"]
    policy = cfg.kind == _lib.NET_POLICY
    base = p["PolicyNet_0"] if policy else p
    gnn = base["GraphTransformerGNN_0"]

    def put(off, arr):
        arr = _np(arr).ravel()
        buf[off:off + arr.size] = arr

    Hh = N_HEADS
    for l in range(cfg.n_layers):
        g = gnn[f"GraphTransformer_{l}"]
        IN, D = L.in_dim[l], L.out_dim[l]
        INP, INA = (IN + 1 + 3) // 4 * 4, IN + 5
        wq, wk, wv = _np(g["Dense_0"]["kernel"]), _np(g["Dense_1"]["kernel"]), _np(g["Dense_2"]["kernel"])
        we, wu = _np(g["Dense_3"]["kernel"]), _np(g["Dense_4"]["kernel"])
        assert wq.shape == (IN, Hh * D) and we.shape == (cfg.edge_dim, Hh * D) and wu.shape == (IN, D)
        put(L.wq[l], wq)
        put(L.bq[l], g["Dense_0"]["bias"])
        wkt = np.zeros((Hh, D, INP), np.float32)
        wkt[:, :, :IN] = wk.reshape(IN, Hh, D).transpose(1, 2, 0)
        wkt[:, :, IN] = _np(g["Dense_1"]["bias"]).reshape(Hh, D)
        put(L.wkt[l], wkt)
        # merged query-key block: qt[h][c] = sum_j (x Wq_h + bq_h)[j] * wkt[h][j][c], linear in x -> one
        # (IN+1) x (H*INP) matrix (last row = the bq part); products in double, rounded once
        wq_aug = np.concatenate([wq, _np(g["Dense_0"]["bias"])[None]], axis=0).astype(np.float64)   # (IN+1, H*D)
        wqk = np.einsum("chj,hjk->chk", wq_aug.reshape(IN + 1, Hh, D), wkt.astype(np.float64))      # (IN+1, H, INP)
        put(L.wqk[l], wqk.reshape(IN + 1, Hh * INP).astype(np.float32))
        wagg = np.zeros((Hh, INA, D), np.float32)
        wagg[:, :IN, :] = wv.reshape(IN, Hh, D).transpose(1, 0, 2)
        wagg[:, IN, :] = _np(g["Dense_2"]["bias"]).reshape(Hh, D)
        wagg[:, IN + 1:, :] = we.reshape(cfg.edge_dim, Hh, D).transpose(1, 0, 2)
        put(L.wagg[l], wagg)
        put(L.wu[l], wu)
        put(L.bu[l], g["Dense_4"]["bias"])

    head = base["PolicyGNNHead" if policy else "ValueGNNHead"]
    put(L.d0w, head["Dense_0"]["kernel"]); put(L.d0b, head["Dense_0"]["bias"])
    put(L.ln0s, head["LayerNorm_0"]["scale"]); put(L.ln0b, head["LayerNorm_0"]["bias"])
    put(L.d1w, head["Dense_1"]["kernel"]); put(L.d1b, head["Dense_1"]["bias"])
    put(L.ln1s, head["LayerNorm_1"]["scale"]); put(L.ln1b, head["LayerNorm_1"]["bias"])

    (cell,) = list(base["RNN_0"].values())        # single GRUCell_* child (rnn.py:19-22)
    put(L.wi, np.concatenate([_np(cell[k]["kernel"]) for k in ("ir", "iz", "in")], axis=1))
    put(L.bi, np.concatenate([_np(cell[k]["bias"]) for k in ("ir", "iz", "in")]))
    put(L.wh, np.concatenate([_np(cell[k]["kernel"]) for k in ("hr", "hz", "hn")], axis=1))
    put(L.bhn, cell["hn"]["bias"])

    out_w = np.zeros((HID, 4), np.float32)
    out_b = np.zeros(4, np.float32)
    if policy:
        nu = cfg.n_out
        out_w[:, :nu] = _np(p["OutputDenseMean"]["kernel"]); out_b[:nu] = _np(p["OutputDenseMean"]["bias"])
        out_w[:, 2:2 + nu] = _np(p["OutputDenseStdTrans"]["kernel"])
        out_b[2:2 + nu] = _np(p["OutputDenseStdTrans"]["bias"])
        # ScaleHid is a Dense without activation in front of the two output Denses (policy.py:66-70):
        # fold it in, (x Ws + bs) Wo + bo = x (Ws Wo) + (bs Wo + bo); products in double, rounded once
        ws, bs = _np(p["ScaleHid"]["kernel"]).astype(np.float64), _np(p["ScaleHid"]["bias"]).astype(np.float64)
        out_b = (bs @ out_w.astype(np.float64) + out_b.astype(np.float64)).astype(np.float32)
        out_w = (ws @ out_w.astype(np.float64)).astype(np.float32)
    else:
        out_w[:, :cfg.n_out] = _np(p["Dense_0"]["kernel"]); out_b[:cfg.n_out] = _np(p["Dense_0"]["bias"])
    put(L.out_w, out_w); put(L.out_b, out_b)
    # tensor-core head operands (include/dgppo_abi.h, tc_head): [hi | lo] TF32 split of each matrix in the
    # K-major canonical layout [K/4][N][4]
    off = L.tc_head
    for w in (buf[L.d0w:L.d0w + HID * HID].reshape(HID, HID), buf[L.d1w:L.d1w + HID * HID].reshape(HID, HID),
              buf[L.wi:L.wi + HID * 3 * HID].reshape(HID, 3 * HID), buf[L.wh:L.wh + HID * 3 * HID].reshape(HID, 3 * HID)):
        blk = tc_operand(w)
        buf[off:off + blk.size] = blk
        off += blk.size
    return buf


def tf32_round(x: np.ndarray) -> np.ndarray:
    """Round fp32 to TF32 (10 explicit mantissa bits), ties away from zero as cvt.rna.tf32.f32 does."""
    u = np.ascontiguousarray(x, np.float32).view(np.uint32)
    return ((u + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)


def tc_operand(w: np.ndarray) -> np.ndarray:
    """(K, N) weight -> flat [hi | lo], each [K/4][N][4] with block[kc][n][j] = W[4 kc + j][n]."""
    w = np.asarray(w, np.float32)
    K, N = w.shape
    hi = tf32_round(w)
    lo = tf32_round(w - hi)
    def canon(a):
        return a.reshape(K // 4, 4, N).transpose(0, 2, 1).reshape(-1)
    return np.concatenate([canon(hi), canon(lo)])


def count_params(tree) -> int:
    if isinstance(tree, dict):
        return sum(count_params(v) for v in tree.values())
    return int(np.asarray(tree).size)
