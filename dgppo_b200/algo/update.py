"""PPO minibatch update of DGPPO: losses, gradients, clipping, Adam.

Mirrors `DGPPO.update_inner`'s scan body and the three update functions
(dgppo/algo/dgppo.py:276-321, dgppo/algo/informarl.py:357-457): `update_Vl`
(chunked BPTT through the centralised value's GRU), `update_Vh` (per-agent
constraint value, no recurrence: the carry is the stored policy rnn state),
`update_policy` (chunked BPTT, PPO-clip + entropy), each followed by
`compute_norm_and_clip` (trainer/utils.py:113-118) and `optax.adam` wrapped in
`apply_if_finite` (informarl.py:131-132, dgppo.py:310-311).

The reference obtains the gradients from jax.grad; here the same forward
functions are written over torch tensors and differentiated by torch autograd
(library arithmetic on the GPU: this is the one part of the training step that
is NOT hand-written CUDA; the forward used for ROLLOUTS stays the kernels).
The GraphTransformer is evaluated in the kernels' regrouped form - receivers
are agents only and masked slots (recv = send = pad) feed nothing an agent reads
 - so activations are (graphs, agents, slots) instead of
(graphs, edges, heads, width).

Data parallel: every rank holds `b / world` environments and an equal share
of each minibatch; the flat gradient of each net is mean-all-reduced (NCCL)
before clipping, which reproduces the single-device mean.
"""
from __future__ import annotations

import math
import os
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from ..trainer import distributed as D

N_HEADS = 3
STD_DEV_INIT_INV = float(np.log(np.exp(0.5) - 1.0))       # policy.py:54-59
STD_DEV_MIN = 1e-5


# ----------------------------------------------------------------- pytrees
def tree_leaves(tree, prefix=()) -> List[Tuple[tuple, object]]:
    """(path, leaf) pairs in sorted-key order (jax's dict flattening order)."""
    if isinstance(tree, dict):
        out = []
        for k in sorted(tree):
            out += tree_leaves(tree[k], prefix + (k,))
        return out
    return [(prefix, tree)]


def tree_map(fn, tree):
    if isinstance(tree, dict):
        return {k: tree_map(fn, v) for k, v in tree.items()}
    return fn(tree)


def to_torch_tree(tree, device, dtype=torch.float32, requires_grad=True):
    return tree_map(lambda a: torch.tensor(np.asarray(a), device=device, dtype=dtype, requires_grad=requires_grad), tree)


def to_numpy_tree(tree):
    return tree_map(lambda t: t.detach().to(torch.float32).cpu().numpy(), tree)


# ------------------------------------------------------------------ layers
SPLIT_K_ROWS, SPLIT_K = 32768, 64


def matmul_rows(x, W):
    """x @ W for x with very many rows (every slot of a minibatch: 1e5 .. 1e6) and a small W.  The weight gradient
    of a plain mm is X^T dY with K = rows - one thin GEMM the library runs on a handful of CTAs; written as a batched
    product over SPLIT_K row blocks against the expanded W, autograd's backward becomes SPLIT_K partial products
    + a sum (split-K), which fills the machine.  DGPPO_UPDATE_SPLITK=0 keeps the plain product."""
    rows = x.numel() // x.shape[-1]
    if rows < SPLIT_K_ROWS or rows % SPLIT_K or os.environ.get("DGPPO_UPDATE_SPLITK", "1") == "0":
        return x @ W
    y = torch.bmm(x.reshape(SPLIT_K, rows // SPLIT_K, x.shape[-1]), W.unsqueeze(0).expand(SPLIT_K, *W.shape))
    return y.reshape(x.shape[:-1] + (W.shape[-1],))


def dense(x, p):
    y = matmul_rows(x, p["kernel"])
    return y + p["bias"] if "bias" in p else y


def layer_norm(x, p, eps=1e-6):
    """flax nn.LayerNorm defaults (eps 1e-6, scale + bias over the last axis).  The library's fused kernel pair
    (forward + backward) instead of ~30 elementwise launches; flax's "fast variance" E[x^2] - E[x]^2 and the
    two-pass variance used here are the same function (1e-7 apart in fp32)."""
    return torch.nn.functional.layer_norm(x, x.shape[-1:], p["scale"], p["bias"], eps)


def mlp_head(x, p):
    for i in range(2):
        x = torch.relu(layer_norm(dense(x, p[f"Dense_{i}"]), p[f"LayerNorm_{i}"]))
    return x


def gru_weights(p_rnn):
    """flax GRUCell leaves (ir, iz, in with bias; hr, hz without; hn with bias: rnn.py / SURVEY A.4) -> the
    (w_ih, w_hh, b_ih, b_hh) of torch.gru_cell, gate order r | z | n.  Built once per loss evaluation."""
    (c,) = list(p_rnn.values())
    w_ih = torch.cat([c["ir"]["kernel"], c["iz"]["kernel"], c["in"]["kernel"]], dim=1).t()
    w_hh = torch.cat([c["hr"]["kernel"], c["hz"]["kernel"], c["hn"]["kernel"]], dim=1).t()
    b_ih = torch.cat([c["ir"]["bias"], c["iz"]["bias"], c["in"]["bias"]])
    zero = torch.zeros_like(c["hn"]["bias"])
    b_hh = torch.cat([zero, zero, c["hn"]["bias"]])
    return w_ih, w_hh, b_ih, b_hh


def gru_apply(w, h, x):
    """r = sigmoid(W_ir x + b_ir + W_hr h), z likewise, n = tanh(W_in x + b_in + r (W_hn h + b_hn)),
    h' = (1 - z) n + z h: flax's GRUCell and torch.gru_cell agree on this form.  Two GEMMs + one fused
    pointwise kernel each way on CUDA."""
    shape = h.shape
    out = torch.gru_cell(x.reshape(-1, x.shape[-1]), h.reshape(-1, shape[-1]), *w)
    return out.reshape(shape)


def gru_cell(p_rnn, h, x):
    return gru_apply(gru_weights(p_rnn), h, x)


class GraphIndex:
    """Static slot structure of an env's graphs: receiver agent i owns `deg` edge slots (agents, goals,
    obstacles / hits); slot (i, t) is edge `eidx[i, t]` of the flat edge list (utils/graph.py:212-247)."""

    def __init__(self, n: int, n_ag: int, n_ao: int, n_nodes: int, device):
        self.n, self.N, self.deg = n, n_nodes, n + n_ag + n_ao
        i = np.arange(n)[:, None]
        e = np.concatenate([i * n + np.arange(n)[None], n * n + i * n_ag + np.arange(n_ag)[None],
                            n * n + n * n_ag + i * n_ao + np.arange(n_ao)[None]], axis=1)
        self.eidx = torch.as_tensor(e.reshape(-1), dtype=torch.long, device=device)     # (n * deg)


def graph_transformer_dense(p, x, edge_feat, sidx, mask, nmask, gi: GraphIndex, d: int, agents_only: bool):
    """One GraphTransformer layer + update (nn/gnn.py:78-117) in receiver-major form, projections as the reference
    writes them (q, k, v per node).  Kept as the A/B partner of `graph_transformer` (DGPPO_UPDATE_GNN=dense).
    x (B,N,in); edge_feat (B,n,deg,4); sidx (B,n,deg) sender node of each slot; mask (B,n,deg) slot is live."""
    B, N, n, H = x.shape[0], gi.N, gi.n, N_HEADS
    q = dense(x[:, :n], p["Dense_0"]).reshape(B, n, H, d)
    k = dense(x, p["Dense_1"]).reshape(B, N, H, d)
    v = dense(x, p["Dense_2"]).reshape(B, N, H, d)
    scores = torch.einsum("bihd,bjhd->bhij", q, k) / math.sqrt(d)                        # (B,H,n,N)
    sx = sidx.unsqueeze(1).expand(B, H, n, gi.deg)
    s = torch.gather(scores, 3, sx)
    m = mask.unsqueeze(1)
    s = s.masked_fill(~m, -torch.inf)
    smax = s.amax(dim=-1, keepdim=True)
    smax = torch.where(torch.isfinite(smax), smax, torch.zeros_like(smax))              # rows without a live slot
    ex = torch.exp(s - smax) * m
    a = ex / ex.sum(-1, keepdim=True).clamp_min(1e-38)                                   # segment_softmax per head
    a_full = torch.zeros((B, H, n, N), dtype=x.dtype, device=x.device).scatter_add(3, sx, a)
    agg_v = torch.einsum("bhij,bjhd->bihd", a_full, v)                                   # sum_e a (W_v x_s + b_v)
    ae = torch.einsum("bhit,bitc->bihc", a, edge_feat)                                   # sum_e a edge_e
    we = p["Dense_3"]["kernel"].reshape(-1, H, d)                                        # (4, H, d)
    agg_e = torch.einsum("bihc,chd->bihd", ae, we)
    agg = (agg_v + agg_e).mean(dim=2)                                                    # mean over heads
    upd = dense(x[:, :n] if agents_only else x, p["Dense_4"])
    if agents_only:
        return torch.relu(upd + agg)
    return torch.relu(torch.cat([upd[:, :n] + agg, upd[:, n:]], dim=1))


def graph_transformer(p, x, edge_feat, sidx, mask, nmask, gi: GraphIndex, d: int, agents_only: bool):
    """The same layer in the kernels' regrouped algebra, which suits the library GEMMs far
    better than per-node key / value projections: with q_h = x_i Wq_h + bq_h,

        score_h(i, j) = q_h . (Wk_h^T x_j + bk_h) = x_j . (x_i Wqk_h + bqk_h) + (x_i wqb_h + bqb_h)
        agg(i)        = 1/H sum_h [ (sum_j a_hij x_j) Wv_h + (sum_j a_hij) bv_h + (sum_t a_hit edge_it) We_h ]

    so the layer is one (B n, in) GEMM for the merged query-key rows, two batched GEMMs over whole graphs
    ((n H, in) x (in, N) scores and (n H, N) x (N, in) weighted sender features), and one (B n, H (in + 5)) x (., d)
    GEMM for the stacked value | bias | edge block - no per-node k / v tensors, no scatter.  The softmax runs over
    the sender NODES of a receiver (every live slot of a receiver is a distinct node): `nmask` (B,n,N).
    Parameters stay the reference's leaves; the merged blocks are formed from them inside the autograd graph."""
    B, N, n, H = x.shape[0], gi.N, gi.n, N_HEADS
    IN = x.shape[-1]
    xa = x[:, :n]
    Wq, bq = p["Dense_0"]["kernel"].reshape(IN, H, d), p["Dense_0"]["bias"].reshape(H, d)
    Wk, bk = p["Dense_1"]["kernel"].reshape(IN, H, d), p["Dense_1"]["bias"].reshape(H, d)
    Wv, bv = p["Dense_2"]["kernel"].reshape(IN, H, d), p["Dense_2"]["bias"].reshape(H, d)
    We = p["Dense_3"]["kernel"].reshape(-1, H, d)                                        # (4, H, d)
    W1 = torch.cat([torch.einsum("ahd,chd->ahc", Wq, Wk).reshape(IN, H * IN),            # x_i -> Wk_h^T q_h   (H in)
                    torch.einsum("ahd,hd->ah", Wq, bk)], dim=1)                          # x_i -> q_h . bk_h   (H)
    b1 = torch.cat([torch.einsum("hd,chd->hc", bq, Wk).reshape(H * IN), (bq * bk).sum(-1)])
    t = matmul_rows(xa, W1) + b1                                                         # (B,n,H in + H)
    qt, qb = t[..., :H * IN].reshape(B, n * H, IN), t[..., H * IN:]
    s = (torch.bmm(qt, x.transpose(1, 2)).reshape(B, n, H, N) + qb.unsqueeze(-1)) / math.sqrt(d)
    m = nmask.unsqueeze(2)                                                               # (B,n,1,N)
    # segment_softmax per head over the live senders: masked entries get exp(-1e30 - max) = 0 exactly; a receiver
    # without any live sender would come out uniform, the mask product makes it all-zero (as the reference's
    # empty segment is)
    a = torch.softmax(s.masked_fill(~m, -1e30), dim=-1) * m
    wx = torch.bmm(a.reshape(B, n * H, N), x).reshape(B, n, H, IN)                       # sum_j a x_j
    # per-slot weights for the edge features (a masked slot points at node 0: zeroed by the slot mask)
    a_slot = a.gather(3, sidx.unsqueeze(2).expand(B, n, H, gi.deg)) * mask.unsqueeze(2)
    ae = (a_slot.unsqueeze(-1) * edge_feat.unsqueeze(2)).sum(3)                          # (B,n,H,4)
    feat = torch.cat([wx, a.sum(-1, keepdim=True), ae], dim=-1).reshape(B * n, H * (IN + 5))
    Wagg = torch.cat([Wv, bv.unsqueeze(0), We], dim=0).permute(1, 0, 2).reshape(H * (IN + 5), d)
    agg = matmul_rows(feat, Wagg).reshape(B, n, d) / H                                   # mean over heads
    upd = dense(xa if agents_only else x, p["Dense_4"])
    if agents_only:
        return torch.relu(upd + agg)
    return torch.relu(torch.cat([upd[:, :n] + agg, upd[:, n:]], dim=1))


def gnn(p, g, gi: GraphIndex, n_layers: int, msg_dim=32, out_dim=64):
    """GraphTransformerGNN (nn/gnn.py:127-142) -> agent embeddings (B,n,64)."""
    x = g["nodes"]
    # sender-node mask of every receiver: node j is the sender of a live slot of agent i (slots -> distinct nodes)
    nmask = torch.zeros((x.shape[0], gi.n, gi.N), dtype=x.dtype, device=x.device).scatter_add_(
        2, g["sidx"], g["mask"].to(x.dtype)) > 0
    layer = graph_transformer_dense if os.environ.get("DGPPO_UPDATE_GNN") == "dense" else graph_transformer
    for i in range(n_layers):
        last = i == n_layers - 1
        x = layer(p[f"GraphTransformer_{i}"], x, g["edge_feat"], g["sidx"], g["mask"], nmask, gi,
                  out_dim if last else msg_dim, agents_only=last)
    return x


def prep_graphs(nodes, edges, recv, send, gi: GraphIndex, dtype):
    """Flat graph arrays (B,N,nd), (B,E,4), (B,E), (B,E) -> the receiver-major inputs of `gnn`."""
    B = nodes.shape[0]
    pad = gi.N - 1
    r = recv[:, gi.eidx].reshape(B, gi.n, gi.deg)
    s = send[:, gi.eidx].reshape(B, gi.n, gi.deg).long()
    mask = r != pad
    return {"nodes": nodes.to(dtype), "edge_feat": edges[:, gi.eidx].reshape(B, gi.n, gi.deg, -1).to(dtype),
            "sidx": torch.where(mask, s, torch.zeros_like(s)), "mask": mask}


# ------------------------------------------------------------- tanh-Normal
def _normal_log_prob(x, loc, scale):
    dd = x / scale - loc / scale
    return -0.5 * dd * dd - (0.5 * math.log(2.0 * math.pi) + torch.log(scale))


def tanh_normal_log_prob(value, loc, scale, threshold=0.999):
    """TanhTransformedDistribution.log_prob, summed over the action axis (distribution.py:25-35)."""
    inv_thr = math.atanh(threshold)
    log_eps = math.log(1.0 - threshold)
    lp_left = torch.special.log_ndtr((-inv_thr - loc) / scale) - log_eps
    lp_right = torch.special.log_ndtr(-((inv_thr - loc) / scale)) - log_eps
    v = torch.clamp(value, -threshold, threshold)
    x = torch.atanh(v)
    fldj = 2.0 * (math.log(2.0) - x - torch.nn.functional.softplus(-2.0 * x))
    inner = _normal_log_prob(x, loc, scale) - fldj
    lp = torch.where(v <= -threshold, lp_left, torch.where(v >= threshold, lp_right, inner))
    return lp.sum(-1)


def tanh_normal_entropy(loc, scale, eps):
    """TanhTransformedDistribution.entropy (distribution.py:37-43): Normal entropy + the Tanh forward
    log-det-Jacobian at ONE reparameterised sample, summed over the action axis."""
    z = loc + scale * eps
    fldj = 2.0 * (math.log(2.0) - z - torch.nn.functional.softplus(-2.0 * z))
    ent = 0.5 + 0.5 * math.log(2.0 * math.pi) + torch.log(scale)
    return (ent + fldj).sum(-1)


# ---------------------------------------------------------------- networks
def policy_step(params, emb, h):
    """Head + GRU + TanhNormal parameters for agent embeddings emb (B,n,64), carry h (B,n,64)."""
    p = params["params"]
    base = p["PolicyNet_0"]
    x = mlp_head(emb, base["PolicyGNNHead"])
    h = gru_cell(base["RNN_0"], h, x)
    f = dense(h, p["ScaleHid"])
    mean = dense(f, p["OutputDenseMean"])
    std = torch.nn.functional.softplus(dense(f, p["OutputDenseStdTrans"]) + STD_DEV_INIT_INV) + STD_DEV_MIN
    return mean, std, h


def value_step(params, emb, h):
    p = params["params"]
    x = mlp_head(emb, p["ValueGNNHead"])
    h = gru_cell(p["RNN_0"], h, x)
    return dense(h, p["Dense_0"]), h


# ------------------------------------------------------------------ losses
def chunk_graphs(rollout_arrays, idx, T, gi, dtype):
    """Graphs of envs `idx`, all T slots, flattened env-major: (mb*T, ...)."""
    nodes, edges, recv, send = (a[idx, :T] for a in rollout_arrays)
    mb = nodes.shape[0]
    return prep_graphs(nodes.reshape((mb * T,) + nodes.shape[2:]), edges.reshape((mb * T,) + edges.shape[2:]),
                       recv.reshape(mb * T, -1), send.reshape(mb * T, -1), gi, dtype)


def _scan_gru(w, x_all):
    """GRU over axis 2 of x_all (mb, C, steps, ..., 64) from a zero carry per chunk (informarl.py:365,412):
    only this recurrence is sequential - the head before it and the output layers after it run once over all
    slots.  -> carries AFTER each step, same shape as x_all."""
    h = torch.zeros_like(x_all[:, :, 0])
    hs = []
    for t in range(x_all.shape[2]):
        h = gru_apply(w, h, x_all[:, :, t])
        hs.append(h)
    return torch.stack(hs, dim=2)


def loss_Vl(params, g, targets, gi, n_layers, rnn_step):
    """update_Vl.get_loss_ (informarl.py:367-374): scan_Vl over chunks of rnn_step slots, zero initial carry.
    g: graphs (mb*T, ...); targets (mb, T)."""
    mb, T = targets.shape
    p = params["params"]
    emb = gnn(p["GraphTransformerGNN_0"], g, gi, n_layers).mean(dim=1)                      # (mb*T, 64)
    x = mlp_head(emb, p["ValueGNNHead"]).reshape(mb, T // rnn_step, rnn_step, -1)
    hs = _scan_gru(gru_weights(p["RNN_0"]), x)
    Vl = dense(hs, p["Dense_0"]).reshape(mb, T)
    return (0.5 * (Vl - targets) ** 2).mean()


def loss_Vh(params, g, rnn_states, targets, gi, n_layers):
    """update_Vh.get_loss (dgppo.py:304-311): no recurrence, the carry is the stored rnn state.
    rnn_states (mb, T, n, 64); targets (mb, T, n, n_cost)."""
    mb, T, n, _ = rnn_states.shape
    emb = gnn(params["params"]["GraphTransformerGNN_0"], g, gi, n_layers)                   # (mb*T, n, 64)
    Vh, _ = value_step(params, emb, rnn_states.reshape(mb * T, n, -1))
    return (0.5 * (Vh.reshape(targets.shape) - targets) ** 2).mean()


def loss_policy(params, g, actions, log_pis_old, adv, eps, gi, n_layers, rnn_step, clip_eps, coef_ent):
    """update_policy.get_loss_ (informarl.py:416-437).  actions (mb,T,n,2), log_pis_old / adv (mb,T,n),
    eps (mb,T,n,2) the N(0,1) draw behind the one-sample entropy estimate."""
    mb, T, n, nu = actions.shape
    p = params["params"]
    base = p["PolicyNet_0"]
    emb = gnn(base["GraphTransformerGNN_0"], g, gi, n_layers)                               # (mb*T, n, 64)
    x = mlp_head(emb, base["PolicyGNNHead"]).reshape(mb, T // rnn_step, rnn_step, n, -1)
    hs = _scan_gru(gru_weights(base["RNN_0"]), x).reshape(mb, T, n, -1)
    f = dense(hs, p["ScaleHid"])
    mean = dense(f, p["OutputDenseMean"])
    std = torch.nn.functional.softplus(dense(f, p["OutputDenseStdTrans"]) + STD_DEV_INIT_INV) + STD_DEV_MIN
    log_pis = tanh_normal_log_prob(actions, mean, std)
    entropy = tanh_normal_entropy(mean, std, eps)
    ratio = torch.exp(log_pis - log_pis_old)
    l1 = -ratio * adv
    l2 = -torch.clamp(ratio, 1.0 - clip_eps, 1.0 + clip_eps) * adv
    loss = torch.maximum(l1, l2).mean() - coef_ent * entropy.mean()
    info = {"policy/clip_frac": (l2 > l1).float().mean(), "policy/entropy": entropy.mean(),
            "policy/total_variation_dist": 0.5 * (ratio - 1.0).abs().mean()}
    return loss, info


# --------------------------------------------------------------- optimiser
class NetTrainState:
    """Parameters of one net as ONE flat fp32 tensor (the leaves are views of it, rebuilt per step), with the
    state of optax.apply_if_finite(optax.adam(lr), ...) (b1 0.9, b2 0.999, eps 1e-8) beside it.  Flat because the
    gradient all-reduce wants one buffer and because the whole step then is a handful of kernels with no host
    synchronisation - which is what lets it be captured in a CUDA graph (GraphedStep)."""

    def __init__(self, tree_np, device, lr: float, dtype=torch.float32, b1=0.9, b2=0.999, eps=1e-8):
        self.lr, self.b1, self.b2, self.eps = lr, b1, b2, eps
        self.meta, off, chunks = [], 0, []
        for path, leaf in tree_leaves(tree_np):
            a = np.asarray(leaf)
            self.meta.append((path, tuple(a.shape), off, a.size))
            chunks.append(torch.as_tensor(a.reshape(-1), dtype=dtype))
            off += a.size
        self.flat = torch.cat(chunks).to(device).requires_grad_(True)
        self.m = torch.zeros_like(self.flat, requires_grad=False)
        self.v = torch.zeros_like(self.flat, requires_grad=False)
        self.count = torch.zeros((), dtype=torch.float64, device=device)         # applied steps
        self.notfinite_count = torch.zeros((), dtype=torch.float64, device=device)

    def state_tensors(self):
        return [self.flat, self.m, self.v, self.count, self.notfinite_count]

    def tree(self):
        """Nested dict of (differentiable) views of the flat parameter tensor."""
        out = {}
        for path, shape, off, size in self.meta:
            node = out
            for k in path[:-1]:
                node = node.setdefault(k, {})
            node[path[-1]] = self.flat[off:off + size].view(shape)
        return out

    def numpy_tree(self):
        flat = self.flat.detach().to(torch.float32).cpu().numpy()
        out = {}
        for path, shape, off, size in self.meta:
            node = out
            for k in path[:-1]:
                node = node.setdefault(k, {})
            node[path[-1]] = flat[off:off + size].reshape(shape).copy()
        return out

    @torch.no_grad()
    def load(self, tree_np):
        for (path, leaf), (p2, shape, off, size) in zip(tree_leaves(tree_np), self.meta):
            assert path == p2
            self.flat[off:off + size].copy_(torch.as_tensor(np.asarray(leaf).reshape(-1), dtype=self.flat.dtype))

    def step(self, loss: torch.Tensor, max_norm: float) -> dict:
        """grad -> mean all-reduce over the ranks (the flat buffer itself) -> has_any_nan_or_inf ->
        compute_norm_and_clip (trainer/utils.py:113-118: g / max(max_norm, |g|) * max_norm) -> Adam, skipped as a
        whole when the gradient is not finite.  No host synchronisation: the skip is a 0 / 1 factor."""
        (g,) = torch.autograd.grad(loss, [self.flat])
        (g,) = D.allreduce_mean_flat([g])
        with torch.no_grad():
            sq = (g * g).sum()
            g_norm = torch.sqrt(sq)
            ok = torch.isfinite(sq)
            f = ok.to(self.flat.dtype)
            g = torch.where(ok, g * (max_norm / torch.clamp(g_norm, min=max_norm)), torch.zeros_like(g))
            self.count += ok.to(torch.float64)
            self.notfinite_count += (~ok).to(torch.float64)
            cnt = torch.clamp(self.count, min=1.0)
            c1 = (1.0 - self.b1 ** cnt).to(self.flat.dtype)
            c2 = (1.0 - self.b2 ** cnt).to(self.flat.dtype)
            self.m += f * (1.0 - self.b1) * (g - self.m)
            self.v += f * (1.0 - self.b2) * (g * g - self.v)
            self.flat -= f * (self.lr / c1) * self.m / (torch.sqrt(self.v / c2) + self.eps)
        return {"grad_norm": g_norm, "has_nan": 1.0 - f}


class GraphedStep:
    """A training step `fn(*tensors) -> dict of 0-dim tensors` captured in a CUDA graph: inputs are copied into
    static buffers, the step (forward, backward, all-reduce, clip, Adam) is replayed with one launch.  The eager
    step issues several thousand small kernels (16 recurrent steps x three nets, forward and backward); replay
    removes their launch cost.  Warm-up iterations run for real (as capture requires), so the optimiser state
    is restored afterwards from a snapshot."""

    def __init__(self, fn, inputs, state_tensors):
        self.fn = fn
        self.static_in = [x.clone() for x in inputs]
        saved = [t.detach().clone() for t in state_tensors]

        def restore():
            with torch.no_grad():
                for t, s_ in zip(state_tensors, saved):
                    t.copy_(s_)
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(2):
                fn(*self.static_in)
        cur.wait_stream(side)
        restore()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = fn(*self.static_in)
        restore()

    def __call__(self, *inputs):
        for s_, x in zip(self.static_in, inputs):
            s_.copy_(x)
        self.graph.replay()
        return self.out


# legacy list-of-leaves optimiser (kept for the unit tests of the Adam / clip arithmetic)
class AdamIfFinite:
    """optax.apply_if_finite(optax.adam(lr), max_consecutive_errors) on a list of leaves
    (b1 0.9, b2 0.999, eps 1e-8): a step with any non-finite gradient is skipped."""

    def __init__(self, leaves: List[torch.Tensor], lr: float, b1=0.9, b2=0.999, eps=1e-8):
        self.leaves, self.lr, self.b1, self.b2, self.eps = leaves, lr, b1, b2, eps
        self.m = [torch.zeros_like(p) for p in leaves]
        self.v = [torch.zeros_like(p) for p in leaves]
        self.count = 0
        self.notfinite_count = 0

    @torch.no_grad()
    def step(self, grads: List[torch.Tensor], finite: bool):
        if not finite:
            self.notfinite_count += 1
            return
        self.count += 1
        c1, c2 = 1.0 - self.b1 ** self.count, 1.0 - self.b2 ** self.count
        torch._foreach_mul_(self.m, self.b1)
        torch._foreach_add_(self.m, grads, alpha=1.0 - self.b1)
        torch._foreach_mul_(self.v, self.b2)
        torch._foreach_addcmul_(self.v, grads, grads, value=1.0 - self.b2)
        denom = torch._foreach_sqrt(self.v)
        torch._foreach_div_(denom, math.sqrt(c2))
        torch._foreach_add_(denom, self.eps)
        torch._foreach_addcdiv_(self.leaves, self.m, denom, value=-self.lr / c1)


def clip_and_step(opt: AdamIfFinite, leaves: List[torch.Tensor], loss: torch.Tensor, max_norm: float) -> dict:
    """grad -> mean all-reduce over ranks (one flat buffer) -> has_any_nan_or_inf -> compute_norm_and_clip
    (trainer/utils.py:113-118: g / max(max_norm, |g|) * max_norm) -> Adam."""
    grads = list(torch.autograd.grad(loss, leaves))
    grads = D.allreduce_mean_flat(grads)
    with torch.no_grad():
        sq = torch.stack([(g * g).sum() for g in grads]).sum()
        g_norm = torch.sqrt(sq)
        finite = bool(torch.isfinite(sq))
        scale = max_norm / torch.clamp(g_norm, min=max_norm)
        grads = [g * scale for g in grads]
    opt.step(grads, finite)
    return {"grad_norm": g_norm, "has_nan": 0.0 if finite else 1.0}
