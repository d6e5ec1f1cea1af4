"""PPO update (dgppo_b200/algo/update.py) on the CPU: the torch forward functions against the NumPy oracle,
autograd gradients against central finite differences of the oracle in float64, the optimiser against a
hand-computed Adam step.  (The update is library arithmetic - torch autograd - so it runs on any device; the
rollout forward stays the CUDA kernels and is tested on the GPU.)"""
import numpy as np
import pytest
import torch

from dgppo_b200.algo import update as U
from dgppo_b200.algo import params as P
from oracle import nn_np


def make_graph(rng, b, n=3, g=3, n_obs_nodes=6, nd=7):
    # LidarSpread-like: agents, goals (n_ag = g), obstacle nodes (n_ao per agent)
    n_ao = n_obs_nodes // n
    N = n + g + n_obs_nodes + 1
    E = n * n + n * g + n * n_ao
    pad = N - 1
    nodes = rng.standard_normal((b, N, nd)).astype(np.float32)
    nodes[:, pad] = 0
    edges = rng.standard_normal((b, E, 4)).astype(np.float32)
    recv = np.zeros((b, E), np.int32); send = np.zeros((b, E), np.int32)
    e = 0
    for i in range(n):
        for j in range(n):
            recv[:, e], send[:, e] = i, j; e += 1
    for i in range(n):
        for q in range(g):
            recv[:, e], send[:, e] = i, n + q; e += 1
    for i in range(n):
        for k in range(n_ao):
            recv[:, e], send[:, e] = i, n + g + i * n_ao + k; e += 1
    m = rng.random((b, E)) < 0.35
    m[:, :n * n][:, ::n + 1] = True          # self edges are always masked (lidar_spread.py:63-65)
    recv[m] = pad; send[m] = pad
    m[0, :] = False                          # one graph with every other slot live
    return dict(nodes=nodes, edges=edges, receivers=recv, senders=send), (n, g, n_ao, N)


def torch_graph(gr, dims, dtype):
    n, g, n_ao, N = dims
    gi = U.GraphIndex(n, g, n_ao, N, torch.device("cpu"))
    tg = U.prep_graphs(torch.tensor(gr["nodes"]), torch.tensor(gr["edges"]), torch.tensor(gr["receivers"]),
                       torch.tensor(gr["senders"]), gi, dtype)
    return tg, gi


def test_forward_matches_oracle():
    rng = np.random.default_rng(0)
    gr, dims = make_graph(rng, 6)
    n = dims[0]
    tg, gi = torch_graph(gr, dims, torch.float64)
    pol = P.init_policy_params(7, 4, 2, 2, seed=1, jitter=0.1)
    h = rng.standard_normal((6, n, 64)).astype(np.float32) * 0.3
    eps = rng.standard_normal((6, n, 2)).astype(np.float32)
    act, lp, h1, (mean, std) = nn_np.policy_forward(pol, gr, h, n, eps=eps, dt=np.float64)
    tp = U.to_torch_tree(pol, "cpu", torch.float64, requires_grad=False)
    emb = U.gnn(tp["params"]["PolicyNet_0"]["GraphTransformerGNN_0"], tg, gi, 2)
    m_t, s_t, h_t = U.policy_step(tp, emb, torch.tensor(h, dtype=torch.float64))
    np.testing.assert_allclose(m_t.numpy(), mean, rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(s_t.numpy(), std, rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(h_t.numpy(), h1, rtol=1e-9, atol=1e-11)
    lp_t = U.tanh_normal_log_prob(torch.tensor(act), m_t, s_t)
    np.testing.assert_allclose(lp_t.numpy(), lp, rtol=1e-8, atol=1e-9)
    # value nets
    vh = P.init_value_params(7, 4, 2, 1, seed=2, jitter=0.1)
    ref = nn_np.vh_forward(vh, gr, h, n, dt=np.float64)
    tv = U.to_torch_tree(vh, "cpu", torch.float64, requires_grad=False)
    out, _ = U.value_step(tv, U.gnn(tv["params"]["GraphTransformerGNN_0"], tg, gi, 1), torch.tensor(h, dtype=torch.float64))
    np.testing.assert_allclose(out.numpy(), ref, rtol=1e-9, atol=1e-11)
    vl = P.init_value_params(7, 4, 1, 2, seed=3, jitter=0.1)
    hl = rng.standard_normal((6, 64)).astype(np.float32) * 0.3
    ref_v, ref_h = nn_np.vl_forward(vl, gr, hl, n, dt=np.float64)
    tl = U.to_torch_tree(vl, "cpu", torch.float64, requires_grad=False)
    v, hh = U.value_step(tl, U.gnn(tl["params"]["GraphTransformerGNN_0"], tg, gi, 2).mean(dim=1), torch.tensor(hl, dtype=torch.float64))
    np.testing.assert_allclose(v[:, 0].numpy(), ref_v, rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(hh.numpy(), ref_h, rtol=1e-9, atol=1e-11)


def test_log_prob_tails_match_oracle():
    rng = np.random.default_rng(1)
    loc = rng.standard_normal((50, 2)); scale = np.abs(rng.standard_normal((50, 2))) + 0.05
    val = np.clip(rng.standard_normal((50, 2)), -1, 1)
    val[:10] = 1.0; val[10:20] = -1.0; val[20:25] = 0.9995
    ref = nn_np.tanh_normal_log_prob(val, loc, scale, dt=np.float64)
    out = U.tanh_normal_log_prob(torch.tensor(val), torch.tensor(loc), torch.tensor(scale))
    # oracle log_ndtr follows tfp's float64 segments; torch.special.log_ndtr is accurate everywhere
    np.testing.assert_allclose(out.numpy(), ref, rtol=1e-6, atol=1e-8)


def _oracle_policy_loss(pol, gr, dims, actions, lp_old, adv, eps, rnn_step, clip_eps, coef_ent):
    """update_policy.get_loss_ restated on the NumPy oracle (float64)."""
    n = dims[0]
    mb, T = adv.shape[:2]
    C = T // rnn_step
    lps = np.zeros((mb, T, n)); ents = np.zeros((mb, T, n))
    for c in range(C):
        h = np.zeros((mb, n, 64))
        for t in range(c * rnn_step, (c + 1) * rnn_step):
            g = {k: v.reshape((mb, T) + v.shape[1:])[:, t] for k, v in gr.items()}
            _, _, h, (mean, std) = nn_np.policy_forward(pol, g, h, n, eps=None, dt=np.float64)
            lps[:, t] = nn_np.tanh_normal_log_prob(actions[:, t], mean, std, dt=np.float64)
            z = mean + std * eps[:, t]
            fldj = 2.0 * (np.log(2.0) - z - nn_np.softplus(-2.0 * z))
            ents[:, t] = (0.5 + 0.5 * np.log(2 * np.pi) + np.log(std) + fldj).sum(-1)
    ratio = np.exp(lps - lp_old)
    l1 = -ratio * adv
    l2 = -np.clip(ratio, 1 - clip_eps, 1 + clip_eps) * adv
    return np.maximum(l1, l2).mean() - coef_ent * ents.mean()


def test_policy_loss_and_gradient_vs_finite_differences():
    rng = np.random.default_rng(2)
    mb, T, rnn_step = 2, 4, 2
    gr, dims = make_graph(rng, mb * T)
    n = dims[0]
    tg, gi = torch_graph(gr, dims, torch.float64)
    pol = P.init_policy_params(7, 4, 2, 2, seed=4, jitter=0.1)
    actions = np.tanh(rng.standard_normal((mb, T, n, 2)) * 0.5)
    lp_old = rng.standard_normal((mb, T, n)) * 0.1 - 1.0
    adv = rng.standard_normal((mb, T, n))
    eps = rng.standard_normal((mb, T, n, 2))
    tp = U.to_torch_tree(pol, "cpu", torch.float64)
    loss, info = U.loss_policy(tp, tg, torch.tensor(actions), torch.tensor(lp_old), torch.tensor(adv),
                               torch.tensor(eps), gi, 2, rnn_step, 0.25, 1e-2)
    ref = _oracle_policy_loss(pol, gr, dims, actions, lp_old, adv, eps, rnn_step, 0.25, 1e-2)
    np.testing.assert_allclose(float(loss), ref, rtol=1e-9)
    leaves = U.tree_leaves(tp)
    grads = torch.autograd.grad(loss, [t for _, t in leaves])
    # central differences on a few entries of every kind of leaf (fp64 oracle)
    pol64 = U.tree_map(lambda a: np.asarray(a, np.float64).copy(), pol)
    np_leaves = dict(U.tree_leaves(pol64))
    checked = 0
    for (path, _), gt in zip(leaves, grads):
        arr = np_leaves[path]
        for _ in range(2):
            ix = tuple(rng.integers(0, s) for s in arr.shape)
            old = arr[ix]
            h = 1e-6
            arr[ix] = old + h
            lp = _oracle_policy_loss(pol64, gr, dims, actions, lp_old, adv, eps, rnn_step, 0.25, 1e-2)
            arr[ix] = old - h
            lm = _oracle_policy_loss(pol64, gr, dims, actions, lp_old, adv, eps, rnn_step, 0.25, 1e-2)
            arr[ix] = old
            fd = (lp - lm) / (2 * h)
            assert abs(fd - float(gt[ix])) <= 1e-4 * max(1e-3, abs(fd)) + 1e-8, (path, ix, fd, float(gt[ix]))
            checked += 1
    assert checked >= 60


def test_value_losses_gradients_finite_differences():
    rng = np.random.default_rng(3)
    mb, T, rnn_step = 2, 4, 2
    gr, dims = make_graph(rng, mb * T)
    n = dims[0]
    tg, gi = torch_graph(gr, dims, torch.float64)
    # Vl
    vl = U.tree_map(lambda a: np.asarray(a, np.float64), P.init_value_params(7, 4, 1, 2, seed=5, jitter=0.1))
    tgt = rng.standard_normal((mb, T))

    def vl_loss_np(p):
        out = np.zeros((mb, T))
        for c in range(T // rnn_step):
            h = np.zeros((mb, 64))
            for t in range(c * rnn_step, (c + 1) * rnn_step):
                g = {k: v.reshape((mb, T) + v.shape[1:])[:, t] for k, v in gr.items()}
                out[:, t], h = nn_np.vl_forward(p, g, h, n, dt=np.float64)
        return (0.5 * (out - tgt) ** 2).mean()
    tv = U.to_torch_tree(vl, "cpu", torch.float64)
    loss = U.loss_Vl(tv, tg, torch.tensor(tgt), gi, 2, rnn_step)
    np.testing.assert_allclose(float(loss), vl_loss_np(vl), rtol=1e-9)
    _fd_check(rng, vl, tv, loss, vl_loss_np)
    # Vh
    vh = U.tree_map(lambda a: np.asarray(a, np.float64), P.init_value_params(7, 4, 2, 1, seed=6, jitter=0.1))
    hs = rng.standard_normal((mb, T, n, 64)) * 0.3
    tgh = rng.standard_normal((mb, T, n, 2))

    def vh_loss_np(p):
        out = nn_np.vh_forward(p, gr, hs.reshape(mb * T, n, 64), n, dt=np.float64)
        return (0.5 * (out.reshape(tgh.shape) - tgh) ** 2).mean()
    th = U.to_torch_tree(vh, "cpu", torch.float64)
    loss = U.loss_Vh(th, tg, torch.tensor(hs), torch.tensor(tgh), gi, 1)
    np.testing.assert_allclose(float(loss), vh_loss_np(vh), rtol=1e-9)
    _fd_check(rng, vh, th, loss, vh_loss_np)


def _fd_check(rng, np_tree, t_tree, loss, loss_np, per_leaf=2):
    leaves = U.tree_leaves(t_tree)
    grads = torch.autograd.grad(loss, [t for _, t in leaves])
    np_leaves = dict(U.tree_leaves(np_tree))
    for (path, _), gt in zip(leaves, grads):
        arr = np_leaves[path]
        for _ in range(per_leaf):
            ix = tuple(rng.integers(0, s) for s in arr.shape)
            old, h = arr[ix], 1e-6
            arr[ix] = old + h; lp = loss_np(np_tree)
            arr[ix] = old - h; lm = loss_np(np_tree)
            arr[ix] = old
            fd = (lp - lm) / (2 * h)
            assert abs(fd - float(gt[ix])) <= 1e-4 * max(1e-3, abs(fd)) + 1e-8, (path, ix, fd, float(gt[ix]))


def test_adam_if_finite_and_clip():
    p = [torch.tensor([1.0, -2.0], dtype=torch.float64, requires_grad=True)]
    opt = U.AdamIfFinite(p, lr=0.1)
    loss = (p[0] ** 2).sum() * 3.0               # grad = 6 p = [6, -12], norm 13.416 -> clipped to norm 2
    info = U.clip_and_step(opt, p, loss, max_norm=2.0)
    g = np.array([6.0, -12.0]); g = g / np.linalg.norm(g) * 2.0
    m, v = 0.1 * g, 0.001 * g * g
    upd = -0.1 * (m / 0.1) / (np.sqrt(v / 0.001) + 1e-8)
    np.testing.assert_allclose(p[0].detach().numpy(), np.array([1.0, -2.0]) + upd, rtol=1e-12)
    np.testing.assert_allclose(float(info["grad_norm"]), np.linalg.norm([6.0, -12.0]), rtol=1e-12)
    before = p[0].detach().clone()
    opt.step([torch.tensor([float("nan"), 0.0], dtype=torch.float64)], finite=False)     # apply_if_finite: skipped
    assert torch.equal(p[0].detach(), before) and opt.count == 1 and opt.notfinite_count == 1
    # below the clip threshold the gradient passes unchanged (g / max(max_norm, |g|) * max_norm)
    q = [torch.tensor([0.1], dtype=torch.float64, requires_grad=True)]
    o2 = U.AdamIfFinite(q, lr=0.01)
    U.clip_and_step(o2, q, (q[0] ** 2).sum(), max_norm=2.0)
    np.testing.assert_allclose(o2.m[0].numpy(), 0.1 * 0.2, rtol=1e-12)


def test_flat_train_state_matches_leafwise_adam():
    """NetTrainState.step (flat buffer, sync-free skip) == clip_and_step on the list of leaves, step by step, and a
    non-finite gradient leaves parameters and moments untouched."""
    rng = np.random.default_rng(0)
    tree = {"params": {"a": {"kernel": rng.standard_normal((3, 4)), "bias": rng.standard_normal(4)},
                       "b": {"kernel": rng.standard_normal((4, 2))}}}
    st = U.NetTrainState(tree, "cpu", lr=0.05, dtype=torch.float64)
    ref_tree = U.to_torch_tree(tree, "cpu", torch.float64)
    ref_leaves = [t for _, t in U.tree_leaves(ref_tree)]
    opt = U.AdamIfFinite(ref_leaves, 0.05)
    x = torch.tensor(rng.standard_normal((5, 3)))

    def loss_of(t):
        return ((x @ t["params"]["a"]["kernel"] + t["params"]["a"]["bias"]) @ t["params"]["b"]["kernel"]).pow(2).sum() * 7.0
    for it in range(4):
        r1 = st.step(loss_of(st.tree()), 2.0)
        r2 = U.clip_and_step(opt, ref_leaves, loss_of(ref_tree), 2.0)
        np.testing.assert_allclose(float(r1["grad_norm"]), float(r2["grad_norm"]), rtol=1e-12)
        got = st.numpy_tree()
        for (path, leaf) in U.tree_leaves(ref_tree):
            node = got
            for k in path:
                node = node[k]
            np.testing.assert_allclose(node, leaf.detach().numpy().astype(np.float32), rtol=1e-6)
    before = [t.clone() for t in st.state_tensors()]
    r = st.step(loss_of(st.tree()) * float("nan"), 2.0)
    assert float(r["has_nan"]) == 1.0 and float(st.notfinite_count) == 1.0 and float(st.count) == 4.0
    for t, b_ in zip(st.state_tensors()[:3], before[:3]):
        assert torch.equal(t.detach(), b_.detach())
    # set_params-style reload keeps the moments
    st.load(tree)
    np.testing.assert_allclose(st.numpy_tree()["params"]["b"]["kernel"], tree["params"]["b"]["kernel"].astype(np.float32))
    assert float(st.m.abs().sum()) > 0


def test_matmul_rows_split_k_equals_plain_product(monkeypatch):
    """The split-K form of the many-rows product (batched over row blocks against the expanded weight) and its
    autograd gradients equal the plain product's."""
    torch.manual_seed(0)
    x = torch.randn(512, 128, 7, dtype=torch.float64, requires_grad=True)          # 65,536 rows: above the threshold
    W = torch.randn(7, 96, dtype=torch.float64, requires_grad=True)
    out = {}
    for flag in ("1", "0"):
        monkeypatch.setenv("DGPPO_UPDATE_SPLITK", flag)
        y = U.matmul_rows(x, W)
        out[flag] = (y.detach(), torch.autograd.grad((y ** 2).sum(), [x, W]))
    assert out["1"][0].shape == (512, 128, 96)
    torch.testing.assert_close(out["1"][0], out["0"][0], rtol=1e-13, atol=1e-13)
    for a, b_ in zip(out["1"][1], out["0"][1]):
        torch.testing.assert_close(a, b_, rtol=1e-12, atol=1e-9)
    # few rows, or a row count the blocks do not divide: the plain product
    small = torch.randn(100, 7, dtype=torch.float64)
    torch.testing.assert_close(U.matmul_rows(small, W.detach()), small @ W.detach())
