"""The PPO update against the REFERENCE'S OWN `DGPPO.update` (dgppo/algo/dgppo.py:136-321,
dgppo/algo/informarl.py:357-457), executed under the NumPy stand-ins by
tools/gen_golden_update_from_reference.py -> tests/golden/ref_update_*.npz.

Checked on the CPU:
  * the oracle (oracle/algo_np.py, oracle/nn_np.py) reproduces every intermediate of the reference's pre-pass:
    Vl scan, Vh with the stored policy carries (and the policy's post-step carry for the final graph), both
    Dec-OCP GAE passes, the CBF advantage merge with the scheduled weight;
  * algo/update.py's three loss functions equal the values of the reference's `get_loss_` closures - to fp32
    rounding against the fp32 run, to ~1e-9 against the same closures evaluated in float64 (oracle/algoshim.x64) -
    and their torch-autograd gradients agree with central finite differences OF THOSE CLOSURES along seeded
    directions confined to groups of parameter leaves (float64 differences: relative 2e-5; the fp32 run's own
    differences within their error bar).
"""
import os

import numpy as np
import pytest
import torch

from dgppo_b200.algo import update as U
from oracle import algo_np, nn_np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ("LidarSpread_n3_obs3", "MPEConnectSpread_n3_obs1", "LidarBicycleTarget_n4_obs3")      # 2 / 3 cost heads, LiDAR / MPE graphs, Spread / Target goals, state_dim 4 / 5
GROUPS = {       # as in tools/gen_golden_update_from_reference.py
    "policy": ("", "GraphTransformer_0", "GraphTransformer_1", "PolicyGNNHead", "RNN_0", "ScaleHid", "OutputDenseMean",
               "OutputDenseStdTrans"),
    "Vl": ("", "GraphTransformer_0", "GraphTransformer_1", "ValueGNNHead", "RNN_0", "Dense_0"),
    "Vh": ("", "GraphTransformer_0", "ValueGNNHead", "RNN_0", "Dense_0"),
}
F = np.float32


def env_cfg(name):
    """The oracle's static env description of a fixture case (tests/golden_util.py)."""
    from tests.golden_util import CASES as ENV_CASES
    return ENV_CASES[name]


def load(name="LidarSpread_n3_obs3"):
    d = np.load(os.path.join(GOLDEN_DIR, f"ref_update_{name}.npz"))
    n, n_obs, b, T, rnn_step, step, train_steps = (int(v) for v in d["meta"])
    trees = {}
    for key in d.files:
        if key.startswith("param:"):
            _, tag, path = key.split(":", 2)
            node = trees.setdefault(tag, {})
            parts = path.split("/")
            for p_ in parts[:-1]:
                node = node.setdefault(p_, {})
            node[parts[-1]] = d[key]
    hyper = dict(zip(("gamma", "gae_lambda", "clip_eps", "coef_ent", "alpha", "cbf_eps", "cbf_weight", "dt",
                      "max_grad_norm"), (float(v) for v in d["hyper"])))
    return d, trees, hyper, (n, n_obs, b, T, rnn_step)


def graphs(d, tag, lo, hi):
    return {k: d[f"{tag}:{k}"][:, lo:hi] for k in ("nodes", "edges", "receivers", "senders")}


def at(g, t):
    return {k: v[:, t] for k, v in g.items()}


# ------------------------------------------------------------------ pre-pass
@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_the_reference_prepass(name):
    d, trees, hp, (n, n_obs, b, T, rnn_step) = load(name)
    assert hp["cbf_weight"] == 2.0                        # step 600 of 1000: past the first schedule boundary
    for tag, vh_key in (("ro", "Vh"), ("det", "Vh_det")):
        g = graphs(d, tag, 0, T + 1)
        rnn = d[f"{tag}:rnn_states"]
        # Vh over graph[t] with the stored carry (dgppo.py:219-220), final graph with the policy's post-step
        # carry from act(next_graph[-1], rnn_states[-1]) (dgppo.py:222-226)
        for t in range(T):
            vh = nn_np.vh_forward(trees["Vh"], at(g, t), rnn[:, t], n)
            np.testing.assert_allclose(vh, d[vh_key][:, t], rtol=2e-5, atol=2e-6)
        _, _, h_fin, _ = nn_np.policy_forward(trees["policy"], at(g, T), rnn[:, T - 1], n, eps=None)
        vh = nn_np.vh_forward(trees["Vh"], at(g, T), h_fin, n)
        np.testing.assert_allclose(vh, d[vh_key][:, T], rtol=2e-5, atol=2e-6)
    # Vl scan from the zero carry, final value on next_graph[-1] (dgppo.py:204-216)
    g = graphs(d, "ro", 0, T + 1)
    h = np.zeros((b, 64), F)
    for t in range(T + 1):
        v, h = nn_np.vl_forward(trees["Vl"], at(g, t), h, n)
        np.testing.assert_allclose(v, d["Vl"][:, t], rtol=2e-5, atol=2e-6)
    # GAE on both records (the deterministic pass uses the stochastic Vl: dgppo.py:268-273), then the merge
    for i in range(b):
        Qh, Ql = algo_np.compute_dec_ocp_gae(d["ro:costs"][i], -d["ro:rewards"][i], d["Vh"][i], d["Vl"][i],
                                             hp["gamma"], hp["gae_lambda"])
        np.testing.assert_allclose(Qh, d["Qh"][i], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(Ql, d["Ql"][i], rtol=1e-5, atol=1e-6)
        Qhd, _ = algo_np.compute_dec_ocp_gae(d["det:costs"][i], -d["det:rewards"][i], d["Vh_det"][i], d["Vl"][i],
                                             hp["gamma"], hp["gae_lambda"])
        np.testing.assert_allclose(Qhd, d["Qh_det"][i], rtol=1e-5, atol=1e-6)
    A, deriv, _, safe = algo_np.cbf_advantage(d["Ql"], d["Vl"], d["Vh"], hp["dt"], hp["alpha"], hp["cbf_eps"],
                                              hp["cbf_weight"])
    near = (np.abs(deriv) < 1e-4).any(-1)                 # a residual within rounding of 0 may flip `is_safe`
    np.testing.assert_allclose(A[~near], d["A"][~near], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(safe.mean(), float(d["safe_data"]), atol=near.mean() + 1e-7)


# -------------------------------------------------------------------- losses
def _torch_inputs(d, T, n, n_obs, dtype=torch.float64, name="LidarSpread_n3_obs3"):
    cfg = env_cfg(name)
    gi = U.GraphIndex(n, cfg.n_ag, cfg.n_ao, cfg.n_nodes, torch.device("cpu"))
    assert cfg.n_nodes == d["ro:nodes"].shape[2] and cfg.n_edges == d["ro:edges"].shape[2]

    def prep(tag):
        a = [torch.tensor(d[f"{tag}:{k}"][:, :T]) for k in ("nodes", "edges", "receivers", "senders")]
        mbT = a[0].shape[0] * T
        return U.prep_graphs(a[0].reshape((mbT,) + a[0].shape[2:]), a[1].reshape((mbT,) + a[1].shape[2:]),
                             a[2].reshape(mbT, -1), a[3].reshape(mbT, -1), gi, dtype)
    return gi, prep("ro"), prep("det")


def _losses(d, trees, hp, dims, dtype=torch.float64, name="LidarSpread_n3_obs3"):
    n, n_obs, b, T, rnn_step = dims
    gi, g, gd = _torch_inputs(d, T, n, n_obs, dtype, name)
    tt = lambda a: torch.tensor(np.asarray(a), dtype=dtype)       # noqa: E731
    tp = {k: U.to_torch_tree(trees[k], "cpu", dtype, requires_grad=True) for k in ("policy", "Vl", "Vh")}
    # the entropy draw: ONE (n, action_dim) sample shared by every graph (distribution.py:37-43 under jit)
    eps = tt(d["entropy_eps"]).expand(b, T, n, 2)
    out = {
        "Vl": (U.loss_Vl(tp["Vl"], g, tt(d["Ql"]), gi, 2, rnn_step), None),
        "Vh": (U.loss_Vh(tp["Vh"], gd, tt(d["det:rnn_states"]), tt(d["Qh_det"]), gi, 1), None),
        "policy": U.loss_policy(tp["policy"], g, tt(d["ro:actions"]), tt(d["ro:log_pis"]), tt(d["A"]), eps, gi, 2,
                                rnn_step, hp["clip_eps"], hp["coef_ent"]),
    }
    return out, tp


@pytest.mark.parametrize("name", CASES)
def test_losses_equal_the_reference_closures(name):
    d, trees, hp, dims = load(name)
    out, _ = _losses(d, trees, hp, dims, name=name)
    for tag in ("Vl", "Vh", "policy"):
        np.testing.assert_allclose(float(out[tag][0]), float(d[f"loss:{tag}"]), rtol=2e-5, err_msg=tag)
        # ... and the same closure evaluated in float64: the two loss FUNCTIONS agree to ~1e-9 (the entropy draw is
        # the fixture's fp32-rounded one here, the unrounded one there)
        np.testing.assert_allclose(float(out[tag][0]), float(d[f"loss64:{tag}"]), rtol=2e-9, err_msg=tag + " (float64)")
    info = out["policy"][1]
    np.testing.assert_allclose(float(info["policy/entropy"]), float(d["aux:policy/entropy"]), rtol=2e-5)
    np.testing.assert_allclose(float(info["policy/total_variation_dist"]),
                               float(d["aux:policy/total_variation_dist"]), rtol=1e-3, atol=1e-6)
    np.testing.assert_allclose(float(info["policy/clip_frac"]), float(d["aux:policy/clip_frac"]), atol=1e-3)


def _flatten(tree, pre=""):
    out = {}
    for k in sorted(tree):
        v = tree[k]
        if isinstance(v, dict):
            out.update(_flatten(v, pre + k + "/"))
        else:
            out[pre + k] = v
    return out


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("tag", ["Vl", "Vh", "policy"])
def test_autograd_gradients_vs_finite_differences_of_the_reference_closures(tag, name):
    d, trees, hp, dims = load(name)
    out, tp = _losses(d, trees, hp, dims, name=name)
    loss = out[tag][0]
    flat_t = _flatten(tp[tag])
    grads = dict(zip(flat_t, torch.autograd.grad(loss, list(flat_t.values()), allow_unused=True)))
    flat_np = _flatten(trees[tag])
    fd = d[f"fd:{tag}"]
    worst = 0.0
    for gi_, group in enumerate(GROUPS[tag]):
        rng = np.random.default_rng(1000 + gi_)            # the generator's `direction`
        dirs = {k: (rng.standard_normal(v.shape).astype(F) if group in k else np.zeros_like(v)) for k, v in flat_np.items()}
        nrm = np.sqrt(sum(float((v.astype(np.float64) ** 2).sum()) for v in dirs.values()))
        dd = sum(float((grads[k].double() * torch.tensor(dirs[k] / F(nrm)).double()).sum())
                 for k in flat_np if grads[k] is not None)
        pn = np.sqrt(sum(float((v.astype(np.float64) ** 2).sum()) for k, v in flat_np.items() if group in k))
        # (1) the closure evaluated in float64 (algoshim.x64), central differences: the reference's directional
        #     derivative to ~1e-7 relative
        #     (steps of 1e-6 / 1e-7 / 1e-8 |p|: a ReLU kink inside the largest step happens once in ~50 directions and
        #     shows as an outlier among the three, so the closest one is compared)
        fd64 = min((float(v) for v in d[f"fd64:{tag}"][gi_]), key=lambda v: abs(v - dd))
        assert abs(dd - fd64) <= 2e-5 * abs(fd64) + 1e-9, (tag, group, dd, d[f"fd64:{tag}"][gi_].tolist())
        worst = max(worst, abs(dd - fd64) / (abs(fd64) + 1e-12))
        # (2) the fp32 run's own differences (steps of 2 % and 1 % of |p|): consistent within their error bar -
        #     step-size spread + fp32 round-off of the loss difference; directions whose derivative is below that
        #     floor, or that cross ReLU / clip kinks within the step, carry no information and are only bounded
        noise = 4e-7 * abs(float(d[f"loss:{tag}"])) / (0.01 * pn)
        bar = 3.0 * abs(fd[gi_, 0] - fd[gi_, 1]) + noise + 0.02 * abs(dd)
        if abs(fd[gi_, 1]) > 5.0 * (noise + abs(fd[gi_, 0] - fd[gi_, 1])):
            assert abs(dd - fd[gi_, 1]) <= bar, (tag, group, dd, fd[gi_].tolist(), bar)
    print(tag, name, "worst relative |autograd - fd64|:", f"{worst:.2e}")


# ------------------------------------------------------------------ rollouts
def _obstacles(d, tag):
    if f"{tag}:obs_theta" not in d.files:        # MPE: the obstacles are nodes of the graph
        return None
    th = d[f"{tag}:obs_theta"].astype(F)
    return dict(center=d[f"{tag}:obs_center"], width=d[f"{tag}:obs_width"], height=d[f"{tag}:obs_height"], theta=th,
                cos=np.cos(th).astype(F), sin=np.sin(th).astype(F), points=d[f"{tag}:obs_points"])


@pytest.mark.parametrize("name", CASES)
def test_oracle_rollout_replays_the_reference_trajectories(name):
    """The reference's `collect` (rollout, trainer/utils.py:22-57: carry stored BEFORE the step) and its
    deterministic `test_rollout` (:60-86: carry stored AFTER the step), both through its own policy, env.step,
    LiDAR and get_graph, against the oracle's scan from the same initial graph: 16 steps, every field."""
    from oracle import env_np
    d, trees, hp, (n, n_obs, b, T, rnn_step) = load(name)
    cfg = env_cfg(name)
    fields = ("n_node", "n_edge", "nodes", "edges", "states", "receivers", "senders", "node_type")
    # deterministic
    out = algo_np.rollout(cfg, trees["policy"], {k: d[f"det:{k}"][:, 0] for k in fields}, _obstacles(d, "det"), None, T)
    np.testing.assert_allclose(out["actions"], d["det:actions"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(out["rnn_states"][:, 1:], d["det:rnn_states"], rtol=1e-5, atol=2e-6)     # post-step
    np.testing.assert_allclose(out["rewards"], d["det:rewards"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(out["costs"], d["det:costs"], rtol=1e-5, atol=1e-6)
    for k in ("receivers", "senders"):
        assert np.array_equal(out[k], d[f"det:{k}"]), k
    np.testing.assert_allclose(out["nodes"], d["det:nodes"], rtol=1e-5, atol=1e-6)
    # stochastic: the draw behind each sampled action is recovered from the reference's action itself
    g = {k: d[f"ro:{k}"][:, 0] for k in fields}
    obst, h = _obstacles(d, "ro"), np.zeros((b, n, 64), F)
    for t in range(T):
        np.testing.assert_allclose(h, d["ro:rnn_states"][:, t], rtol=1e-5, atol=2e-6)                   # pre-step
        _, _, _, (mean, std) = nn_np.policy_forward(trees["policy"], g, h, n, eps=None)
        a_ref = d["ro:actions"][:, t]
        eps = ((np.arctanh(a_ref.astype(np.float64)) - mean) / std).astype(F)
        a, lp, h, _ = nn_np.policy_forward(trees["policy"], g, h, n, eps=eps)
        np.testing.assert_allclose(a, a_ref, rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(lp, d["ro:log_pis"][:, t], rtol=1e-4, atol=1e-4)
        g, r, c, _ = env_np.env_step(cfg, g, a_ref, obst)
        np.testing.assert_allclose(r, d["ro:rewards"][:, t], rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(c, d["ro:costs"][:, t], rtol=1e-5, atol=1e-6)
        for k in ("receivers", "senders"):
            assert np.array_equal(g[k], d[f"ro:{k}"][:, t + 1]), (k, t)
        np.testing.assert_allclose(g["nodes"], d["ro:nodes"][:, t + 1], rtol=1e-5, atol=1e-6)
