"""Generate tests/golden/ref_update_*.npz by EXECUTING THE REFERENCE'S OWN `DGPPO.update`
(/root/reference/dgppo/algo/dgppo.py:136-321, informarl.py:357-457) under the NumPy stand-ins
(oracle/jaxshim.py, flaxshim.py, algoshim.py):

    python tools/gen_golden_update_from_reference.py        # writes tests/golden/ref_update_<case>.npz

One small problem (LidarSpread n = 3, obs = 3, 4 envs, T = 16, rnn_step = 8) goes through the reference's
`algo.collect` (its own rollout, policy and env code), then `algo.update(rollout, step)`: deterministic
rollout, Vl scan, Vh on both records, both Dec-OCP GAE passes, the CBF advantage merge, and the three loss
functions of the minibatch scan.  The fixture keeps

  * the three parameter pytrees and both rollouts (graphs, actions, log_pis, carries, rewards, costs),
  * every intermediate of the pre-pass (Vl, Vh, Qh, Ql, Vh_det, Qh_det, the merged advantage A),
  * the VALUES of the reference's own `get_loss_` closures (update_Vl / update_Vh / update_policy) at the
    initial parameters, and their central finite differences along seeded directions confined to groups of
    parameter leaves, at two step sizes (the spread between the two is the error bar the tests use).

`jax.value_and_grad` is a stand-in that records the closure (oracle/algoshim.py), so the numbers are the
reference's arithmetic; no autodiff is involved on the reference side.  tests/test_update_reference.py checks
algo/update.py (losses and autograd gradients) and the oracle against them on the CPU,
tests/test_gpu_update_reference.py the kernels' pre-pass on the GPU.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import algoshim  # noqa: E402
from oracle import jaxshim as J  # noqa: E402

algoshim.install()
import jax  # noqa: E402  (the shim)

OUT = os.path.join(ROOT, "tests", "golden")
GRAPH_FIELDS = ("n_node", "n_edge", "nodes", "edges", "states", "receivers", "senders", "node_type")
ENTROPY_SEED = 4242

# direction groups: a finite difference along a random direction confined to the leaves whose path contains
# the substring (all leaves for "")
GROUPS = {
    "policy": ("", "GraphTransformer_0", "GraphTransformer_1", "PolicyGNNHead", "RNN_0", "ScaleHid", "OutputDenseMean",
               "OutputDenseStdTrans"),
    "Vl": ("", "GraphTransformer_0", "GraphTransformer_1", "ValueGNNHead", "RNN_0", "Dense_0"),
    "Vh": ("", "GraphTransformer_0", "ValueGNNHead", "RNN_0", "Dense_0"),
}


def flatten(tree, pre=""):
    out = {}
    for k in sorted(tree):
        v = tree[k]
        if isinstance(v, dict):
            out.update(flatten(v, pre + k + "/"))
        else:
            out[pre + k] = np.asarray(v, np.float32)
    return out


def unflatten(flat):
    tree = {}
    for path, v in flat.items():
        node = tree
        parts = path.split("/")
        for p_ in parts[:-1]:
            node = node.setdefault(p_, {})
        node[parts[-1]] = v
    return tree


def direction(flat, group, seed):
    """Unit-norm random direction over the leaves of `group` (sorted path order), zeros elsewhere."""
    rng = np.random.default_rng(seed)
    d = {k: (rng.standard_normal(v.shape).astype(np.float32) if group in k else np.zeros_like(v)) for k, v in flat.items()}
    nrm = np.sqrt(sum(float((v.astype(np.float64) ** 2).sum()) for v in d.values()))
    return {k: (v / np.float32(nrm)).astype(np.float32) for k, v in d.items()}


def stack_graphs(graph, next_graph):
    """(b, T, ...) graph + next_graph -> (b, T + 1, ...) per field."""
    return {k: np.concatenate([np.asarray(getattr(graph, k)), np.asarray(getattr(next_graph, k))[:, -1:]], axis=1)
            for k in GRAPH_FIELDS}


def closure_vars(f):
    return dict(zip(f.__code__.co_freevars, [c.cell_contents for c in f.__closure__]))


CASES = {   # name -> (module, class, n_agents, n_obs)
    "LidarSpread_n3_obs3": ("dgppo.env.lidar_env.lidar_spread", "LidarSpread", 3, 3),
    "MPEConnectSpread_n3_obs1": ("dgppo.env.mpe.mpe_connect_spread", "MPEConnectSpread", 3, 1),     # three cost heads
    "LidarBicycleTarget_n4_obs3": ("dgppo.env.lidar_env.lidar_bicycle_target", "LidarBicycleTarget", 4, 3),   # state_dim 5, one goal per agent
}


def run(name, b=4, T=16, rnn_step=8, step=600, train_steps=1000):
    import importlib
    import dgppo.algo.dgppo as M
    mod, cls, n, n_obs = CASES[name]
    Env = getattr(importlib.import_module(mod), cls)
    params = dict(Env.PARAMS)
    params["n_obs"] = n_obs
    env = Env(num_agents=n, area_size=None, max_step=T, dt=0.03, params=params)
    algo = M.DGPPO(env=env, node_dim=env.node_dim, edge_dim=env.edge_dim, state_dim=env.state_dim,
                   action_dim=env.action_dim, n_agents=n, batch_size=b * T, rnn_step=rnn_step, seed=0,
                   train_steps=train_steps)
    ro = algo.collect(algo.params, jax.random.split(jax.random.PRNGKey(5), b))

    gae_calls, det = [], []
    orig_gae, orig_det = M.compute_dec_ocp_gae, algo.det_rollout_fn

    def rec_gae(*a, **k):
        out = orig_gae(*a, **k)
        gae_calls.append((dict(k), out))
        return out

    def rec_det(*a, **k):
        det.append(orig_det(*a, **k))
        return det[-1]
    M.compute_dec_ocp_gae, algo.det_rollout_fn = rec_gae, rec_det
    orig_shuffle = np.random.shuffle
    np.random.shuffle = lambda x: None                  # minibatch order = env order (the losses are means)
    del algoshim.CAPTURED[:]
    try:
        with algoshim.trace_constants(ENTROPY_SEED):
            info = algo.update(ro, step)
    finally:
        M.compute_dec_ocp_gae, np.random.shuffle = orig_gae, orig_shuffle
    (f_Vl, p_Vl), (f_Vh, p_Vh), (f_pi, p_pi) = algoshim.CAPTURED
    det_ro = det[0]
    assert len(gae_calls) == 2 * b

    save = {"meta": np.array([n, n_obs, b, T, rnn_step, step, train_steps]),
            "hyper": np.array([algo.gamma, algo.gae_lambda, algo.clip_eps, algo.coef_ent, algo.alpha, algo.cbf_eps,
                               algo.cbf_schedule_fn(step), env.dt, algo.max_grad_norm], np.float64),
            "entropy_eps": algoshim.entropy_eps(ENTROPY_SEED, n, env.action_dim)}
    for tag, r in (("ro", ro), ("det", det_ro)):
        for k, v in stack_graphs(r.graph, r.next_graph).items():
            save[f"{tag}:{k}"] = v
        save[f"{tag}:actions"] = np.asarray(r.actions)
        save[f"{tag}:rnn_states"] = np.asarray(r.rnn_states).reshape(b, T, n, 64)
        save[f"{tag}:rewards"] = np.asarray(r.rewards)
        save[f"{tag}:costs"] = np.asarray(r.costs)
        es = r.graph.env_states                                               # each rollout has its own reset
        if hasattr(es, "obstacle"):
            for k in ("center", "width", "height", "theta", "points"):
                save[f"{tag}:obs_{k}"] = np.asarray(getattr(es.obstacle, k))[:, 0]                  # static over t
        elif es.obs is not None:
            save[f"{tag}:mpe_obs"] = np.asarray(es.obs)[:, 0]
    save["ro:log_pis"] = np.asarray(ro.log_pis)
    # pre-pass intermediates: the GAE calls' inputs and outputs (dgppo.py:232-237, 268-273), per env
    for tag, calls in (("", gae_calls[:b]), ("_det", gae_calls[b:])):
        save["Vh" + tag] = np.stack([np.asarray(k["Tp1ah_Vh"]) for k, _ in calls])
        save["Qh" + tag] = np.stack([np.asarray(o[0]) for _, o in calls])
        if not tag:
            save["Vl"] = np.stack([np.asarray(k["Tp1_Vl"]) for k, _ in calls])
            save["Ql"] = np.stack([np.asarray(o[1]) for _, o in calls])
    cv = closure_vars(f_pi)
    save["A"] = np.asarray(cv["bcTa_A"]).reshape(b, T, n)
    assert np.array_equal(np.asarray(cv["bcTa_log_pis_old"]).reshape(b, T, n), save["ro:log_pis"])
    assert np.array_equal(np.asarray(closure_vars(f_Vl)["bcT_targets"]).reshape(b, T), save["Ql"])
    assert np.array_equal(np.asarray(closure_vars(f_Vh)["bcTah_Qh_det"]).reshape(save["Qh_det"].shape), save["Qh_det"])
    save["safe_data"] = np.asarray(info["eval/safe_data"], np.float32)

    # losses of the reference's own closures + finite differences along seeded directions
    for tag, f, p in (("Vl", f_Vl, p_Vl), ("Vh", f_Vh, p_Vh), ("policy", f_pi, p_pi)):
        flat = flatten(p)
        for k, v in flat.items():
            save[f"param:{tag}:{k}"] = v

        def loss_at(fl):
            with algoshim.trace_constants(ENTROPY_SEED):
                out = f(unflatten(fl))
            return out if tag == "Vl" else out[0], (None if tag == "Vl" else out[1])
        l0, aux = loss_at(flat)
        save[f"loss:{tag}"] = np.asarray(l0, np.float32)
        if tag == "policy":
            for k in ("policy/clip_frac", "policy/entropy", "policy/total_variation_dist"):
                save["aux:" + k] = np.asarray(aux[k], np.float32)
        fds = []
        for gi_, group in enumerate(GROUPS[tag]):
            d = direction(flat, group, seed=1000 + gi_)
            pn = np.sqrt(sum(float((v.astype(np.float64) ** 2).sum()) for k, v in flat.items() if group in k))
            row = []
            for rel in (0.02, 0.01):                     # step = rel * |parameters of the group|
                h = np.float32(rel * pn)
                lp, _ = loss_at({k: (v + h * d[k]).astype(np.float32) for k, v in flat.items()})
                lm, _ = loss_at({k: (v - h * d[k]).astype(np.float32) for k, v in flat.items()})
                row.append((float(lp) - float(lm)) / (2.0 * float(h)))
            fds.append(row)
            print(f"  {tag:6s} group {group or 'all':20s} |p| {pn:8.3f}  fd {row[0]: .6e} {row[1]: .6e}")
        save[f"fd:{tag}"] = np.array(fds, np.float64)                  # (groups, 2)
        # the same closures evaluated in float64 (algoshim.x64): the loss to ~1e-15 and its directional derivatives
        # by central differences with steps of 1e-6 / 1e-7 / 1e-8 |p| - no round-off floor, no truncation to speak of
        with algoshim.x64():
            flat64 = {k: v.astype(np.float64) for k, v in flat.items()}
            l64, _ = loss_at(flat64)
            save[f"loss64:{tag}"] = np.asarray(l64, np.float64)
            fd64 = []
            for gi_, group in enumerate(GROUPS[tag]):
                d = direction(flat, group, seed=1000 + gi_)
                pn = np.sqrt(sum(float((v.astype(np.float64) ** 2).sum()) for k, v in flat.items() if group in k))
                row = []
                for rel in (1e-6, 1e-7, 1e-8):      # three steps: a ReLU / clip kink inside one of them shows as an outlier
                    h = rel * pn
                    lp, _ = loss_at({k: v + h * d[k].astype(np.float64) for k, v in flat64.items()})
                    lm, _ = loss_at({k: v - h * d[k].astype(np.float64) for k, v in flat64.items()})
                    row.append((float(lp) - float(lm)) / (2.0 * h))
                fd64.append(row)
                print(f"  {tag:6s} group {group or 'all':20s} fd64 " + " ".join(f"{v: .9e}" for v in row))
            save[f"fd64:{tag}"] = np.array(fd64, np.float64)           # (groups, 3)
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, f"ref_update_{name}.npz"), **save)
    print("update", name, {k: float(v) for k, v in info.items()})


if __name__ == "__main__":
    for case in (sys.argv[1:] or list(CASES)):
        run(case)
