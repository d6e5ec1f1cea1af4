"""Generate tests/golden/ref_*.npz by EXECUTING THE REFERENCE'S OWN SOURCE
(/root/reference/dgppo/env/..., dgppo/algo/utils.py) under oracle/jaxshim.py
(a NumPy stand-in for jax; jax itself is not installable in this image).  Where a real
jax / flax / jraph / tfp stack is importable the script uses it instead of the stand-ins
and writes the same files: that run pins the third-party arithmetic the stand-ins restate.

    python tools/gen_golden_from_reference.py          # writes tests/golden/

The fixtures hold, per env config, a few environments x a few steps of: agent /
goal / obstacle state, the action fed to env.step, and everything env.step
returns (all GraphsTuple arrays, reward, cost).  They travel with the repo;
/root/reference does not exist on the GPU box.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
try:                    # a real JAX stack, when one exists: the same script then pins the fixtures against XLA itself
    import flax  # noqa: F401,E402
    import jax  # noqa: E402
    import jraph  # noqa: F401,E402
    import tensorflow_probability.substrates.jax  # noqa: F401,E402
    REAL_JAX = True
except ImportError:     # this image: NumPy stand-ins for jax / flax.linen / jraph / tfp (oracle/jaxshim.py, flaxshim.py)
    from oracle import flaxshim, jaxshim  # noqa: E402,F401
    flaxshim.install()
    import jax  # noqa: E402  (the shim)
    REAL_JAX = False
import jax.numpy as jnp  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

CASES = {   # name -> (module, class, n_agents, n_obs, n_envs, n_steps)
    "LidarSpread_n3_obs3": ("dgppo.env.lidar_env.lidar_spread", "LidarSpread", 3, 3, 6, 6),
    "LidarSpread_n8_obs8": ("dgppo.env.lidar_env.lidar_spread", "LidarSpread", 8, 8, 4, 5),
    "LidarTarget_n5_obs2": ("dgppo.env.lidar_env.lidar_target", "LidarTarget", 5, 2, 4, 5),
    "LidarBicycleTarget_n4_obs3": ("dgppo.env.lidar_env.lidar_bicycle_target", "LidarBicycleTarget", 4, 3, 4, 5),
    "MPESpread_n8_obs3": ("dgppo.env.mpe.mpe_spread", "MPESpread", 8, 3, 4, 5),
    "LidarSpread_n4_obs0": ("dgppo.env.lidar_env.lidar_spread", "LidarSpread", 4, 0, 3, 4),
    "MPETarget_n6_obs3": ("dgppo.env.mpe.mpe_target", "MPETarget", 6, 3, 4, 5),
    "MPECorridor_n5_obs2": ("dgppo.env.mpe.mpe_corridor", "MPECorridor", 5, 2, 4, 5),
    "LidarLine_n4_obs3": ("dgppo.env.lidar_env.lidar_line", "LidarLine", 4, 3, 4, 5),
    "MPELine_n3_obs3": ("dgppo.env.mpe.mpe_line", "MPELine", 3, 3, 4, 5),
    "MPELine_n5_obs3": ("dgppo.env.mpe.mpe_line", "MPELine", 5, 3, 4, 5),
    "MPEFormation_n4_obs3": ("dgppo.env.mpe.mpe_formation", "MPEFormation", 4, 3, 4, 5),
    "MPEConnectSpread_n3_obs1": ("dgppo.env.mpe.mpe_connect_spread", "MPEConnectSpread", 3, 1, 4, 5),
    # BASELINE C4 size (bicycle, n = 16) through the reference's own reset; C5 size (n = 64, obs = 64) from synthetic
    # states (its reset may not terminate at the default area: SURVEY.md appendix B.5), see run_case(synthetic=True)
    "LidarBicycleTarget_n16_obs3": ("dgppo.env.lidar_env.lidar_bicycle_target", "LidarBicycleTarget", 16, 3, 2, 3),
    "LidarSpread_n64_obs64_synth": ("dgppo.env.lidar_env.lidar_spread", "LidarSpread", 64, 64, 1, 2),
}
GRAPH_FIELDS = ("n_node", "n_edge", "nodes", "edges", "states", "receivers", "senders", "node_type")


def run_case(name, mod, cls, n, n_obs, n_envs, n_steps):
    import importlib
    Env = getattr(importlib.import_module(mod), cls)
    params = dict(Env.PARAMS)
    params["n_obs"] = n_obs
    env = Env(num_agents=n, area_size=None, max_step=128, dt=0.03, params=params)
    import zlib
    rng = np.random.default_rng(zlib.crc32(name.encode()))
    out = {k: [] for k in GRAPH_FIELDS}
    out.update(action=[], reward=[], cost=[])
    obst = {k: [] for k in ("center", "width", "height", "theta", "points")}
    mpe_obs = []
    for e in range(n_envs):
        if name.endswith("_synth"):
            # reset's tail on synthetic states (lidar_env/base.py:121-124): uniform positions, no rejection sampling
            from dgppo.env.lidar_env.base import LidarEnvState
            A = env.area_size
            pos = rng.uniform(0, A, (n, 2)).astype(np.float32)
            vel = rng.uniform(-0.5, 0.5, (n, 2)).astype(np.float32)
            gp = np.concatenate([rng.uniform(0, A, (n, 2)), np.zeros((n, 2))], axis=1).astype(np.float32)
            obstacles = env.create_obstacles(jnp.array(rng.uniform(0, A, (n_obs, 2)).astype(np.float32)),
                                             jnp.array(rng.uniform(0.1, 0.3, n_obs).astype(np.float32)),
                                             jnp.array(rng.uniform(0.1, 0.3, n_obs).astype(np.float32)),
                                             jnp.array(rng.uniform(0, 2 * np.pi, n_obs).astype(np.float32)))
            states = jnp.array(np.concatenate([pos, vel], axis=1))
            es0 = LidarEnvState(states, jnp.array(gp), obstacles)
            g = env.get_graph(es0, env.get_lidar_data(states, obstacles))
        else:
            g = env.reset(jax.random.PRNGKey(100 + e))
        es = g.env_states
        if hasattr(es, "obstacle") and es.obstacle is not None:
            for k in obst:
                obst[k].append(np.asarray(getattr(es.obstacle, k)))
        if hasattr(es, "obs") and es.obs is not None:
            mpe_obs.append(np.asarray(es.obs))
        per = {k: [np.asarray(getattr(g, k))] for k in GRAPH_FIELDS}
        acts, rews, costs = [], [], []
        for t in range(n_steps):
            # actions beyond [-1, 1] exercise clip_action; scale grows to push agents around
            a = rng.uniform(-1.3, 1.3, (n, 2)).astype(np.float32)
            g, r, c, done, info = env.step(g, jnp.array(a))
            assert not bool(done)
            for k in GRAPH_FIELDS:
                per[k].append(np.asarray(getattr(g, k)))
            acts.append(a); rews.append(np.asarray(r)); costs.append(np.asarray(c))
        for k in GRAPH_FIELDS:
            out[k].append(np.stack(per[k]))
        out["action"].append(np.stack(acts)); out["reward"].append(np.stack(rews)); out["cost"].append(np.stack(costs))
    save = {k: np.stack(v) for k, v in out.items()}          # (envs, T+1 | T, ...)
    if obst["center"]:
        save.update({"obs_" + k: np.stack(v) for k, v in obst.items()})
    if mpe_obs:
        save["mpe_obs"] = np.stack(mpe_obs)
    save["meta"] = np.array([n, n_obs, n_envs, n_steps])
    np.savez_compressed(os.path.join(OUT, f"ref_{name}.npz"), **save)
    print(name, {k: v.shape for k, v in save.items() if k in ("nodes", "edges", "receivers", "reward", "cost")})


def run_gae():
    from dgppo.algo.utils import compute_dec_ocp_gae
    rng = np.random.default_rng(0)
    cases = {}
    for i, (T, a, nh) in enumerate([(16, 3, 2), (128, 8, 2), (7, 1, 1)]):
        hs = np.clip(rng.normal(-0.7, 0.3, (T, a, nh)), -1, 1).astype(np.float32)
        l = rng.uniform(0, 0.02, T).astype(np.float32)
        Vh = rng.normal(0, 0.5, (T + 1, a, nh)).astype(np.float32)
        Vl = rng.normal(0, 0.5, T + 1).astype(np.float32)
        Qh, Ql = compute_dec_ocp_gae(jnp.array(hs), jnp.array(l), jnp.array(Vh), jnp.array(Vl), 0.99, 0.95)
        for k, v in dict(hs=hs, l=l, Vh=Vh, Vl=Vl, Qh=np.asarray(Qh), Ql=np.asarray(Ql)).items():
            cases[f"c{i}_{k}"] = v
    np.savez_compressed(os.path.join(OUT, "ref_gae.npz"), **cases)
    print("gae", [k for k in cases if k.endswith("Qh")])


def _flatten(tree, pre=""):
    out = {}
    for k, v in tree.items():
        if isinstance(v, dict):
            out.update(_flatten(v, pre + k + "/"))
        else:
            out[pre + k] = np.asarray(v, np.float32)
    return out


def run_nn(only=()):
    """Policy (mode + sample + log_pi), Vh and Vl forward of the reference's own module code
    (nn/gnn.py, nn/mlp.py, nn/rnn.py, algo/module/{policy,value,distribution}.py) under the
    flax / jraph / tfp stand-ins, on graphs taken from the env fixtures."""
    from dgppo.algo.module.policy import PPOPolicy
    from dgppo.algo.module.value import ValueNet
    from dgppo.utils.graph import GraphsTuple
    for name, n, node_dim in (("LidarSpread_n3_obs3", 3, 7), ("LidarBicycleTarget_n4_obs3", 4, 8),
                              ("MPESpread_n8_obs3", 8, 7), ("LidarBicycleTarget_n16_obs3", 16, 8)):
        if only and name not in only:
            continue
        d = np.load(os.path.join(OUT, f"ref_{name}.npz"))
        nominal = GraphsTuple(nodes=jnp.zeros((n, node_dim)), edges=jnp.zeros((n, 4)), states=jnp.zeros((n, 4)),
                              n_node=jnp.array(n), n_edge=jnp.array(n), senders=jnp.arange(n),
                              receivers=jnp.arange(n), node_type=jnp.zeros((n,)), env_states=jnp.zeros((n,)))
        rnn0 = jnp.zeros((1, n, 1, 64))
        pol = PPOPolicy(node_dim=node_dim, edge_dim=4, n_agents=n, action_dim=2, use_rnn=True, rnn_layers=1,
                        gnn_layers=2, gnn_out_dim=64)
        p_pol = pol.dist.init(jax.random.PRNGKey(1), nominal, rnn0, n)
        Vh = ValueNet(node_dim=node_dim, edge_dim=4, n_agents=n, n_out=2, use_rnn=True, gnn_layers=1,
                      gnn_out_dim=64, use_lstm=False, decompose=True, use_global_info=False, n_heads=3)
        p_vh = Vh.net.init(jax.random.PRNGKey(2), nominal, rnn0, n)
        Vl = ValueNet(node_dim=node_dim, edge_dim=4, n_agents=n, use_rnn=True, rnn_layers=1, gnn_layers=2,
                      gnn_out_dim=64, use_lstm=False, decompose=False)
        p_vl = Vl.net.init(jax.random.PRNGKey(3), nominal, jnp.zeros((1, 1, 1, 64)), n)
        rng = np.random.default_rng(7)
        E, T1 = d["nodes"].shape[0], d["nodes"].shape[1]
        sel = [(e, t) for e in range(E) for t in (0, T1 - 1)]
        out = {k: [] for k in ("nodes", "edges", "receivers", "senders", "rnn", "eps", "act_mode", "rnn_out",
                               "act_sample", "log_pi", "vh", "vl_rnn", "vl", "vl_rnn_out")}
        for e, t in sel:
            G = GraphsTuple(nodes=jnp.array(d["nodes"][e, t]), edges=jnp.array(d["edges"][e, t]),
                            states=jnp.array(d["states"][e, t]), n_node=jnp.array(d["n_node"][e, t]),
                            n_edge=jnp.array(d["n_edge"][e, t]), senders=jnp.array(d["senders"][e, t]),
                            receivers=jnp.array(d["receivers"][e, t]), node_type=jnp.array(d["node_type"][e, t]),
                            env_states=None)
            rnn = (rng.standard_normal((1, n, 1, 64)) * 0.5).astype(np.float32)
            a_mode, rnn_out = pol.get_action(p_pol, G, jnp.array(rnn))
            key = jax.random.PRNGKey(1000 + 17 * e + t)
            a_s, lp, rnn_out2 = pol.sample_action(p_pol, G, jnp.array(rnn), key)
            assert np.array_equal(np.asarray(rnn_out), np.asarray(rnn_out2))
            if REAL_JAX:    # recover the draw behind the sample from the distribution itself: eps = (atanh a - loc) / scale
                dist, _ = pol.dist.apply(p_pol, G, jnp.array(rnn), n_agents=n)
                base = dist.distribution.distribution           # Independent -> TanhTransformed -> Normal
                eps = ((np.arctanh(np.asarray(a_s, np.float64)) - np.asarray(base.loc, np.float64))
                       / np.asarray(base.scale, np.float64)).astype(np.float32)
            else:
                eps = jaxshim._gen(key).standard_normal((n, 2)).astype(np.float32)     # the draw Normal.sample made
            vh, _ = Vh.get_value(p_vh, G, jnp.array(rnn))
            vl_rnn = (rng.standard_normal((1, 1, 1, 64)) * 0.5).astype(np.float32)
            vl, vl_rnn_out = Vl.get_value(p_vl, G, jnp.array(vl_rnn))
            for k, v in dict(nodes=d["nodes"][e, t], edges=d["edges"][e, t], receivers=d["receivers"][e, t],
                             senders=d["senders"][e, t], rnn=rnn.reshape(n, 64), eps=eps, act_mode=a_mode,
                             rnn_out=np.asarray(rnn_out).reshape(n, 64), act_sample=a_s, log_pi=lp, vh=vh,
                             vl_rnn=vl_rnn.reshape(64), vl=np.asarray(vl).reshape(()),
                             vl_rnn_out=np.asarray(vl_rnn_out).reshape(64)).items():
                out[k].append(np.asarray(v))
        save = {k: np.stack(v) for k, v in out.items()}
        for tag, tree in (("policy", p_pol), ("vh", p_vh), ("vl", p_vl)):
            for k, v in _flatten(tree).items():
                save[f"param:{tag}:{k}"] = v
        np.savez_compressed(os.path.join(OUT, f"ref_nn_{name}.npz"), **save)
        print("nn", name, save["act_mode"].shape, save["vh"].shape, save["vl"].shape)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    only = sys.argv[1:]                      # optional: names of env cases to (re)generate
    for name, spec in CASES.items():
        if not only or name in only:
            run_case(name, *spec)
    if not only:
        run_gae()
        run_nn()
    elif any(o.startswith("nn:") for o in only):          # e.g. nn:LidarBicycleTarget_n16_obs3
        run_nn([o[3:] for o in only if o.startswith("nn:")])
