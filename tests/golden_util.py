"""Loading of the reference-generated fixtures (tests/golden/ref_*.npz, made by
tools/gen_golden_from_reference.py from the reference's own source)."""
import os

import numpy as np

from oracle import env_np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GRAPH_FIELDS = ("n_node", "n_edge", "nodes", "edges", "states", "receivers", "senders", "node_type")

CASES = {
    "LidarSpread_n3_obs3": env_np.EnvCfg(env_np.LIDAR_SPREAD, n=3, n_obs=3),
    "LidarSpread_n8_obs8": env_np.EnvCfg(env_np.LIDAR_SPREAD, n=8, n_obs=8),
    "LidarTarget_n5_obs2": env_np.EnvCfg(env_np.LIDAR_TARGET, n=5, n_obs=2),
    "LidarBicycleTarget_n4_obs3": env_np.EnvCfg(env_np.LIDAR_BICYCLE_TARGET, n=4, n_obs=3),
    "MPESpread_n8_obs3": env_np.EnvCfg(env_np.MPE_SPREAD, n=8, n_obs=3),
    "LidarSpread_n4_obs0": env_np.EnvCfg(env_np.LIDAR_SPREAD, n=4, n_obs=0),
    "MPETarget_n6_obs3": env_np.EnvCfg(env_np.MPE_TARGET, n=6, n_obs=3),
    "MPECorridor_n5_obs2": env_np.EnvCfg(env_np.MPE_CORRIDOR, n=5, n_obs=2, area=1.0, obs_radius=0.2),
    "LidarLine_n4_obs3": env_np.EnvCfg(env_np.LIDAR_LINE, n=4, n_obs=3),
    "MPELine_n3_obs3": env_np.EnvCfg(env_np.MPE_LINE, n=3, n_obs=3),
    "MPELine_n5_obs3": env_np.EnvCfg(env_np.MPE_LINE, n=5, n_obs=3),
    "MPEFormation_n4_obs3": env_np.EnvCfg(env_np.MPE_FORMATION, n=4, n_obs=3),
    "MPEConnectSpread_n3_obs1": env_np.EnvCfg(env_np.MPE_CONNECT_SPREAD, n=3, n_obs=1, area=1.0, obs_radius=0.25),
    "LidarBicycleTarget_n16_obs3": env_np.EnvCfg(env_np.LIDAR_BICYCLE_TARGET, n=16, n_obs=3),      # BASELINE C4 size
    "LidarSpread_n64_obs64_synth": env_np.EnvCfg(env_np.LIDAR_SPREAD, n=64, n_obs=64),             # BASELINE C5 size
}


def load(name):
    d = np.load(os.path.join(GOLDEN_DIR, f"ref_{name}.npz"))
    cfg = CASES[name]
    obstacles = None
    if "obs_center" in d.files:
        th = d["obs_theta"].astype(np.float32)
        obstacles = dict(center=d["obs_center"], width=d["obs_width"], height=d["obs_height"], theta=th,
                         cos=np.cos(th).astype(np.float32), sin=np.sin(th).astype(np.float32),
                         points=d["obs_points"])
    return cfg, d, obstacles


def graph_at(d, t):
    return {k: d[k][:, t] for k in GRAPH_FIELDS}


NN_CASES = {"LidarSpread_n3_obs3": 7, "LidarBicycleTarget_n4_obs3": 8, "MPESpread_n8_obs3": 7,
            "LidarBicycleTarget_n16_obs3": 8}


def load_nn(name):
    """Reference network fixture: graphs, carries, draws, outputs and the three flax-shaped
    parameter pytrees (rebuilt from the flattened 'param:<net>:<path>' entries)."""
    d = np.load(os.path.join(GOLDEN_DIR, f"ref_nn_{name}.npz"))
    trees = {"policy": {}, "vh": {}, "vl": {}}
    for key in d.files:
        if not key.startswith("param:"):
            continue
        _, tag, path = key.split(":", 2)
        node = trees[tag]
        parts = path.split("/")
        for p_ in parts[:-1]:
            node = node.setdefault(p_, {})
        node[parts[-1]] = d[key]
    graph = {k: d[k] for k in ("nodes", "edges", "receivers", "senders")}
    return CASES[name], d, graph, trees
