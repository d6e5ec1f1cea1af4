"""Tiny driver for ncu: a few launches of each hot kernel on synthetic C3-like
states (no oracle involved).  usage: python tools/profile_kernels.py [workload] [envs] [T]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from dgppo_b200.algo import make_algo  # noqa: E402
from dgppo_b200.env import make_env  # noqa: E402
from dgppo_b200.env.envs import LidarEnvState, MPEEnvState, Rectangle, rect_record  # noqa: E402


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "C3"
    w = bench.WORKLOADS[wl]
    b = int(sys.argv[2]) if len(sys.argv) > 2 else w["envs"]
    T = int(sys.argv[3]) if len(sys.argv) > 3 else 4
    n = w["n"]
    dev = torch.device("cuda", 0)
    env = make_env(w["env"], num_agents=n, num_obs=w["obs"], max_step=T)
    algo = make_algo("dgppo", env=env, node_dim=env.node_dim, edge_dim=env.edge_dim, state_dim=env.state_dim,
                     action_dim=env.action_dim, n_agents=n, batch_size=b * T)
    rng = np.random.default_rng(0)
    A = env.area_size
    agent = np.concatenate([rng.uniform(0, A, (b, n, 2)), rng.uniform(-.5, .5, (b, n, env.state_dim - 2))], -1)
    goal = np.zeros((b, n, env.state_dim)); goal[..., :2] = rng.uniform(0, A, (b, n, 2))
    agent_d = torch.as_tensor(agent, dtype=torch.float32, device=dev)
    goal_d = torch.as_tensor(goal, dtype=torch.float32, device=dev)
    if w["env"].startswith("Lidar"):
        rec = rect_record(rng.uniform(0, A, (b, w["obs"], 2)).astype(np.float32),
                          rng.uniform(0.1, 0.3, (b, w["obs"])).astype(np.float32),
                          rng.uniform(0.1, 0.3, (b, w["obs"])).astype(np.float32),
                          rng.uniform(0, 2 * np.pi, (b, w["obs"])).astype(np.float32))
        es = LidarEnvState(agent_d, goal_d, Rectangle.from_record(rec, dev))
        g0 = env.get_graph(es, env.get_lidar_data(agent_d, es.obstacle))
    else:
        o = np.zeros((b, w["obs"], 4)); o[..., :2] = rng.uniform(0.15, A - 0.15, (b, w["obs"], 2))
        g0 = env.get_graph(MPEEnvState(agent_d, goal_d, torch.as_tensor(o, dtype=torch.float32, device=dev)))
    eps = torch.randn((b, T, n, 2), device=dev)
    ro = algo.collect(algo.params, None, eps=eps, graph0=g0)
    torch.cuda.synchronize()
    print("ok", tuple(ro.actions.shape), float(ro.rewards.mean()))


if __name__ == "__main__":
    main()
