"""Kernel-level profile of the PPO minibatch step (algo/update.py) on the GPU: time of one `algo.update` over 4
minibatches of 128 envs x 128 steps (C3 shapes: LidarSpread n = 8, obs = 8; the reference default batch_size 16384)
for both formulations of the GraphTransformer layer (DGPPO_UPDATE_GNN = dense | regrouped), eager and CUDA-graph
replay, then the torch.profiler table of the eager regrouped step (which library kernels the time goes to)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile

from dgppo_b200.algo import make_algo
from dgppo_b200.env import make_env

b, n, T = 512, 8, 128
env = make_env("LidarSpread", num_agents=n, num_obs=8, max_step=T)


def fresh():
    algo = make_algo("dgppo", env=env, node_dim=7, edge_dim=4, state_dim=4, action_dim=2, n_agents=n, batch_size=16384)
    return algo, algo.collect(algo.params, np.arange(b, dtype=np.uint64))


for gnn_mode in ("dense", "regrouped"):
    for graph in ("0", "1"):
        os.environ["DGPPO_UPDATE_GNN"], os.environ["DGPPO_UPDATE_GRAPH"] = gnn_mode, graph
        algo, ro = fresh()
        for _ in range(2):
            algo.update(ro, 0)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        info = algo.update(ro, 0)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) * 1e3
        pre = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        pre[0].record(); algo.prepass(ro, 0); pre[1].record(); torch.cuda.synchronize()
        print(f"GNN {gnn_mode:9s} graph replay {graph}: update {dt:7.1f} ms (4 minibatches; pre-pass {pre[0].elapsed_time(pre[1]):.1f} ms of it)"
              f"  policy/loss {info['policy/loss']:.5f} Vl/loss {info['Vl/loss']:.6f}", flush=True)
        del algo, ro
        torch.cuda.empty_cache()
os.environ["DGPPO_UPDATE_GNN"], os.environ["DGPPO_UPDATE_GRAPH"] = "regrouped", "0"
algo, ro = fresh()
for _ in range(2):
    algo.update(ro, 0)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    algo.update(ro, 0)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=70))
