import os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from dgppo_b200.algo import make_algo
from dgppo_b200.env import make_env
b, n, T = 512, 8, 128
env = make_env("LidarSpread", num_agents=n, num_obs=8, max_step=T)
for splitk, tf32 in (("0", "0"), ("1", "0"), ("1", "1"), ("0", "0"), ("1", "0")):
    os.environ["DGPPO_UPDATE_SPLITK"], os.environ["DGPPO_UPDATE_TF32"] = splitk, tf32
    algo = make_algo("dgppo", env=env, node_dim=7, edge_dim=4, state_dim=4, action_dim=2, n_agents=n, batch_size=16384)
    ro = algo.collect(algo.params, np.arange(b, dtype=np.uint64))
    for _ in range(2):
        algo.update(ro, 0)
    torch.cuda.synchronize(); ts = []
    for _ in range(3):
        t0 = time.perf_counter(); info = algo.update(ro, 0); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    print(f"splitk {splitk} tf32 {tf32}: update {min(ts):7.1f} ms (4 minibatches)  policy/loss {info['policy/loss']:.5f} Vl/loss {info['Vl/loss']:.6f}", flush=True)
    del algo, ro; torch.cuda.empty_cache()
