import time, sys, os
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from dgppo_b200.algo import make_algo
from dgppo_b200.env import make_env
b, n, T = 4096, 8, 128
env = make_env("LidarSpread", num_agents=n, num_obs=8, max_step=T)
algo = make_algo("dgppo", env=env, node_dim=7, edge_dim=4, state_dim=4, action_dim=2, n_agents=n, batch_size=16384)
keys = np.arange(b, dtype=np.uint64)
ro = algo.collect(algo.params, keys)
torch.cuda.synchronize()
def tm(name, fn, reps=2):
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); out = fn(); torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) * 1e3
    print(f"{name:28s} {dt:8.1f} ms"); return out
tm("packed policy", lambda: algo.packed("policy", None))
tm("packed Vh", lambda: algo.packed("Vh", None))
det = tm("det_rollout_fn", lambda: algo.det_rollout_fn(algo.params, keys))
Vl, _ = tm("scan_Vl", lambda: algo.scan_Vl(ro))
Vh = tm("_value_record Vh", lambda: algo._value_record("Vh", ro, None))
tm("_record_arrays", lambda: algo._record_arrays(ro))
tm("_record_arrays det", lambda: algo._record_arrays(det))
tm("_value_record Vh det", lambda: algo._value_record("Vh", det, None))
tm("gae", lambda: algo.gae(ro.costs, -ro.rewards, Vh, Vl))
tm("update", lambda: algo.update(ro, 0))

# ---- update() body, step by step (dgppo_b200/algo/dgppo.py: update)
import numpy as np
print("--- update body")
key = algo._np_rng.integers(0, 2 ** 31 - 1, size=b)
det_rollout = tm("det_rollout_fn", lambda: algo.det_rollout_fn(algo.params, key), 1)
Vl, Vl_carries = tm("scan_Vl", lambda: algo.scan_Vl(ro), 1)
Vh = tm("Vh", lambda: algo._value_record("Vh", ro, None), 1)
Qh, Ql = tm("gae", lambda: algo.gae(ro.costs, -ro.rewards, Vh, Vl), 1)
A = tm("cbf", lambda: algo.cbf_advantage(Ql, Vl, Vh, 0), 1)
Vh_det = tm("Vh det", lambda: algo._value_record("Vh", det_rollout, None), 1)
tm("gae det", lambda: algo.gae(det_rollout.costs, -det_rollout.rewards, Vh_det, Vl), 1)
def host():
    idx = np.arange(b); algo._np_rng.shuffle(idx)
    rnn_chunk_ids = np.array(np.array_split(np.arange(T), T // algo.rnn_step))
    batch_idx = np.array(np.array_split(idx, idx.shape[0] // (algo.batch_size // T)))
    return float(A[3].float().mean())
tm("host part", host, 1)
print("--- update() repeated")
for i in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter(); algo.update(ro, 0); t1 = time.perf_counter()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"update call {i}: host returned after {(t1 - t0) * 1e3:.1f} ms, device done after {(t2 - t0) * 1e3:.1f} ms")
import cProfile, pstats
pr = cProfile.Profile(); pr.enable(); algo.update(ro, 0); torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
