"""Fixed-data checks of the update machinery: Vl / Vh regression on one rollout must converge; the PPO
surrogate on one rollout with fixed advantages must increase."""
import sys, os
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from dgppo_b200.algo import make_algo, update as U
from dgppo_b200.env import make_env

T = 128
env = make_env("LidarSpread", num_agents=3, num_obs=3, max_step=T)
algo = make_algo("dgppo", env=env, node_dim=env.node_dim, edge_dim=env.edge_dim, state_dim=env.state_dim,
                 action_dim=env.action_dim, n_agents=3, batch_size=16384, seed=0)
ro = algo.collect(algo.params, np.arange(128))
pp = algo.prepass(ro, 0)
b = 128
d = env.graph_dims()
gi = U.GraphIndex(3, d.n_ag, d.n_ao, d.n_nodes, algo.device)
arrays = algo._record_arrays(ro)
ix = torch.arange(b, device=algo.device)
g = U.chunk_graphs(arrays, ix, T, gi, torch.float32)
st = algo._train_state("Vl")
print("Ql mean/std", float(pp["bT_Ql"].mean()), float(pp["bT_Ql"].std()), "Vl", float(pp["bTp1_Vl"].mean()), float(pp["bTp1_Vl"].std()))
for it in range(201):
    loss = U.loss_Vl(st["tree"], g, pp["bT_Ql"], gi, 2, 16)
    r = U.clip_and_step(st["opt"], st["leaves"], loss, 2.0)
    if it % 40 == 0:
        print("Vl fit", it, float(loss), float(r["grad_norm"]))
stp = algo._train_state("policy")
A = pp["bTa_A"]
A = -(pp["bT_Ql"] - pp["bTp1_Vl"][:, :-1])
A = ((A - A.mean(1, keepdim=True)) / (A.std(1, keepdim=True) + 1e-8))[:, :, None].expand(b, T, 3).contiguous()
eps = torch.randn(ro.actions.shape, device=algo.device)
for it in range(41):
    loss, info = U.loss_policy(stp["tree"], g, ro.actions, ro.log_pis, A, eps, gi, 2, 16, 0.25, 0.0)
    r = U.clip_and_step(stp["opt"], stp["leaves"], loss, 2.0)
    if it % 5 == 0:
        print("pi fit", it, "surrogate loss", float(loss), "gn", float(r["grad_norm"]), "clip", float(info["policy/clip_frac"]),
              "tv", float(info["policy/total_variation_dist"]), "ent", float(info["policy/entropy"]))
