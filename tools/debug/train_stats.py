"""Per-update statistics of a short DGPPO training run (debugging aid)."""
import sys, os
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from dgppo_b200.algo import make_algo
from dgppo_b200.env import make_env

T = 128
env = make_env("LidarSpread", num_agents=3, num_obs=3, max_step=T)
kw = {}
for a in sys.argv[1:]:
    k, v = a.split("=")
    kw[k] = float(v)
algo = make_algo("dgppo", env=env, node_dim=env.node_dim, edge_dim=env.edge_dim, state_dim=env.state_dim,
                 action_dim=env.action_dim, n_agents=3, batch_size=16384, seed=0, **kw)
if os.environ.get("VH_INIT") == "safe":      # diagnostic: start from a Vh that calls every state safe (small negative output)
    t = algo.params["Vh"]
    t["params"]["Dense_0"]["kernel"] = (t["params"]["Dense_0"]["kernel"] * 0.01).astype(np.float32)
    t["params"]["Dense_0"]["bias"] = np.full_like(t["params"]["Dense_0"]["bias"], -0.5)
    algo.set_params("Vh", t)
    algo.invalidate("Vh")
if os.environ.get("FORCE_SAFE") == "1":       # plain PPO on the reward: A = -normalised(Ql - Vl)
    orig = algo.cbf_advantage
    def patched(Ql, Vl, Vh, step):
        A, deriv, acbf, safe = orig(Ql, Vl, Vh, step)
        Al = Ql - Vl[:, :-1]
        Al = (Al - Al.mean(1, keepdim=True)) / (Al.std(1, keepdim=True, unbiased=False) + 1e-8)
        if os.environ.get("A_MODE") == "random":      # pure noise advantages: what does the optimiser do with them?
            return torch.randn_like(A), deriv, acbf, torch.ones_like(safe)
        if os.environ.get("A_MODE") == "random_t":    # noise shared by the agents of a step (as the real Al is)
            return torch.randn_like(A[:, :, :1]).expand_as(A).contiguous(), deriv, acbf, torch.ones_like(safe)
        sgn = float(os.environ.get("A_SIGN", "-1")); return (sgn * Al)[:, :, None].expand_as(A).contiguous(), deriv, acbf, torch.ones_like(safe)
    algo.cbf_advantage = patched
rng = np.random.default_rng(0)
WARM = int(os.environ.get("WARM", "0"))
PRINT = int(os.environ.get("PRINT", "4"))
for step in range(int(os.environ.get("STEPS", "40"))):
    if WARM:
        algo._train_state("policy")["opt"].lr = 0.0 if step < WARM else algo._lrs["policy"]
    keys = rng.integers(0, 2**31 - 1, size=128)
    ro = algo.collect(algo.params, keys)
    info = algo.update(ro, step)
    pp = algo.last_prepass
    if step % PRINT == 0:
        c = ro.costs
        A = pp["bTa_A"]
        print(f"step {step}: reward {float(ro.rewards.sum(1).mean()):.3f} cost>0 {float((c > 0).float().mean()):.3f} "
              f"Vh mean {float(pp['bTp1ah_Vh'].mean()):.3f} std {float(pp['bTp1ah_Vh'].std()):.3f} "
              f"Qh {float(pp['bTah_Qh'].mean()):.3f} deriv mean {float(pp['bTah_cbf_deriv'].mean()):.2f} "
              f"safe {float(pp['bTa_is_safe'].float().mean()):.3f} A mean {float(A.mean()):.2f} std {float(A.std()):.2f} "
              f"min {float(A.min()):.1f} | ent {info['policy/entropy']:.2f} clip {info['policy/clip_frac']:.2f} "
              f"tv {info['policy/total_variation_dist']:.3f} gn {info['policy/grad_norm']:.2f} "
              f"|a| {float(ro.actions.abs().mean()):.2f} logpi {float(ro.log_pis.mean()):.2f}")
        det = pp["det_rollout"]
        print(f"      det: cost mean {float(det.costs.mean()):.3f} cost>0 {float((det.costs > 0).float().mean()):.3f} "
              f"Vh_det {float(pp['bTp1ah_Vh_det'].mean()):.3f} Qh_det {float(pp['bTah_Qh_det'].mean()):.3f} "
              f"stoch cost mean {float(c.mean()):.3f}  Vh/loss {info['Vh/loss_Vh']:.4f} "
              f"Vh[t=0] {float(pp['bTp1ah_Vh'][:, 0].mean()):.3f} Vh[t=64] {float(pp['bTp1ah_Vh'][:, 64].mean()):.3f} Vh[T] {float(pp['bTp1ah_Vh'][:, -1].mean()):.3f} "
              f"Qh[t=0] {float(pp['bTah_Qh'][:, 0].mean()):.3f} Qh[64] {float(pp['bTah_Qh'][:, 64].mean()):.3f} Qh[T-1] {float(pp['bTah_Qh'][:, -1].mean()):.3f} "
              f"hmax {float(c.amax(-1).mean()):.3f}")
