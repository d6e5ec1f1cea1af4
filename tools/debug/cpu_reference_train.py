"""DGPPO training WITHOUT the CUDA kernels (debugging aid; nothing here is a product path).

Rollouts, Vl / Vh, the Dec-OCP GAE and the CBF advantage merge come from the NumPy oracle (oracle/*.py, pinned
bit for bit / to fp32 rounding against the reference's own source); the minibatch update is algo/update.py on
the CPU (its losses and gradients are pinned against the reference's `get_loss_` closures:
tests/test_update_reference.py).  The loop therefore shares NO kernel with `train.py`.  If it shows the same
learning curve as the GPU run (profiles/r2_train_lidarspread_n3.log: the deterministic policy saturates
within ~50 updates, recovers after ~3000), the curve is a property of the algorithm as the reference writes it,
not of the kernels.

    python tools/debug/cpu_reference_train.py [updates] [n_env] [seed]      # LidarSpread n = 3, obs = 3, T = 128
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from dgppo_b200.algo import params as P      # noqa: E402
from dgppo_b200.algo import update as U      # noqa: E402
from oracle import algo_np, env_np, nn_np, reset_np      # noqa: E402

F = np.float32
UPDATES = int(sys.argv[1]) if len(sys.argv) > 1 else 60
B = int(sys.argv[2]) if len(sys.argv) > 2 else 128
SEED = int(sys.argv[3]) if len(sys.argv) > 3 else 0
N, N_OBS, T, RNN_STEP = 3, 3, 128, 16
GAMMA, LAM, CLIP, COEF_ENT, ALPHA, CBF_EPS, MAX_NORM = 0.99, 0.95, 0.25, 1e-2, 10.0, 1e-2, 2.0
TRAIN_STEPS = 200000            # cbf schedule of the default run: weight 1 until step 100000

cfg = env_np.EnvCfg(env_np.LIDAR_SPREAD, n=N, n_obs=N_OBS, max_step=T)
rays = env_np.ray_table(cfg.n_rays, cfg.comm_radius)
torch.manual_seed(SEED)
rng = np.random.default_rng(SEED)
trees = {"policy": P.init_policy_params(cfg.node_dim, cfg.edge_dim, 2, 2, seed=SEED),
         "Vl": P.init_value_params(cfg.node_dim, cfg.edge_dim, 1, 2, seed=SEED + 1),
         "Vh": P.init_value_params(cfg.node_dim, cfg.edge_dim, cfg.n_cost, 1, seed=SEED + 2)}
states = {k: U.NetTrainState(trees[k], "cpu", lr) for k, lr in (("policy", 3e-4), ("Vl", 1e-3), ("Vh", 1e-3))}
gi = U.GraphIndex(N, cfg.n_ag, cfg.n_ao, cfg.n_nodes, torch.device("cpu"))
entropy_eps = torch.randn((N, 2))


def reset(keys):
    ag, gl, recs = zip(*[reset_np.reset_states(cfg, int(k))[:3] for k in keys])
    rec = np.stack(recs)
    obstacles = dict(center=rec[..., 0:2], width=rec[..., 2], height=rec[..., 3], theta=rec[..., 4], cos=rec[..., 5],
                     sin=rec[..., 6], points=rec[..., 8:16].reshape(rec.shape[:-1] + (4, 2)))
    agent, goal = np.stack(ag), np.stack(gl)
    return env_np.reset_graph(cfg, agent, goal, obstacles, None, rays), obstacles


def values(tr, ro, det_mode):
    """Vh over the T + 1 graphs with the stored policy carries (final graph: the policy's post-step carry)."""
    b = ro["nodes"].shape[0]
    rnn = ro["rnn_states"]                                        # (b, T + 1, n, 64): carry BEFORE step t at [t]
    carr = rnn[:, 1:T + 1] if det_mode else rnn[:, :T]            # test_rollout stores the post-step carry
    g = {k: ro[k][:, :T].reshape((b * T,) + ro[k].shape[2:]) for k in ("nodes", "edges", "receivers", "senders")}
    Vh = nn_np.vh_forward(tr["Vh"], g, carr.reshape(b * T, N, 64), N).reshape(b, T, N, -1)
    gT = {k: ro[k][:, T] for k in ("nodes", "edges", "receivers", "senders")}
    _, _, h_fin, _ = nn_np.policy_forward(tr["policy"], gT, carr[:, -1], N, eps=None)
    return np.concatenate([Vh, nn_np.vh_forward(tr["Vh"], gT, h_fin, N)[:, None]], axis=1), carr


def main():
    t0 = time.time()
    for step in range(UPDATES):
        tr = {k: {"params": s.numpy_tree()["params"]} for k, s in states.items()}
        g0, obst = reset(rng.integers(0, 2 ** 31 - 1, size=B))
        eps = rng.standard_normal((B, T, N, 2)).astype(F)
        ro = algo_np.rollout(cfg, tr["policy"], g0, obst, eps, T, rays=rays)
        g0d, obstd = reset(rng.integers(0, 2 ** 31 - 1, size=B))
        det = algo_np.rollout(cfg, tr["policy"], g0d, obstd, None, T, rays=rays)
        # ---- pre-pass (dgppo.py:204-273)
        h = np.zeros((B, 64), F)
        Vl = np.zeros((B, T + 1), F)
        for t in range(T + 1):
            Vl[:, t], h = nn_np.vl_forward(tr["Vl"], {k: ro[k][:, t] for k in ("nodes", "edges", "receivers", "senders")}, h, N)
        Vh, carr = values(tr, ro, False)
        Vh_det, carr_det = values(tr, det, True)
        Ql = np.zeros((B, T), F)
        Qh_det = np.zeros((B, T, N, cfg.n_cost), F)
        for i in range(B):
            _, Ql[i] = algo_np.compute_dec_ocp_gae(ro["costs"][i], -ro["rewards"][i], Vh[i], Vl[i], GAMMA, LAM)
            Qh_det[i], _ = algo_np.compute_dec_ocp_gae(det["costs"][i], -det["rewards"][i], Vh_det[i], Vl[i], GAMMA, LAM)
        w = 1.0 * (2 if step >= TRAIN_STEPS // 2 else 1) * (2 if step >= 3 * TRAIN_STEPS // 4 else 1)
        A, deriv, _, safe = algo_np.cbf_advantage(Ql, Vl, Vh, cfg.dt, ALPHA, CBF_EPS, w)
        # ---- minibatch scan: one minibatch of all B envs when batch_size = B * T (the default 128 x 128 = 16384)
        tt = lambda a: torch.tensor(np.ascontiguousarray(a))      # noqa: E731

        def graphs(r):
            a = [tt(r[k][:, :T]) for k in ("nodes", "edges", "receivers", "senders")]
            return U.prep_graphs(a[0].reshape((B * T,) + a[0].shape[2:]), a[1].reshape((B * T,) + a[1].shape[2:]),
                                 a[2].reshape(B * T, -1), a[3].reshape(B * T, -1), gi, torch.float32)
        g, gd = graphs(ro), graphs(det)
        loss = U.loss_Vl(states["Vl"].tree(), g, tt(Ql), gi, 2, RNN_STEP)
        states["Vl"].step(loss, MAX_NORM)
        loss_h = U.loss_Vh(states["Vh"].tree(), gd, tt(carr_det), tt(Qh_det), gi, 1)
        states["Vh"].step(loss_h, MAX_NORM)
        loss_p, info = U.loss_policy(states["policy"].tree(), g, tt(ro["actions"]), tt(ro["log_pis"]), tt(A),
                                     entropy_eps.expand(B, T, N, 2), gi, 2, RNN_STEP, CLIP, COEF_ENT)
        r = states["policy"].step(loss_p, MAX_NORM)
        if step % 5 == 0 or step == UPDATES - 1:
            print(f"step {step:4d} {time.time() - t0:6.0f}s | stoch reward {ro['rewards'].sum(1).mean():7.3f} "
                  f"det reward {det['rewards'].sum(1).mean():7.3f} det unsafe {float((det['costs'].max((1, 2, 3)) >= 1e-6).mean()):.2f} "
                  f"det |a| {np.abs(det['actions']).mean():.3f} | safe {safe.mean():.3f} A mean {A.mean():6.2f} | "
                  f"ent {float(info['policy/entropy']):6.2f} tv {float(info['policy/total_variation_dist']):.3f} "
                  f"gn {float(r['grad_norm']):.2f} Vh loss {float(loss_h):.4f}", flush=True)


if __name__ == "__main__":
    main()
