#!/usr/bin/env python
"""bench.py - rollout agent-steps/s of the DGPPO hot path on B200.

    python bench.py --gpus N --steps K --warmup W          # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W  # CPU arm (oracle port)

One "step" is one full rollout (algo.collect: T = 128 env steps) of the
workload's environments on each rank: BASELINE.json configs[2], LidarSpread
n=8 obs=8 32 rays, 4096 envs per GPU ("scaling": "weak"; envs are independent,
no data-path collective).  `value` = agent-steps of all ranks / max-over-ranks
device time, inputs resident in HBM; `e2e` = the same through algo.collect from
pinned HOST state buffers with the H2D copy of the initial states and the D2H
read of rewards+costs inside the timed region.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    "C1": dict(env="LidarSpread", n=3, obs=3, envs=4096),
    "C2": dict(env="MPESpread", n=8, obs=3, envs=4096),
    "C3": dict(env="LidarSpread", n=8, obs=8, envs=4096),
    "C4": dict(env="LidarBicycleTarget", n=16, obs=3, envs=2048),
    "C5": dict(env="LidarSpread", n=64, obs=64, envs=1024),
}
T_STEPS = 128
METRIC = "rollout agent-steps/sec (LidarSpread n=8, 4096 envs)"
UNIT = "agent-steps/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


def oracle_cfg(w):
    from oracle import env_np
    return env_np.EnvCfg(env_np.KIND_BY_NAME[w["env"]], n=w["n"], n_obs=w["obs"])


# ------------------------------------------------------------------ CPU arm
def _cpu_shard(job):
    """One worker process: the oracle rollout of `b` environments for `T` steps, BLAS pinned to 1 thread."""
    w, b, T, seed = job
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=1)
    except Exception:
        pass
    from dgppo_b200.algo import params as P
    from oracle import algo_np, env_np
    cfg = oracle_cfg(w)
    tree = P.init_policy_params(cfg.node_dim, 4, 2, 2, seed=0)
    agent, goal, obst, mo = env_np.synthetic_states(cfg, b, seed)
    g0 = env_np.reset_graph(cfg, agent, goal, obst, mo)
    eps = np.random.default_rng(seed).standard_normal((b, T, cfg.n, 2)).astype(np.float32)
    t0 = time.perf_counter()
    algo_np.rollout(cfg, tree, g0, obst, eps, T)
    return time.perf_counter() - t0


def cpu_rollout_rate(w, target_s=12.0, seed=0):
    """Time the oracle port (NumPy restatement of the reference rollout) on ALL host cores: one worker
    process per core, each rolling out its own shard of environments (the same data parallelism the
    reference's vmap exposes), on a bounded sample of the workload grown until a round takes
    ~target_s.  -> agent-steps/s (all workers, wall clock), sample description, workers."""
    import multiprocessing as mp
    from concurrent.futures import ProcessPoolExecutor
    workers = max(1, min(os.cpu_count() or 1, 64))      # one process per core; capped so start-up stays short
    n = w["n"]
    b, T = 8, 4
    rate, sample = 0.0, ""
    with ProcessPoolExecutor(max_workers=workers, mp_context=mp.get_context("spawn")) as pool:
        list(pool.map(_cpu_shard, [(w, 2, 1, 0)] * workers))            # start the workers, import numpy
        while True:
            jobs = [(w, b, T, seed * 1000 + i) for i in range(workers)]
            t0 = time.perf_counter()
            list(pool.map(_cpu_shard, jobs))
            dt = time.perf_counter() - t0
            rate = workers * b * T * n / dt
            sample = (f"{workers} processes x {b} envs x {T} steps of {w['env']} n={n} obs={w['obs']} "
                      f"({dt:.1f} s wall)")
            if dt >= target_s * 0.5 or b * T >= 64 * 32:
                break
            grow = min(8.0, max(2.0, target_s / max(dt, 1e-3)))
            if T < 32:
                T = int(min(32, T * 2)); grow /= 2
            b = int(min(64, max(b + 1, b * grow)))
    return rate, sample, workers


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    vals = []
    sample = ""
    for i in range(args.warmup + args.steps):
        rate, sample, cores = cpu_rollout_rate(w, target_s=max(4.0, min(20.0, 60.0 / max(1, args.steps))), seed=i)
        if i >= args.warmup:
            vals.append(rate)
    v = float(np.mean(vals))
    ms = 1e3 * (w["envs"] * T_STEPS * w["n"]) / v
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {w['env']} n={w['n']} obs={w['obs']} 32 rays, "
                               f"{w['envs']} envs/GPU x T={T_STEPS}", "note": "CPU arm runs a bounded sample"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "what": "NumPy restatement of the reference rollout (jax is not installable in this image)"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------ clocks
class ClockSampler(threading.Thread):
    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._halt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._halt.wait(0.05)

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------ GPU arm
def run_gpu(args):
    import torch
    import torch.distributed as dist
    from dgppo_b200 import _lib
    from dgppo_b200.algo import make_algo
    from dgppo_b200.env import make_env
    from dgppo_b200.env.envs import LidarEnvState, MPEEnvState, Rectangle, rect_record
    from dgppo_b200.trainer.rollout import RolloutRecord

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    w = WORKLOADS[args.workload]
    b = args.envs or w["envs"]
    if args.scaling == "strong":
        b = b // world
    n, T = w["n"], T_STEPS

    env = make_env(w["env"], num_agents=n, num_obs=w["obs"], max_step=T)
    algo = make_algo("dgppo", env=env, node_dim=env.node_dim, edge_dim=env.edge_dim, state_dim=env.state_dim,
                     action_dim=env.action_dim, n_agents=n, batch_size=min(16384, b * T), seed=rank)
    d = env.graph_dims()

    # synthetic reset states, one independent shard per rank, in PINNED host memory
    rng = np.random.default_rng(1000 + rank)
    A = env.area_size
    pos = rng.uniform(0, A, (b, n, 2)).astype(np.float32)
    if env.state_dim == 5:
        th = rng.uniform(0, 2 * np.pi, (b, n)).astype(np.float32)
        rest = np.stack([np.cos(th), np.sin(th), rng.uniform(-0.5, 0.5, (b, n)).astype(np.float32)], -1)
    else:
        vmax = 1.0 if w["env"].startswith("MPE") else 0.5
        rest = rng.uniform(-vmax, vmax, (b, n, 2)).astype(np.float32)
    agent_h = torch.from_numpy(np.concatenate([pos, rest], -1).astype(np.float32)).pin_memory()
    goal_np = np.zeros((b, n, env.state_dim), np.float32)
    goal_np[..., :2] = rng.uniform(0, A, (b, n, 2))
    goal_h = torch.from_numpy(goal_np).pin_memory()
    lidar = w["env"].startswith("Lidar")
    if lidar:
        rec = rect_record(rng.uniform(0, A, (b, w["obs"], 2)).astype(np.float32),
                          rng.uniform(0.1, 0.3, (b, w["obs"])).astype(np.float32),
                          rng.uniform(0.1, 0.3, (b, w["obs"])).astype(np.float32),
                          rng.uniform(0, 2 * np.pi, (b, w["obs"])).astype(np.float32))
        obs_h = torch.from_numpy(rec).pin_memory()
    else:
        o = np.zeros((b, w["obs"], 4), np.float32)
        o[..., :2] = rng.uniform(0.15, A - 0.15, (b, w["obs"], 2))
        obs_h = torch.from_numpy(o).pin_memory()
    rew_h = torch.empty((b, T), dtype=torch.float32).pin_memory()
    cost_h = torch.empty((b, T, n, 2), dtype=torch.float32).pin_memory()
    h2d = agent_h.numel() * 4 + goal_h.numel() * 4 + obs_h.numel() * 4
    d2h = rew_h.numel() * 4 + cost_h.numel() * 4

    record = RolloutRecord(env, b, T, dev, stochastic=True)
    prof = _lib.lib().dgppo_prof_create(T)

    def reset_graph(agent_d, goal_d, obs_d):
        if lidar:
            es = LidarEnvState(agent_d, goal_d, Rectangle.from_record(obs_d, dev))
            return env.get_graph(es, env.get_lidar_data(agent_d, es.obstacle))
        return env.get_graph(MPEEnvState(agent_d, goal_d, obs_d))

    agent_d, goal_d, obs_d = agent_h.to(dev), goal_h.to(dev), obs_h.to(dev)

    def step_resident(use_prof):
        """Hot path with inputs already in HBM: noise draw + reset graph + T-step rollout."""
        g0 = reset_graph(agent_d, goal_d, obs_d)
        eps = torch.randn((b, T, n, 2), device=dev, dtype=torch.float32)
        if use_prof:      # per-kernel event timing: one stream, no overlap between env groups
            chunks, algo.rollout_chunks = algo.rollout_chunks, 1
            try:
                return algo.collect(algo.params, None, eps=eps, graph0=g0, record=record, prof=prof)
            finally:
                algo.rollout_chunks = chunks
        return algo.collect(algo.params, None, eps=eps, graph0=g0, record=record)

    def step_e2e():
        a = agent_h.to(dev, non_blocking=True)
        g = goal_h.to(dev, non_blocking=True)
        o = obs_h.to(dev, non_blocking=True)
        g0 = reset_graph(a, g, o)
        eps = torch.randn((b, T, n, 2), device=dev, dtype=torch.float32)
        ro = algo.collect(algo.params, None, eps=eps, graph0=g0, record=record)
        rew_h.copy_(ro.rewards, non_blocking=True)
        cost_h.copy_(ro.costs, non_blocking=True)
        return ro

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(args.warmup):
        step_resident(False)
    sampler = ClockSampler(local)
    sampler.start()
    ms = timed(lambda: step_resident(False), args.steps)
    clocks = sampler.stop()
    # per-kernel device time: the same rollout once more on ONE stream with CUDA events recorded on
    # that stream around every kernel (dgppo_prof_*); not part of `value`
    ms_prof = timed(lambda: step_resident(True), 1)
    sums = (C.c_float * 4)()
    maxs = (C.c_float * 4)()
    _lib.check(_lib.lib().dgppo_prof_read(prof, sums, maxs), "dgppo_prof_read")
    kern_ms = {k: float(sums[i]) for i, k in enumerate(("policy", "step", "lidar", "graph"))}

    for _ in range(max(1, args.warmup // 2)):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)

    # the same through the full public call: algo.collect(params, keys) = device-side reset (K0 rejection
    # sampler) + LiDAR + graph + rollout.  Reported beside the metric; the metric itself uses synthetic states.
    keys = np.arange(b, dtype=np.uint64) + 7919 * (rank + 1)
    try:
        algo.collect(algo.params, keys, record=record)
        ms_api = timed(lambda: algo.collect(algo.params, keys, record=record), max(1, args.steps // 2))
        ms_api /= max(1, args.steps // 2)
    except RuntimeError as exc:       # e.g. C5 at the default area: the reference's sampler cannot place 64 + 64
        ms_api, api_err = None, str(exc)

    # update pre-pass (dgppo.py:204-273) on the record just collected: Vl scan,
    # Vh over all (b, T+1) graphs, Dec-OCP GAE, CBF advantage merge.  Reported beside the metric.
    prepass = None
    if not args.no_prepass:
        if ms_api is not None:
            ro = algo.collect(algo.params, keys, record=record)
        else:
            ro = step_resident(False)

        def ev_time(fn):
            fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn()
            e1.record()
            torch.cuda.synchronize()
            return out, e0.elapsed_time(e1)

        (Vl, _), ms_vl = ev_time(lambda: algo.scan_Vl(ro))
        Vh, ms_vh = ev_time(lambda: algo._value_record("Vh", ro, None))
        (Qh, Ql), ms_gae = ev_time(lambda: algo.gae(ro.costs, -ro.rewards, Vh, Vl))
        _, ms_cbf = ev_time(lambda: algo.cbf_advantage(Ql, Vl, Vh, 0))
        try:
            _, ms_upd = ev_time(lambda: algo.update(ro, 0))
        except RuntimeError:               # the deterministic rollout resets through the sampler (C5: infeasible area)
            ms_upd = None
        prepass = {"ms": {"scan_Vl": ms_vl, "Vh": ms_vh, "gae": ms_gae, "cbf_advantage": ms_cbf,
                          "update_prepass_total": ms_upd},
                   "graphs": b * (T + 1), "note": "components: one pass over the stochastic record; update_prepass_total = "
                   "algo.update(): deterministic rollout (incl. reset) + Vl scan + Vh and GAE on both records + "
                   "CBF advantage merge (dgppo.py:136-273), no gradient step"}
        del Vl, Vh, Qh, Ql, ro

    units = b * T * n * world
    value = units * args.steps / (ms * 1e-3)
    e2e = units * args.steps / (ms_e2e * 1e-3)

    if rank == 0:
        hbm, which = peaks()
        # dominant kernel: the policy forward (K4a).  Algorithmic bytes per env-step:
        # nodes + edges + recv + send + rnn in/out + eps + action + log_pi
        pol_bytes = 4 * (d.n_nodes * d.node_dim + d.n_edges * 4 + 2 * d.n_edges + 2 * n * 64 + n * 2 + n * 2 + n)
        ach = pol_bytes * b / (kern_ms["policy"] / T * 1e-3) / 1e9
        rec_bytes = 4 * (d.n_nodes * d.node_dim + d.n_edges * 4 + d.n_nodes * d.state_dim + 2 * d.n_edges
                         + d.n_nodes + 2 + n * 2 + n * 64 + 1 + n * 2 + n) + 1
        rollout_gbs = rec_bytes * b * T / (ms / args.steps * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {w['env']} n={n} obs={w['obs']} 32 rays, "
                                   f"{b} envs/GPU x T={T} (one step = one full rollout)",
                       "l2": "record written per step is %.1f GB >> 126 MB L2" % (record.nbytes() / 1e9)},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / args.steps},
            # our kernels inside the timed region, per rollout: every env group (stream) launches
            # T x {gnn_layers, head, env_step, [lidar], build_graph}; the reset graph adds [lidar] + build_graph
            "gpu_launches": args.steps * (algo._n_chunks(b) * (5 if lidar else 4) * T + (2 if lidar else 1)),
            "clocks": clocks,
            "roofline": {"kernel": "K4a policy forward = gnn_layers_kernel<2> + head_kernel_wide", "bound": "hbm",
                         "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm,
                         # dram__bytes_read + dram__bytes_write per launch pair from the committed ncu --set full
                         # capture (C3, 4096 envs): gnn_layers 27.63 MB + head 17.51 MB read, 0.03 MB written
                         "traffic": 45.17e6 if (args.workload == "C3" and b == 4096) else None,
                         "traffic_source": "profiles/r1_all_v6.ncu.txt",
                         "algorithmic_bytes_per_launch": pol_bytes * b,
                         "peak_source": which,
                         "note": "FP32-FFMA bound kernel reported against HBM as BASELINE's metric asks; "
                                 "see DESIGN.md for the compute roofline"},
            "kernel_ms_per_rollout": dict(kern_ms, total_one_stream=ms_prof),
            "rollout_streams": algo._n_chunks(b),
            "api_collect_with_reset": ({"ms_per_step": ms_api, "value": units / (ms_api * 1e-3), "unit": UNIT}
                                       if ms_api is not None else {"unavailable": api_err}),
            "rollout_hbm": {"unique_record_bytes_per_env_step": rec_bytes, "achieved_gbs": rollout_gbs,
                            "frac_of_hbm": rollout_gbs / hbm},
        }
        if prepass is not None:
            line["update_prepass"] = prepass
        if not args.no_cpu and world == 1:
            rate, sample, cores = cpu_rollout_rate(w, target_s=12.0)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": sample}
        print(json.dumps(line))
    _lib.lib().dgppo_prof_destroy(prof)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", type=str, default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", type=str, default="C3", choices=list(WORKLOADS))
    ap.add_argument("--no-prepass", action="store_true", help="skip the update pre-pass timing (Vl / Vh / GAE / CBF)")
    ap.add_argument("--envs", type=int, default=None, help="envs per GPU (default: the workload's)")
    ap.add_argument("--scaling", type=str, default="weak", choices=["weak", "strong"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
