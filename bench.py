#!/usr/bin/env python
"""bench.py - rollout agent-steps/s of the DGPPO hot path on B200.

    python bench.py --gpus N --steps K --warmup W          # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W  # CPU arm (oracle port)

One "step" is one full rollout (algo.collect: T = 128 env steps) of the
workload's environments: BASELINE.json configs[2], LidarSpread n=8 obs=8 32
rays, 4096 envs IN TOTAL, sharded over the ranks ("scaling": "strong": 4096 /
N envs per GPU - the curve BASELINE.json names; envs are independent, no
data-path collective).  `--scaling weak` keeps 4096 envs per GPU instead; for
N > 1 the default run measures that too and reports it under "weak_scaling".
`value` = agent-steps of all ranks / max-over-ranks device time, inputs
resident in HBM; `e2e` = the same through algo.collect from pinned HOST state
buffers with the H2D copy of the initial states and the D2H read of
rewards+costs inside the timed region.  Prints ONE JSON line on rank 0.
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


CPU_SAMPLE_ENVS, CPU_SAMPLE_STEPS = 16, 32     # per worker process: a FIXED, stated sample (not grown to a time budget)


def cpu_rollout_rate(w, seed=0, rounds=1):
    """Time the oracle port (NumPy restatement of the reference rollout) on ALL host cores: one worker
    process per core, each rolling out its own shard of CPU_SAMPLE_ENVS environments for CPU_SAMPLE_STEPS
    steps (the same data parallelism the reference's vmap exposes).  The rate is per agent-step, so the
    full workload's time is an EXTRAPOLATION from this sample.  -> agent-steps/s, sample description, workers."""
    import multiprocessing as mp
    from concurrent.futures import ProcessPoolExecutor
    workers = max(1, min(os.cpu_count() or 1, 64))      # one process per core; capped so start-up stays short
    n = w["n"]
    b, T = CPU_SAMPLE_ENVS, CPU_SAMPLE_STEPS
    if w["n"] >= 64:
        b, T = 2, 8                                      # C5 graphs are 60x larger
    with ProcessPoolExecutor(max_workers=workers, mp_context=mp.get_context("spawn")) as pool:
        list(pool.map(_cpu_shard, [(w, 2, 1, 0)] * workers))            # start the workers, import numpy
        t0 = time.perf_counter()
        for r in range(rounds):
            list(pool.map(_cpu_shard, [(w, b, T, seed * 1000 + r * 100 + i) for i in range(workers)]))
        dt = time.perf_counter() - t0
    rate = rounds * workers * b * T * n / dt
    sample = (f"{rounds} round(s) of {workers} processes x {b} envs x {T} steps of {w['env']} n={n} obs={w['obs']} "
              f"({dt:.1f} s wall); extrapolated per agent-step to the full workload")
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
        rate, sample, cores = cpu_rollout_rate(w, seed=i)
        if i >= args.warmup:
            vals.append(rate)
    v = float(np.mean(vals))
    ms = 1e3 * (w["envs"] * T_STEPS * w["n"]) / v
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {w['env']} n={w['n']} obs={w['obs']} 32 rays, "
                               f"{w['envs']} envs in total x T={T_STEPS}",
                   "note": "CPU arm: each step times a fixed bounded sample; value and ms_per_step are EXTRAPOLATED "
                           "from it per agent-step (the full workload was not run on the CPU)"},
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

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    w = WORKLOADS[args.workload]
    total = args.envs or w["envs"]
    b = total // world if args.scaling == "strong" else total
    m = measure(args, w, b, world, rank, local, full=True)
    weak = None
    if world > 1 and args.scaling == "strong":        # the weak curve beside the named (strong) one
        weak = measure(args, w, total, world, rank, local, full=False)
    if rank == 0:
        line = m["line"]
        if weak is not None:
            line["weak_scaling"] = {"value": weak["value"], "unit": UNIT, "ms_per_step": weak["ms_per_step"],
                                    "envs_per_gpu": total, "e2e": weak["e2e"]}
        print(json.dumps(line))
    if world > 1:
        # leave without tearing NCCL down: the 8-rank run of this script sat in destroy_process_group() until the
        # launcher's time limit after the line had been printed (update steps captured in CUDA graphs held the
        # communicator).  Every rank has finished its work here; exit code 0 for torchrun.
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def traffic_record(workload, b):
    """DRAM bytes per launch of each kernel from the committed `ncu --set full` capture of this build
    (profiles/traffic.json, written by tools/ncu_traffic.py together with the commit it was taken at)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(p):
        return None
    with open(p) as f:
        rec = json.load(f)
    if rec.get("workload") != workload or rec.get("envs") != b:
        return None
    return rec


def measure(args, w, b, world, rank, local, full):
    import torch
    import torch.distributed as dist
    from dgppo_b200 import _lib
    from dgppo_b200.algo import make_algo
    from dgppo_b200.env import make_env
    from dgppo_b200.env.envs import LidarEnvState, MPEEnvState, Rectangle, rect_record
    from dgppo_b200.trainer.rollout import RolloutRecord

    dev = torch.device("cuda", local)
    n, T = w["n"], T_STEPS

    env = make_env(w["env"], num_agents=n, num_obs=w["obs"], max_step=T)
    # the update's minibatch: the reference default 16384 steps = 128 envs x T, split evenly over the ranks
    algo = make_algo("dgppo", env=env, node_dim=env.node_dim, edge_dim=env.edge_dim, state_dim=env.state_dim,
                     action_dim=env.action_dim, n_agents=n, batch_size=min(16384, b * T * world), seed=rank)
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

    record = RolloutRecord(env, b, T, dev, stochastic=True, compact=args.record == "compact")
    algo.compact_record = args.record == "compact"
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
        if use_prof:      # per-kernel event timing: one stream, no graph, no overlap between env groups
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
    for _ in range(max(1, args.warmup // 2)):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    units = b * T * n * world
    value = units * args.steps / (ms * 1e-3)
    e2e = {"value": units * args.steps / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
           "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / args.steps}
    if not full:
        _lib.lib().dgppo_prof_destroy(prof)
        return {"value": value, "ms_per_step": ms / args.steps, "e2e": e2e}

    # per-kernel device time: the same rollout once more on ONE stream with CUDA events recorded on
    # that stream around every kernel (dgppo_prof_*); not part of `value`
    ms_prof = timed(lambda: step_resident(True), 1)
    sums = (C.c_float * 4)()
    maxs = (C.c_float * 4)()
    _lib.check(_lib.lib().dgppo_prof_read(prof, sums, maxs), "dgppo_prof_read")
    kern_ms = {k: float(sums[i]) for i, k in enumerate(("policy", "step", "lidar", "graph"))}

    # the same through the full public call: algo.collect(params, keys) = device-side reset (K0 rejection
    # sampler) + LiDAR + graph + rollout.  Reported beside the metric; the metric itself uses synthetic states.
    keys = np.arange(b, dtype=np.uint64) + 7919 * (rank + 1)
    try:
        algo.collect(algo.params, keys, record=record)
        ms_api = timed(lambda: algo.collect(algo.params, keys, record=record), max(1, args.steps // 2))
        ms_api /= max(1, args.steps // 2)
    except RuntimeError as exc:       # e.g. C5 at the default area: the reference's sampler cannot place 64 + 64
        ms_api, api_err = None, str(exc)

    # update (dgppo.py:136-321) on the record just collected: the pre-pass
    # (Vl scan, Vh over all (b, T+1) graphs, Dec-OCP GAE, CBF advantage merge, deterministic rollout) and the
    # PPO minibatch scan with its gradient all-reduce.  Reported beside the metric.
    upd = None
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
        ms_pre = ms_upd = upd_err = None
        try:
            _, ms_pre = ev_time(lambda: algo.prepass(ro, 0))
            _, ms_upd = ev_time(lambda: algo.update(ro, 0))
        except Exception as exc:           # e.g. the deterministic rollout resets through the sampler (C5: infeasible
            upd_err = f"{type(exc).__name__}: {exc}"[:200]       # area); an auxiliary leg must not cost the metric line
        gae_bytes = 4 * (T * n * 2 + T + (T + 1) * n * 2 + (T + 1) + T * n * 2 + T)
        upd = {"ms": {"scan_Vl": ms_vl, "Vh": ms_vh, "gae": ms_gae, "cbf_advantage": ms_cbf,
                      "prepass_total": ms_pre, "update_total": ms_upd},
               "error": upd_err, "graphs": b * (T + 1), "minibatches": max(1, b // max(1, (algo.batch_size // world) // T)),
               "gae_gbs": gae_bytes * b / (ms_gae * 1e-3) / 1e9,
               "note": "components: one pass over the stochastic record; prepass_total = deterministic rollout (incl. "
               "reset) + Vl scan + Vh and GAE on both records + CBF advantage merge (dgppo.py:136-273); "
               "update_total = algo.update(): pre-pass + the PPO minibatch scan (update_Vl, update_Vh, update_policy "
               "with torch-autograd gradients, flat gradient all-reduce over the ranks, clip, Adam)"}
        del Vl, Vh, Qh, Ql, ro

    k3_ms = 0.0
    if args.record == "compact":                # K3 over the whole record in one launch (what a GraphsTuple view costs)
        torch.cuda.synchronize()
        record._materialized = None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        goal_b = goal_d if goal_d.shape[1] == env.num_goals else goal_d[:, :env.num_goals]
        record._goal, record._obstacles = goal_b.contiguous(), (None if lidar else obs_d)
        record.materialize()                    # first call: the caching allocator obtains the output buffers
        record._materialized = None
        torch.cuda.synchronize()
        e0.record()
        record.materialize()
        e1.record()
        torch.cuda.synchronize()
        k3_ms = e0.elapsed_time(e1)
        record._materialized = None
        torch.cuda.empty_cache()
    line = None
    if rank == 0:
        hbm, which = peaks()
        sd, k_top = d.state_dim, (env.params.get("top_k_rays", 0) if lidar else 0)
        n_obs_pts = n * k_top if lidar else w["obs"]
        # ALGORITHMIC bytes per env-step of each kernel
        by = {
            "step": 4 * (n * sd + n * sd + n * 2 + n_obs_pts * (2 if lidar else 4) + n * sd + 1 + n * 2),
            "lidar": 4 * (n * 2 + w["obs"] * 13 + n * k_top * 2) if lidar else 0,
            "graph": 4 * (n * sd + n * sd + n_obs_pts * (2 if lidar else 4) + d.n_nodes * d.node_dim + d.n_edges * 4
                          + d.n_nodes * sd + 2 * d.n_edges + d.n_nodes + 2),
            # the GraphsTuple the policy forward consumes (nodes + edges + recv + send) + rnn in/out +
            # eps + action + log_pi.  (With the compact record the kernel reads the state instead and never
            # materialises the graph: `bytes_moved_compact` below, and `traffic` is accordingly smaller.)
            "policy": 4 * (d.n_nodes * d.node_dim + d.n_edges * 4 + 2 * d.n_edges + 2 * n * 64 + n * 2 + n * 2 + n),
        }
        policy_compact = 4 * (n * sd + env.num_goals * sd + n_obs_pts * (2 if lidar else 4) + 2 * n * 64 + n * 2 + n * 2 + n)
        per_launch_us = {k: kern_ms[k] / T * 1e3 for k in kern_ms}
        fp32_peak = 148 * 128 * 2 * 1.965e9 / 1e12          # FFMA lanes x 2 flop x max SM clock, TFLOP/s
        rk = {}
        if args.record == "compact":        # K3 is not in the loop: time it where it runs, building all b x (T + 1) graphs
            per_launch_us["graph"] = k3_ms * 1e3 / (T + 1)
        for k in ("step", "graph", "lidar", "policy"):
            if by[k] == 0 or per_launch_us[k] <= 0:
                continue
            gbs = by[k] * b / (per_launch_us[k] * 1e-6) / 1e9
            rk[k] = {"bound": "hbm" if k in ("step", "graph") else "fp32 issue", "algorithmic_bytes_per_launch": by[k] * b,
                     "us_per_launch": per_launch_us[k], "achieved_gbs": gbs, "frac_of_hbm": gbs / hbm}
        if args.workload == "C3":       # MAC counts of DESIGN.md section 4 (regrouped GNN 0.22 M + head 0.30 M per env-step)
            flops = 2 * 0.52e6 * b
            rk["policy"]["fp32_tflops"] = flops / (per_launch_us["policy"] * 1e-6) / 1e12
            rk["policy"]["frac_of_fp32_peak"] = rk["policy"]["fp32_tflops"] / fp32_peak
            rk["policy"]["bytes_moved_compact"] = policy_compact * b
            rk["policy"]["note"] = ("head GEMMs run as 3xTF32 on tcgen05 (0.79 M tensor MAC per env-step), the GNN "
                                    "layers on the FFMA pipe; counted here as the fp32 work they replace")
        if upd is not None:
            rk["gae"] = {"bound": "hbm", "achieved_gbs": upd["gae_gbs"], "frac_of_hbm": upd["gae_gbs"] / hbm,
                         "us_per_launch": upd["ms"]["gae"] * 1e3}
        tr = traffic_record(args.workload, b)
        ach = rk["policy"]["achieved_gbs"]
        if args.record == "compact":    # agent state + hits + action + rnn + reward + cost + done + log_pi
            rec_bytes = 4 * (n * sd + (n * k_top * 2 if lidar else 0) + n * 2 + n * 64 + 1 + n * env.n_cost + n) + 1
        else:
            rec_bytes = 4 * (d.n_nodes * d.node_dim + d.n_edges * 4 + d.n_nodes * d.state_dim + 2 * d.n_edges
                             + d.n_nodes + 2 + n * 2 + n * 64 + 1 + n * 2 + n) + 1
        rollout_gbs = rec_bytes * b * world * T / (ms / args.steps * 1e-3) / 1e9
        chunks = algo._n_chunks(b)
        kernels_per_rollout = chunks * (5 if lidar else 4) * T + (2 if lidar else 1)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {w['env']} n={n} obs={w['obs']} 32 rays, "
                                   + (f"{b * world} envs in total, {b} envs/GPU" if args.scaling == "strong"
                                      else f"{b} envs/GPU") + f" x T={T} (one step = one full rollout)",
                       "l2": "record written per step is %.1f GB >> 126 MB L2" % (record.nbytes() / 1e9)},
            "e2e": e2e,
            # our kernels inside the timed region, per rollout and rank: every env group replays its captured
            # graph of T x {gnn_layers, head_tc, env_step, [lidar], build_graph}; the reset graph adds
            # [lidar] + build_graph.  host_launches: what the host submits (graph launches + reset kernels)
            "gpu_launches": args.steps * kernels_per_rollout,
            "host_launches": args.steps * (chunks + (2 if lidar else 1)),
            "clocks": clocks,
            "roofline": {"kernel": "K4a policy forward = gnn_layers_kernel<2> + head_tc_kernel", "bound": "hbm",
                         "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm,
                         "traffic": (tr["policy_pair_bytes"] if tr else None),
                         "traffic_source": (f"profiles/traffic.json (ncu --set full at commit {tr['commit']})" if tr else None),
                         "algorithmic_bytes_per_launch": by["policy"] * b,
                         "peak_source": which,
                         "note": "issue / tensor bound kernel pair reported against HBM as BASELINE's metric asks; "
                                 "roofline_per_kernel has the HBM-bound kernels and the compute view"},
            "roofline_per_kernel": rk,
            "kernel_ms_per_rollout": dict(kern_ms, total_one_stream=ms_prof,
                                          **({"graph_all_slots_one_launch": k3_ms} if args.record == "compact" else {})),
            "rollout_streams": chunks, "record": args.record,
            "api_collect_with_reset": ({"ms_per_step": ms_api, "value": units / (ms_api * 1e-3), "unit": UNIT}
                                       if ms_api is not None else {"unavailable": api_err}),
            "rollout_hbm": {"unique_record_bytes_per_env_step": rec_bytes, "achieved_gbs": rollout_gbs,
                            "frac_of_hbm": rollout_gbs / (hbm * world)},
        }
        if upd is not None:
            line["update"] = upd
        if not args.no_cpu and world == 1:
            rate, sample, cores = cpu_rollout_rate(w, rounds=2)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": sample}
    _lib.lib().dgppo_prof_destroy(prof)
    del record, algo
    torch.cuda.empty_cache()
    return {"line": line, "value": value, "ms_per_step": ms / args.steps, "e2e": e2e}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", type=str, default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", type=str, default="C3", choices=list(WORKLOADS))
    ap.add_argument("--no-prepass", action="store_true", help="skip the update pre-pass timing (Vl / Vh / GAE / CBF)")
    ap.add_argument("--envs", type=int, default=None, help="envs per GPU (default: the workload's)")
    ap.add_argument("--scaling", type=str, default="strong", choices=["weak", "strong"],
                    help="strong (default): the workload's envs in total, sharded over the ranks; weak: per GPU")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--record", type=str, default=os.environ.get("DGPPO_BENCH_RECORD", "compact"), choices=["compact", "full"],
                    help="rollout record: compact (K3's inputs per slot, graphs on demand) or full (GraphsTuple arrays)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
