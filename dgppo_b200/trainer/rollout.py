"""Batched rollout on the B200 kernels.

Replaces `rollout` / `test_rollout` (dgppo/trainer/utils.py:22-86) as they
are used by `algo.collect` and `DGPPO.det_rollout_fn` (informarl.py:177-186,
dgppo.py:108-117): one call of `dgppo_rollout` runs the T-step scan for all
`b` environments of this rank and fills a (b, T+1, ...) record in HBM.

`next_graph[t]` is `graph[t+1]` (trainer/utils.py:48-51), so the record holds
T+1 graphs once and exposes `graph` / `next_graph` as two views of it.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

from .. import _lib
from ..env.base import MultiAgentEnv, ptr, stream_ptr
from ..env.envs import LidarEnv, LidarEnvState, MPEEnvState
from ..utils.graph import GraphsTuple
from .data import Rollout

RNN_DIM = 64

_POOLS = {}


def _submit_pool(n: int):
    from concurrent.futures import ThreadPoolExecutor
    if n not in _POOLS:
        _POOLS[n] = ThreadPoolExecutor(max_workers=n, thread_name_prefix="dgppo-rollout")
    return _POOLS[n]


class RolloutRecord:
    """Device buffers of one (b, T+1, ...) rollout record + kernel workspaces."""

    def __init__(self, env: MultiAgentEnv, b: int, T: int, device: torch.device, stochastic: bool,
                 compact: bool = False):
        """compact=True: the record keeps K3's INPUTS per slot (agent states, LiDAR hits) instead of the
        GraphsTuple arrays - 2.9 KB instead of 10.7 KB per env-step at C3; graphs are built
        on demand (`materialize`, `graphs_of`), the kernels read the state record directly (*_from_state)."""
        d = env.graph_dims()
        n = env.num_agents
        self.env, self.b, self.T, self.d, self.n = env, b, T, d, n
        self.compact = compact
        f32 = dict(dtype=torch.float32, device=device)
        i32 = dict(dtype=torch.int32, device=device)
        P = T + 1
        self.agent_rec = self.hits_rec = None
        self.nodes = self.edges = self.states = self.receivers = self.senders = self.node_type = None
        self.n_node = self.n_edge = None
        self._materialized = None
        if compact:
            self.agent_rec = torch.empty((b, P, n, d.state_dim), **f32)
            if isinstance(env, LidarEnv) and d.n_obs_nodes > 0:
                self.hits_rec = torch.empty((b, P, n, env.params["top_k_rays"], 2), **f32)
        else:
            self.nodes = torch.empty((b, P, d.n_nodes, d.node_dim), **f32)
            self.edges = torch.empty((b, P, d.n_edges, 4), **f32)
            self.states = torch.empty((b, P, d.n_nodes, d.state_dim), **f32)
            self.receivers = torch.empty((b, P, d.n_edges), **i32)
            self.senders = torch.empty((b, P, d.n_edges), **i32)
            self.node_type = torch.empty((b, P, d.n_nodes), **i32)
            self.n_node = torch.empty((b, P), **i32)
            self.n_edge = torch.empty((b, P), **i32)
        self.rnn = torch.empty((b, P, n, RNN_DIM), **f32)
        self.actions = torch.empty((b, T, n, 2), **f32)
        self.log_pis = torch.empty((b, T, n), **f32) if stochastic else None
        self.rewards = torch.empty((b, T), **f32)
        self.costs = torch.empty((b, T, n, env.n_cost), **f32)
        self.dones = torch.zeros((b, T), dtype=torch.bool, device=device)
        self.agent_ws = None if compact else torch.empty((2, b, n, d.state_dim), **f32)
        self.hits_ws = self.hits_ws2 = None
        if not compact and isinstance(env, LidarEnv) and d.n_obs_nodes > 0:
            self.hits_ws = torch.empty((b, n, env.params["top_k_rays"], 2), **f32)
            self.hits_ws2 = torch.empty_like(self.hits_ws)      # second buffer: LiDAR look-ahead (dgppo_rollout)
        self._init_graph_state()

    def _init_graph_state(self):
        # captured-rollout state (dgppo_rollout_graph_*): the graph bakes in device pointers, so the inputs
        # that change from call to call (eps, goal, obstacles, packed parameters) live in buffers owned here
        self.graph_inputs: dict = {}
        self._graph = None
        self._graph_key = None

    def persistent(self, name: str, src: torch.Tensor) -> torch.Tensor:
        """Copy `src` into this record's persistent buffer `name` (allocated on first use / shape change)."""
        buf = self.graph_inputs.get(name)
        if buf is None or buf.shape != src.shape or buf.dtype != src.dtype:
            buf = torch.empty_like(src, memory_format=torch.contiguous_format)
            self.graph_inputs[name] = buf
            self._drop_graph()
        buf.copy_(src)
        return buf

    def _drop_graph(self):
        if getattr(self, "_graph", None):
            _lib.lib().dgppo_rollout_graph_destroy(self._graph)
        self._graph, self._graph_key = None, None

    def __del__(self):
        try:
            self._drop_graph()
        except Exception:       # interpreter shutdown
            pass

    def nbytes(self) -> int:
        ts = [self.nodes, self.edges, self.states, self.receivers, self.senders, self.node_type,
              self.n_node, self.n_edge, self.agent_rec, self.hits_rec, self.rnn, self.actions, self.rewards,
              self.costs, self.dones, self.log_pis]
        return sum(t.numel() * t.element_size() for t in ts if t is not None)

    def bytes_per_env_step(self) -> float:
        return self.nbytes() / (self.b * self.T)

    _TENSORS = ("nodes", "edges", "states", "receivers", "senders", "node_type", "n_node", "n_edge", "rnn",
                "actions", "log_pis", "rewards", "costs", "dones", "hits_ws", "hits_ws2", "agent_rec", "hits_rec")

    # ---- compact record: graphs on demand ------------------------------------------------------------
    def graphs_of(self, env_idx: Optional[torch.Tensor], goal: torch.Tensor, obstacles: Optional[torch.Tensor]):
        """Graph arrays (nodes, edges, states, receivers, senders, node_type, n_node, n_edge), each
        (len(env_idx), T+1, ...), of the selected environments (all when None): ONE launch of K3
        (dgppo_build_graph) over the stored states - bit for bit what the full record would hold."""
        assert self.compact
        env, d, n = self.env, self.d, self.n
        ag = self.agent_rec if env_idx is None else self.agent_rec[env_idx]
        m, P = ag.shape[0], self.T + 1
        goal_s = goal if env_idx is None else goal[env_idx]
        gl = goal_s.unsqueeze(1).expand(m, P, *goal_s.shape[1:]).reshape((m * P,) + tuple(goal_s.shape[1:])).contiguous()
        if self.hits_rec is not None:
            hr = self.hits_rec if env_idx is None else self.hits_rec[env_idx]
            ob = hr.reshape(m * P, n, -1, 2).contiguous()
        elif obstacles is not None and d.n_obs_nodes > 0:
            o = obstacles if env_idx is None else obstacles[env_idx]
            ob = o.unsqueeze(1).expand(m, P, *o.shape[1:]).reshape((m * P,) + tuple(o.shape[1:])).contiguous()
        else:
            ob = None
        g = env._graph_kernel(ag.reshape(m * P, n, d.state_dim).contiguous(), gl, ob, None)
        return {k: getattr(g, k).reshape((m, P) + tuple(getattr(g, k).shape[1:]))
                for k in ("nodes", "edges", "states", "receivers", "senders", "node_type", "n_node", "n_edge")}

    def materialize(self):
        """All (b, T+1) graphs of a compact record (cached until the record is overwritten)."""
        if self._materialized is None:
            self._materialized = self.graphs_of(None, self._goal, self._obstacles)
        return self._materialized

    def env_slice(self, lo: int, hi: int) -> "RolloutRecord":
        """A view of environments [lo, hi) of this record (shares memory)."""
        v = object.__new__(RolloutRecord)
        v.env, v.T, v.d, v.n, v.b = self.env, self.T, self.d, self.n, hi - lo
        v.compact, v._materialized = self.compact, None
        for k in self._TENSORS:
            t = getattr(self, k)
            setattr(v, k, None if t is None else t[lo:hi])
        # the ping-pong state workspace must be contiguous per chunk: give the view its own
        v.agent_ws = None if self.compact else torch.empty((2, hi - lo) + tuple(self.agent_ws.shape[2:]),
                                                           dtype=torch.float32, device=self.rnn.device)
        v._init_graph_state()
        return v

    def graph_view(self, lo: int, hi: int, env_states) -> GraphsTuple:
        s = slice(lo, hi)
        return GraphsTuple(self.n_node[:, s], self.n_edge[:, s], self.nodes[:, s], self.edges[:, s],
                           self.states[:, s], self.receivers[:, s], self.senders[:, s],
                           self.node_type[:, s], env_states)


class LazyGraphsTuple:
    """GraphsTuple-shaped view of slots [lo, hi) of a COMPACT record: the array fields are built by K3 from the
    stored states on first access (`RolloutRecord.materialize`, one launch over all slots) and are then ordinary
    tensors; `env_states` is available without building anything.  Iteration / indexing / `_replace` /
    `type_states` behave like the GraphsTuple they stand for (utils/graph.py)."""

    _FIELDS = GraphsTuple._FIELDS

    def __init__(self, record: "RolloutRecord", lo: int, hi: int, env_states):
        self.record, self.lo, self.hi, self.env_states, self.connectivity = record, lo, hi, env_states, None

    def as_tuple(self) -> GraphsTuple:
        m, s = self.record.materialize(), slice(self.lo, self.hi)
        return GraphsTuple(m["n_node"][:, s], m["n_edge"][:, s], m["nodes"][:, s], m["edges"][:, s], m["states"][:, s],
                           m["receivers"][:, s], m["senders"][:, s], m["node_type"][:, s], self.env_states)

    def __getattr__(self, name):
        if name in GraphsTuple._FIELDS or name in ("is_single", "n_graphs", "batch_shape", "type_nodes", "type_states",
                                                   "map_arrays", "_replace", "without_edge"):
            return getattr(self.as_tuple(), name)
        raise AttributeError(name)

    def __iter__(self):
        return iter(self.as_tuple())

    def __getitem__(self, i):
        return self.as_tuple()[i]

    def __len__(self):
        return len(GraphsTuple._FIELDS)


def run_rollout(env: MultiAgentEnv, net_cfg: _lib.DgppoNetCfg, params_dev: torch.Tensor,
                graph0: GraphsTuple, eps: Optional[torch.Tensor], T: int,
                init_rnn_state: Optional[torch.Tensor] = None,
                record: Optional[RolloutRecord] = None, test_mode: bool = False,
                prof=None) -> Rollout:
    """Run T steps from the batched, already-reset `graph0` (b, ...).

    eps: (b, T, n, 2) N(0,1) draws for the stochastic policy (`algo.step`), or
    None for the deterministic one (`algo.act`).  test_mode selects
    `test_rollout`'s convention of emitting the POST-step rnn state
    (trainer/utils.py:73-77) instead of `rollout`'s pre-step one (:50-51).
    """
    b = graph0.nodes.shape[0]
    dev = graph0.nodes.device
    n = env.num_agents
    if record is None:
        record = RolloutRecord(env, b, T, dev, stochastic=eps is not None,
                               compact=os.environ.get("DGPPO_COMPACT", "0") == "1")
    rec, d = record, record.d
    assert rec.b == b and rec.T == T
    es = graph0.env_states

    # slot 0 <- the reset graph; workspaces <- the reset state
    compact = rec.compact
    rec._materialized = None
    if not compact:
        for name in ("nodes", "edges", "states", "receivers", "senders", "node_type", "n_node", "n_edge"):
            getattr(rec, name)[:, 0].copy_(getattr(graph0, name))
    if init_rnn_state is None:
        rec.rnn[:, 0].zero_()
    else:   # (rnn_layers=1, n, n_carries=1, 64) as algo.init_rnn_state (informarl.py:114-124)
        rec.rnn[:, 0].copy_(init_rnn_state.reshape(n, RNN_DIM).to(dev))
    (rec.agent_rec[:, 0] if compact else rec.agent_ws[0]).copy_(es.agent)
    goal = es.goal.contiguous()
    obstacles, rays = None, None
    if isinstance(es, LidarEnvState):
        if (rec.hits_rec if compact else rec.hits_ws) is not None:
            obstacles = es.obstacle.record.contiguous()
            rays = env.ray_dirs(dev)
            k = env.params["top_k_rays"]
            hits0 = graph0.states[:, n + env.num_goals:n + env.num_goals + n * k, :2].reshape(b, n, k, 2)
            (rec.hits_rec[:, 0] if compact else rec.hits_ws).copy_(hits0)
    elif isinstance(es, MPEEnvState) and es.obs is not None:
        obstacles = es.obs.contiguous()
    if eps is not None:
        eps = eps.contiguous()
        assert eps.shape == (b, T, n, 2) and eps.dtype == torch.float32
    assert (eps is None) == (rec.log_pis is None), "record built for the other policy mode"

    # Captured rollout (default): one graph launch instead of 5 T kernel launches.  The per-call inputs are
    # copied into buffers the record owns, whose addresses the graph has baked in.
    use_graph = prof is None and os.environ.get("DGPPO_GRAPH", "1") != "0"
    if use_graph:
        goal = rec.persistent("goal", goal)
        params_dev = rec.persistent("params", params_dev)
        if obstacles is not None:
            obstacles = rec.persistent("obstacles", obstacles)
        if rays is not None:
            rays = rec.persistent("rays", rays)
        if eps is not None:
            eps = rec.persistent("eps", eps)

    buf = _lib.DgppoRolloutBuffers(
        ptr(rec.nodes), ptr(rec.edges), ptr(rec.states), ptr(rec.receivers), ptr(rec.senders),
        ptr(rec.node_type), ptr(rec.n_node), ptr(rec.n_edge), ptr(rec.rnn), ptr(eps),
        ptr(rec.actions), ptr(rec.log_pis) if eps is not None else None, ptr(rec.rewards), ptr(rec.costs),
        ptr(rec.agent_ws), ptr(rec.hits_ws), ptr(goal), ptr(obstacles), ptr(rays), ptr(rec.hits_ws2),
        ptr(rec.agent_rec), ptr(rec.hits_rec))
    rec._goal, rec._obstacles = goal, obstacles        # what graphs_of / the *_from_state kernels need later
    cfg = env.env_cfg()
    if use_graph:
        key = (bytes(cfg), bytes(net_cfg), T, b, os.environ.get("DGPPO_HEAD"), os.environ.get("DGPPO_LIDAR_AHEAD"))
        if rec._graph is None or rec._graph_key != key:
            rec._drop_graph()
            rc = C.c_int32(0)
            # capture on a private stream: wait for the copies above first so nothing of this stream is captured
            handle = _lib.lib().dgppo_rollout_graph_create(C.byref(cfg), C.byref(net_cfg), ptr(params_dev),
                                                           C.byref(buf), T, b, C.byref(rc))
            if not handle:
                _lib.check(rc.value or -1, "dgppo_rollout_graph_create")
            rec._graph, rec._graph_key = handle, key
        _lib.check(_lib.lib().dgppo_rollout_graph_launch(rec._graph, stream_ptr()), "dgppo_rollout_graph_launch")
    else:
        _lib.check(_lib.lib().dgppo_rollout(stream_ptr(), C.byref(cfg), C.byref(net_cfg), ptr(params_dev),
                                             C.byref(buf), T, b, prof), "dgppo_rollout")

    return _as_rollout(env, rec, es, eps is not None, T, test_mode)


def _as_rollout(env, rec, es, stochastic: bool, T: int, test_mode: bool) -> Rollout:
    n = env.num_agents

    def env_view(lo, hi):
        if rec.compact:     # per-slot agent states; goals are static (the reference repeats them per step)
            ag = rec.agent_rec[:, lo:hi]
            gl = es.goal.unsqueeze(1).expand(ag.shape[0], hi - lo, *es.goal.shape[1:])
        else:
            st = rec.states[:, lo:hi]
            ag, gl = st[:, :, :n], st[:, :, n:n + env.num_goals]
        if isinstance(es, LidarEnvState):
            return LidarEnvState(ag, gl, es.obstacle)
        return MPEEnvState(ag, gl, es.obs)
    rnn = rec.rnn[:, 1:] if test_mode else rec.rnn[:, :T]
    if rec.compact:
        return Rollout(graph=LazyGraphsTuple(rec, 0, T, env_view(0, T)), actions=rec.actions,
                       rnn_states=rnn.unsqueeze(2).unsqueeze(4), rewards=rec.rewards, costs=rec.costs, dones=rec.dones,
                       log_pis=rec.log_pis if stochastic else None,
                       next_graph=LazyGraphsTuple(rec, 1, T + 1, env_view(1, T + 1)))
    return Rollout(
        graph=rec.graph_view(0, T, env_view(0, T)),
        actions=rec.actions,
        rnn_states=rnn.unsqueeze(2).unsqueeze(4),           # (b, T, rnn_layers=1, n, n_carries=1, 64)
        rewards=rec.rewards, costs=rec.costs, dones=rec.dones,
        log_pis=rec.log_pis if stochastic else None,
        next_graph=rec.graph_view(1, T + 1, env_view(1, T + 1)))


def _slice_env_states(es, lo, hi):
    if isinstance(es, LidarEnvState):
        ob = es.obstacle
        if ob is not None:
            ob = type(ob)(*[t[lo:hi] for t in ob])
        return LidarEnvState(es.agent[lo:hi], es.goal[lo:hi], ob)
    return MPEEnvState(es.agent[lo:hi], es.goal[lo:hi], None if es.obs is None else es.obs[lo:hi])


def run_rollout_chunked(env, net_cfg, params_dev, graph0, eps, T, init_rnn_state=None, record=None,
                        test_mode=False, n_chunks: int = 2, prof=None) -> Rollout:
    """Same as run_rollout, with the environments split into `n_chunks` contiguous
    groups that run on separate CUDA streams: the groups are independent
    (informarl.py:183-184), so one group's env kernels (K1-K3, small, issue-bound)
    overlap the other group's policy kernels (latency-bound, one CTA per SM)."""
    b = graph0.nodes.shape[0]
    dev = graph0.nodes.device
    if n_chunks <= 1 or b < 2 * n_chunks:
        return run_rollout(env, net_cfg, params_dev, graph0, eps, T, init_rnn_state, record, test_mode, prof)
    if record is None:
        record = RolloutRecord(env, b, T, dev, stochastic=eps is not None,
                               compact=os.environ.get("DGPPO_COMPACT", "0") == "1")
    if not hasattr(record, "_chunks") or len(record._chunks) != n_chunks:
        bounds = [(i * b) // n_chunks for i in range(n_chunks + 1)]
        record._chunks = [(lo, hi, record.env_slice(lo, hi), torch.cuda.Stream(device=dev))
                          for lo, hi in zip(bounds[:-1], bounds[1:])]
    cur = torch.cuda.current_stream(dev)
    start = torch.cuda.Event()
    start.record(cur)
    def submit(i):
        lo, hi, sub, st = record._chunks[i]
        torch.cuda.set_device(dev)
        st.wait_event(start)
        with torch.cuda.stream(st):
            g0 = graph0.map_arrays(lambda t: t[lo:hi])._replace(env_states=_slice_env_states(graph0.env_states, lo, hi))
            run_rollout(env, net_cfg, params_dev, g0, None if eps is None else eps[lo:hi], T, init_rnn_state,
                        sub, test_mode, prof if i == 0 else None)
        done = torch.cuda.Event()
        done.record(st)
        return done

    # DGPPO_ROLLOUT_THREADS=1: one submitting thread per env group (the launch sequence of a group is a C
    # loop that releases the GIL), so the groups' kernels enter their streams side by side instead of one
    # group after the other.  Only matters when submission is slow (e.g. with DGPPO_LIDAR_AHEAD=1, whose
    # cross-stream events triple the host cost per step); measured +-0 otherwise, so off by default.
    if os.environ.get("DGPPO_ROLLOUT_THREADS", "0") == "1":
        pool = _submit_pool(n_chunks)
        dones = list(pool.map(submit, range(n_chunks)))
    else:
        dones = [submit(i) for i in range(n_chunks)]
    for done in dones:
        cur.wait_event(done)
    record._goal = graph0.env_states.goal.contiguous()
    es0 = graph0.env_states
    record._obstacles = (es0.obstacle.record.contiguous() if isinstance(es0, LidarEnvState) and es0.obstacle is not None
                         else (es0.obs.contiguous() if isinstance(es0, MPEEnvState) and es0.obs is not None else None))
    record._materialized = None
    return _as_rollout(env, record, graph0.env_states, eps is not None, T, test_mode)
