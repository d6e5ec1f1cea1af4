"""LidarEnv / MPE environment families on the B200 kernels.

Mirrors dgppo/env/lidar_env/{base,lidar_spread,lidar_target,
lidar_bicycle_target}.py and dgppo/env/mpe/{base,mpe_spread}.py: same class
names, PARAMS, dims, `reset / step / get_graph / get_cost / get_lidar_data /
agent_step_euler / state_lim / action_lim`.  The arithmetic of step, LiDAR and
graph construction runs in libdgppo_b200.so (K1/K2/K3), and so does `reset`'s
obstacle sampling + rejection sampler (K0 `dgppo_reset`; env/utils.py:139-244).
"""
from __future__ import annotations

import ctypes as C
from typing import NamedTuple, Optional, Tuple

import numpy as np
import torch

from .. import _lib
from ..utils.graph import GraphsTuple
from .base import MultiAgentEnv, StepResult, dev_f32, ptr, require_cuda, stream_ptr

F = np.float32


# ------------------------------------------------------------- obstacles
class Rectangle(NamedTuple):
    """Rectangle (dgppo/env/obstacle.py:30-56), batched over leading axes,
    plus the packed device record the kernels read (DGPPO_OBS_STRIDE floats)."""
    type: torch.Tensor
    center: torch.Tensor
    width: torch.Tensor
    height: torch.Tensor
    theta: torch.Tensor
    points: torch.Tensor
    record: torch.Tensor

    @staticmethod
    def create(center, width, height, theta, device=None) -> "Rectangle":
        rec = rect_record(np.asarray(center, F), np.asarray(width, F), np.asarray(height, F),
                          np.asarray(theta, F))
        return Rectangle.from_record(rec, device)

    @staticmethod
    def from_record(rec: np.ndarray, device=None) -> "Rectangle":
        device = require_cuda() if device is None else device
        r = dev_f32(rec, device)
        return Rectangle(torch.zeros(r.shape[:-1] + (1,), device=device), r[..., 0:2], r[..., 2], r[..., 3],
                         r[..., 4], r[..., 8:16].reshape(r.shape[:-1] + (4, 2)), r)

    @property
    def n(self) -> int:
        return self.center.shape[-2]


def rect_record(center, width, height, theta) -> np.ndarray:
    """Rectangle.create (obstacle.py:39-56) in fp32, op by op, into the
    16-float record of include/dgppo_abi.h."""
    c = np.cos(theta).astype(F)
    s = np.sin(theta).astype(F)
    hw = (width / F(2)).astype(F)
    hh = (height / F(2)).astype(F)
    bx = np.stack([hw, -hw, -hw, hw], axis=-1)
    by = np.stack([hh, hh, -hh, -hh], axis=-1)
    px = ((c[..., None] * bx).astype(F) + ((-s)[..., None] * by).astype(F)).astype(F) + center[..., 0:1]
    py = ((s[..., None] * bx).astype(F) + (c[..., None] * by).astype(F)).astype(F) + center[..., 1:2]
    rec = np.zeros(center.shape[:-1] + (_lib.OBS_STRIDE,), F)
    rec[..., 0:2] = center
    rec[..., 2], rec[..., 3], rec[..., 4], rec[..., 5], rec[..., 6] = width, height, theta, c, s
    rec[..., 8:16:2] = px.astype(F)
    rec[..., 9:16:2] = py.astype(F)
    return rec


def _as_seeds(key) -> Tuple[np.ndarray, bool]:
    """`key` may be an int, a sequence / array of ints (one env per entry), or
    a (b, 2) uint32 array shaped like a batch of jax PRNG keys."""
    if isinstance(key, torch.Tensor):
        key = key.detach().cpu().numpy()
    k = np.asarray(key)
    if k.ndim == 0:
        return k.reshape(1).astype(np.uint64), True
    if k.ndim == 2:
        k = (k[:, 0].astype(np.uint64) << np.uint64(32)) | k[:, 1].astype(np.uint64)
    return k.astype(np.uint64), False


def _reset_kernel(env, seeds: np.ndarray, dev, obs_len=(0.0, 0.0), theta=(0.0, 0.0), defer_check=False):
    """K0: sample obstacles / agents / goals for len(seeds) environments on the device."""
    b, n, sd = len(seeds), env.num_agents, env.state_dim
    n_obs = int(env.params.get("n_obs", 0))
    lidar = isinstance(env, LidarEnv)
    keys = torch.from_numpy(seeds.astype(np.int64)).to(dev)           # same 64 bits, signed container
    agent = torch.empty((b, n, sd), dtype=torch.float32, device=dev)
    goal = torch.empty((b, env.num_goals, sd), dtype=torch.float32, device=dev)
    obst = torch.empty((b, n_obs, _lib.OBS_STRIDE if lidar else 4), dtype=torch.float32, device=dev) \
        if n_obs > 0 else None
    draws = torch.empty((b,), dtype=torch.int32, device=dev)
    cfg = env.env_cfg()
    _lib.check(_lib.lib().dgppo_reset(stream_ptr(), C.byref(cfg), ptr(keys), float(obs_len[0]), float(obs_len[1]),
                                       float(theta[0]), float(theta[1]), ptr(agent), ptr(goal), ptr(obst),
                                       ptr(draws), b), "dgppo_reset")
    env._reset_flags = draws
    if not defer_check:
        check_reset(env)
    return agent, goal, obst


def check_reset(env) -> None:
    """Raise if the last device-side reset flagged an environment whose area cannot hold the agents
    (K0 gives up after 16 restarts; the reference's sampler loops for ever: env/utils.py:229-232).
    Reads one flag vector back: callers that want to keep the stream busy (algo.collect) reset with
    the check deferred and call this after the rollout has been enqueued."""
    draws = getattr(env, "_reset_flags", None)
    if draws is None:
        return
    env._reset_flags = None
    if bool((draws < 0).any()):
        n, n_obs = env.num_agents, int(env.params.get("n_obs", 0))
        raise RuntimeError(f"reset: area_size={env.area_size} cannot hold {n} agents + goals"
                           f"{' + ' + str(n_obs) + ' obstacles' if n_obs else ''} at the required spacing "
                           "(the reference's sampler does not terminate for this configuration)")


def _batchify(*ts):
    """Add a leading batch axis to single-env tensors; report whether it was added."""
    return [None if t is None else t.unsqueeze(0) for t in ts]


# ------------------------------------------------------------- Lidar envs
class LidarEnvState(NamedTuple):
    agent: torch.Tensor
    goal: torch.Tensor
    obstacle: Optional[Rectangle]

    @property
    def n_agent(self) -> int:
        return self.agent.shape[-2]


class _KernelEnv(MultiAgentEnv):
    """Shared kernel plumbing of the Lidar and MPE families."""

    AGENT, GOAL, OBS = 0, 1, 2

    @property
    def edge_dim(self) -> int:
        return 4

    @property
    def action_dim(self) -> int:
        return 2

    @property
    def n_cost(self) -> int:
        return 2

    @property
    def cost_components(self) -> Tuple[str, ...]:
        return "agent collisions", "obs collisions"

    def action_lim(self):
        return -torch.ones(2), torch.ones(2)

    # --- K1
    def _step_kernel(self, agent, goal, obs_nodes, action):
        b, n = agent.shape[0], self.num_agents
        dev = agent.device
        nxt = torch.empty_like(agent)
        reward = torch.empty((b,), dtype=torch.float32, device=dev)
        cost = torch.empty((b, n, self.n_cost), dtype=torch.float32, device=dev)
        cfg = self.env_cfg()
        _lib.check(_lib.lib().dgppo_env_step(stream_ptr(), C.byref(cfg), ptr(agent), ptr(goal), ptr(obs_nodes),
                                              ptr(action), ptr(nxt), ptr(reward), ptr(cost), 1, b),
                   "dgppo_env_step")
        return nxt, reward, cost

    # --- K3
    def _graph_kernel(self, agent, goal, obs_nodes, env_states) -> GraphsTuple:
        b = agent.shape[0]
        dev = agent.device
        d = self.graph_dims()
        f32, i32 = dict(dtype=torch.float32, device=dev), dict(dtype=torch.int32, device=dev)
        nodes = torch.empty((b, d.n_nodes, d.node_dim), **f32)
        edges = torch.empty((b, d.n_edges, 4), **f32)
        states = torch.empty((b, d.n_nodes, d.state_dim), **f32)
        recv = torch.empty((b, d.n_edges), **i32)
        send = torch.empty((b, d.n_edges), **i32)
        ntype = torch.empty((b, d.n_nodes), **i32)
        n_node = torch.empty((b,), **i32)
        n_edge = torch.empty((b,), **i32)
        cfg = self.env_cfg()
        _lib.check(_lib.lib().dgppo_build_graph(stream_ptr(), C.byref(cfg), ptr(agent), ptr(goal), ptr(obs_nodes),
                                                 ptr(nodes), ptr(edges), ptr(states), ptr(recv), ptr(send),
                                                 ptr(ntype), ptr(n_node), ptr(n_edge), 1, b),
                   "dgppo_build_graph")
        return GraphsTuple(n_node, n_edge, nodes, edges, states, recv, send, ntype, env_states)

    def _slices(self, graph: GraphsTuple):
        n = self.num_agents
        agent = graph.type_states(type_idx=0, n_type=n).contiguous()
        goal = graph.type_states(type_idx=1, n_type=self.num_goals).contiguous()
        return agent, goal

    @staticmethod
    def _squeeze_graph(g: GraphsTuple, env_states) -> GraphsTuple:
        return g.map_arrays(lambda t: t[0])._replace(env_states=env_states)


class LidarEnv(_KernelEnv):
    """dgppo/env/lidar_env/base.py:35-281."""

    PARAMS = {
        "car_radius": 0.05, "comm_radius": 0.5, "n_rays": 32, "obs_len_range": [0.1, 0.3],
        "n_obs": 3, "default_area_size": 1.5, "dist2goal": 0.01, "top_k_rays": 8,
    }
    KIND = -1

    def __init__(self, num_agents, area_size=None, max_step=128, dt=0.03, params=None):
        area_size = type(self).PARAMS["default_area_size"] if area_size is None else area_size
        super().__init__(num_agents, area_size, max_step, dt, params)
        self.num_goals = self._num_agents
        self._rays = None

    @property
    def state_dim(self) -> int:
        return 4

    @property
    def node_dim(self) -> int:
        return 7

    # theta range of the obstacle sampler differs per env (lidar_env/base.py:106
    # vs lidar_bicycle_target.py:74)
    _OBS_THETA = (0.0, 2 * np.pi)

    def ray_dirs(self, device) -> torch.Tensor:
        """End-point offsets of the beams (env/utils.py:51-55): data for K2."""
        if self._rays is None or self._rays.device != device:
            nb, rng_ = self._params["n_rays"], self._params["comm_radius"]
            th = np.linspace(-np.pi, np.pi - 2 * np.pi / nb, nb).astype(F)
            tab = np.stack([(np.cos(th).astype(F) * F(rng_)).astype(F),
                            (np.sin(th).astype(F) * F(rng_)).astype(F)], axis=-1)
            self._rays = dev_f32(tab, device)
        return self._rays

    def reset(self, key, defer_check: bool = False) -> GraphsTuple:
        """LidarEnv.reset (lidar_env/base.py:89-124): sampling on the device (K0), then LiDAR + graph."""
        dev = require_cuda()
        seeds, single = _as_seeds(key)
        assert self._params["n_obs"] >= 0
        agent, goal, rec = _reset_kernel(self, seeds, dev, self._params["obs_len_range"], self._OBS_THETA,
                                         defer_check=defer_check)
        obstacles = Rectangle.from_record(rec, dev) if rec is not None else None
        env_state = LidarEnvState(agent, goal, obstacles)
        lidar = self.get_lidar_data(env_state.agent, obstacles)
        g = self.get_graph(env_state, lidar)
        if single:
            es = LidarEnvState(env_state.agent[0], env_state.goal[0],
                               None if obstacles is None else Rectangle(*[t[0] for t in obstacles]))
            return self._squeeze_graph(g, es)
        return g

    def get_lidar_data(self, states: torch.Tensor, obstacles: Optional[Rectangle]) -> Optional[torch.Tensor]:
        """get_lidar_data (lidar_env/base.py:126-140) -> (…, n, top_k, 2)."""
        if self._params["n_obs"] == 0:
            return None
        single = states.ndim == 2
        st = states.unsqueeze(0) if single else states
        rec = obstacles.record.unsqueeze(0) if single else obstacles.record
        st, rec = st.contiguous(), rec.contiguous()
        b, n = st.shape[0], self.num_agents
        hits = torch.empty((b, n, self._params["top_k_rays"], 2), dtype=torch.float32, device=st.device)
        cfg = self.env_cfg()
        _lib.check(_lib.lib().dgppo_lidar(stream_ptr(), C.byref(cfg), ptr(st), ptr(rec),
                                           ptr(self.ray_dirs(st.device)), ptr(hits), b), "dgppo_lidar")
        return hits[0] if single else hits

    def agent_step_euler(self, agent_states: torch.Tensor, action: torch.Tensor) -> torch.Tensor:
        """agent_step_euler (lidar_env/base.py:142-149 | lidar_bicycle_target.py:92-111);
        `action` is expected already clipped, as in the reference call site."""
        single = agent_states.ndim == 2
        a, u = (agent_states.unsqueeze(0), action.unsqueeze(0)) if single else (agent_states, action)
        a, u = a.contiguous(), u.contiguous()
        goal = torch.zeros_like(a)
        hits = None
        if self._params["n_obs"] > 0:
            hits = torch.zeros((a.shape[0], self.num_agents, self._params["top_k_rays"], 2),
                               dtype=torch.float32, device=a.device)
        nxt, _, _ = self._step_kernel(a, goal, hits, u)
        return nxt[0] if single else nxt

    def _hits_of(self, graph: GraphsTuple) -> Optional[torch.Tensor]:
        if self._params["n_obs"] == 0:
            return None
        n, k = self.num_agents, self._params["top_k_rays"]
        h = graph.type_states(type_idx=2, n_type=k * n)[..., :2]
        return h.reshape(h.shape[:-2] + (n, k, 2)).contiguous()

    def step(self, graph: GraphsTuple, action: torch.Tensor, get_eval_info: bool = False) -> StepResult:
        """LidarEnv.step (lidar_env/base.py:151-174)."""
        single = graph.is_single
        g = graph.map_arrays(lambda t: t.unsqueeze(0)) if single else graph
        act = (action.unsqueeze(0) if single else action).contiguous().float()
        agent, goal = self._slices(g)
        hits = self._hits_of(g)
        obstacles = graph.env_states.obstacle if self._params["n_obs"] > 0 else None
        nxt, reward, cost = self._step_kernel(agent, goal, hits, act)
        if obstacles is not None and single:
            obstacles_b = Rectangle(*[t.unsqueeze(0) for t in obstacles])
        else:
            obstacles_b = obstacles
        lidar_next = self.get_lidar_data(nxt, obstacles_b)
        next_state = LidarEnvState(nxt, goal, obstacles_b)
        ng = self.get_graph(next_state, lidar_next)
        done = torch.zeros(reward.shape, dtype=torch.bool, device=reward.device)
        if single:
            ng = self._squeeze_graph(ng, LidarEnvState(nxt[0], goal[0], obstacles))
            return StepResult(ng, reward[0], cost[0], done[0], {})
        return StepResult(ng, reward, cost, done, {})

    def get_cost(self, graph: GraphsTuple) -> torch.Tensor:
        """get_cost (lidar_env/base.py:180-207)."""
        single = graph.is_single
        g = graph.map_arrays(lambda t: t.unsqueeze(0)) if single else graph
        agent, goal = self._slices(g)
        zero = torch.zeros(agent.shape[:2] + (2,), dtype=torch.float32, device=agent.device)
        _, _, cost = self._step_kernel(agent, goal, self._hits_of(g), zero)
        return cost[0] if single else cost

    def get_reward(self, graph: GraphsTuple, action: torch.Tensor) -> torch.Tensor:
        """get_reward (lidar_spread.py:35-52 | lidar_target.py:35-52); `action` clipped."""
        single = graph.is_single
        g = graph.map_arrays(lambda t: t.unsqueeze(0)) if single else graph
        act = (action.unsqueeze(0) if single else action).contiguous().float()
        agent, goal = self._slices(g)
        _, reward, _ = self._step_kernel(agent, goal, self._hits_of(g), act)
        return reward[0] if single else reward

    def get_graph(self, state: LidarEnvState, lidar_data: Optional[torch.Tensor] = None) -> GraphsTuple:
        """get_graph (lidar_env/base.py:227-271)."""
        single = state.agent.ndim == 2
        if single:
            agent, goal, lidar = _batchify(state.agent, state.goal, lidar_data)
        else:
            agent, goal, lidar = state.agent, state.goal, lidar_data
        agent, goal = agent.contiguous(), goal.contiguous()
        if self._params["n_obs"] > 0:
            assert lidar is not None, "lidar_data is required when n_obs > 0"
            lidar = lidar.reshape(agent.shape[0], self.num_agents, self._params["top_k_rays"], 2).contiguous()
        else:
            lidar = None
        g = self._graph_kernel(agent, goal, lidar, state)
        return self._squeeze_graph(g, state) if single else g

    def state_lim(self, state=None):
        A = self.area_size
        return torch.tensor([0., 0., -0.5, -0.5]), torch.tensor([A, A, 0.5, 0.5])


class LidarSpread(LidarEnv):
    """dgppo/env/lidar_env/lidar_spread.py."""
    PARAMS = dict(LidarEnv.PARAMS)
    KIND = 0


class LidarTarget(LidarEnv):
    """dgppo/env/lidar_env/lidar_target.py."""
    PARAMS = dict(LidarEnv.PARAMS)
    KIND = 1


class LidarLine(LidarSpread):
    """dgppo/env/lidar_env/lidar_line.py: two landmark nodes; the n goals the reward uses lie evenly on the
    segment between them (landmark2goal, :128-133).  Env kind 6."""
    PARAMS = dict(LidarEnv.PARAMS)
    KIND = 6

    def __init__(self, num_agents, area_size=None, max_step=128, dt=0.03, params=None):
        super().__init__(num_agents, area_size, max_step, dt, params)
        self.num_goals = 2

    def landmark2goal(self, landmarks: torch.Tensor) -> torch.Tensor:
        """(…, 2, 2) landmark positions -> (…, n, 2) goals (lidar_line.py:128-133)."""
        n_int = self.num_agents - 1
        k = torch.arange(0, n_int + 1, device=landmarks.device, dtype=landmarks.dtype)[:, None]
        return landmarks[..., 0:1, :] + k * (landmarks[..., 1:2, :] - landmarks[..., 0:1, :]) / n_int


class LidarBicycleTarget(LidarTarget):
    """dgppo/env/lidar_env/lidar_bicycle_target.py."""
    PARAMS = dict(LidarEnv.PARAMS)
    KIND = 2
    _OBS_THETA = (-np.pi, np.pi)

    @property
    def state_dim(self) -> int:
        return 5      # x, y, cos(theta), sin(theta), v

    @property
    def node_dim(self) -> int:
        return 8

    def state_lim(self, state=None):
        A = self.area_size
        return torch.tensor([0., 0., -1., -1., -0.5]), torch.tensor([A, A, 1., 1., 0.5])


# ------------------------------------------------------------------- MPE
class MPEEnvState(NamedTuple):
    agent: torch.Tensor
    goal: torch.Tensor
    obs: Optional[torch.Tensor]

    @property
    def n_agent(self) -> int:
        return self.agent.shape[-2]


class MPE(_KernelEnv):
    """dgppo/env/mpe/base.py:30-251."""

    PARAMS = {"car_radius": 0.05, "comm_radius": 0.5, "n_obs": 3, "obs_radius": 0.05,
              "default_area_size": 1.0, "dist2goal": 0.01}

    def __init__(self, num_agents, area_size=None, max_step=128, dt=0.03, params=None):
        area_size = type(self).PARAMS["default_area_size"] if area_size is None else area_size
        super().__init__(num_agents, area_size, max_step, dt, params)
        self.num_goals = self._num_agents

    @property
    def state_dim(self) -> int:
        return 4

    @property
    def node_dim(self) -> int:
        return 7

    def reset(self, key, defer_check: bool = False) -> GraphsTuple:
        """MPE.reset (mpe/base.py:81-127): sampling on the device (K0), then the graph."""
        dev = require_cuda()
        seeds, single = _as_seeds(key)
        agent, goal, obs = _reset_kernel(self, seeds, dev, defer_check=defer_check)
        env_state = MPEEnvState(agent, goal, obs)
        g = self.get_graph(env_state)
        if single:
            es = MPEEnvState(agent[0], goal[0], None if obs is None else obs[0])
            return self._squeeze_graph(g, es)
        return g

    def _obs_of(self, graph: GraphsTuple) -> Optional[torch.Tensor]:
        if self._params["n_obs"] == 0:
            return None
        return graph.type_states(type_idx=2, n_type=self._params["n_obs"]).contiguous()

    def agent_step_euler(self, agent_states, action):
        single = agent_states.ndim == 2
        a, u = (agent_states.unsqueeze(0), action.unsqueeze(0)) if single else (agent_states, action)
        a, u = a.contiguous(), u.contiguous()
        obs = None
        if self._params["n_obs"] > 0:
            obs = torch.zeros((a.shape[0], self._params["n_obs"], 4), dtype=torch.float32, device=a.device)
        nxt, _, _ = self._step_kernel(a, torch.zeros_like(a), obs, u)
        return nxt[0] if single else nxt

    def step(self, graph: GraphsTuple, action: torch.Tensor, get_eval_info: bool = False) -> StepResult:
        """MPE.step (mpe/base.py:137-158)."""
        single = graph.is_single
        g = graph.map_arrays(lambda t: t.unsqueeze(0)) if single else graph
        act = (action.unsqueeze(0) if single else action).contiguous().float()
        agent, goal = self._slices(g)
        obs = self._obs_of(g)
        nxt, reward, cost = self._step_kernel(agent, goal, obs, act)
        ng = self.get_graph(MPEEnvState(nxt, goal, obs))
        done = torch.zeros(reward.shape, dtype=torch.bool, device=reward.device)
        if single:
            ng = self._squeeze_graph(ng, MPEEnvState(nxt[0], goal[0], None if obs is None else obs[0]))
            return StepResult(ng, reward[0], cost[0], done[0], {})
        return StepResult(ng, reward, cost, done, {})

    def get_cost(self, graph: GraphsTuple) -> torch.Tensor:
        single = graph.is_single
        g = graph.map_arrays(lambda t: t.unsqueeze(0)) if single else graph
        agent, goal = self._slices(g)
        zero = torch.zeros(agent.shape[:2] + (2,), dtype=torch.float32, device=agent.device)
        _, _, cost = self._step_kernel(agent, goal, self._obs_of(g), zero)
        return cost[0] if single else cost

    def get_reward(self, graph: GraphsTuple, action: torch.Tensor) -> torch.Tensor:
        single = graph.is_single
        g = graph.map_arrays(lambda t: t.unsqueeze(0)) if single else graph
        act = (action.unsqueeze(0) if single else action).contiguous().float()
        agent, goal = self._slices(g)
        _, reward, _ = self._step_kernel(agent, goal, self._obs_of(g), act)
        return reward[0] if single else reward

    def get_graph(self, env_state: MPEEnvState, lidar_data=None) -> GraphsTuple:
        """get_graph (mpe/base.py:211-241)."""
        single = env_state.agent.ndim == 2
        if single:
            agent, goal, obs = _batchify(env_state.agent, env_state.goal, env_state.obs)
        else:
            agent, goal, obs = env_state.agent, env_state.goal, env_state.obs
        obs = obs.contiguous() if (obs is not None and self._params["n_obs"] > 0) else None
        g = self._graph_kernel(agent.contiguous(), goal.contiguous(), obs, env_state)
        return self._squeeze_graph(g, env_state) if single else g

    def state_lim(self, state=None):
        A = self.area_size
        return torch.tensor([0., 0., -1., -1.]), torch.tensor([A, A, 1., 1.])


class MPESpread(MPE):
    """dgppo/env/mpe/mpe_spread.py."""
    PARAMS = {"car_radius": 0.05, "comm_radius": 0.5, "n_obs": 3, "obs_radius": 0.05,
              "default_area_size": 1.5, "dist2goal": 0.01}
    KIND = 3


class MPETarget(MPE):
    """dgppo/env/mpe/mpe_target.py: MPE dynamics / cost / obstacles with one paired goal per agent
    (reward mpe_target.py:32-49, edge blocks :51-80).  First env family widened through the same
    kernels."""
    PARAMS = {"car_radius": 0.05, "comm_radius": 0.5, "n_obs": 3, "obs_radius": 0.05,
              "default_area_size": 1.5, "dist2goal": 0.01}
    KIND = 4


class MPELine(MPESpread):
    """dgppo/env/mpe/mpe_line.py: two landmark nodes, goals on the segment between them (landmark2goal,
    :119-128: interior points for n <= 3, end points included otherwise).  Env kind 7."""
    PARAMS = dict(MPESpread.PARAMS)
    KIND = 7

    def __init__(self, num_agents, area_size=None, max_step=128, dt=0.03, params=None):
        super().__init__(num_agents, area_size, max_step, dt, params)
        self.num_goals = 2

    def landmark2goal(self, landmarks: torch.Tensor) -> torch.Tensor:
        n = self.num_agents
        d = landmarks[..., 1:2, :] - landmarks[..., 0:1, :]
        if n <= 3:
            n_int, k = n + 1, torch.arange(1, n + 1, device=landmarks.device, dtype=landmarks.dtype)[:, None]
        else:
            n_int, k = n - 1, torch.arange(0, n, device=landmarks.device, dtype=landmarks.dtype)[:, None]
        return landmarks[..., 0:1, :] + k * d / n_int


class MPEFormation(MPESpread):
    """dgppo/env/mpe/mpe_formation.py: one landmark node, goals on a circle of radius comm_radius around it
    (landmark2goal, :93-96).  Env kind 8; the circle offsets travel to the kernels as a device table."""
    PARAMS = dict(MPESpread.PARAMS)
    KIND = 8

    def __init__(self, num_agents, area_size=None, max_step=128, dt=0.03, params=None):
        super().__init__(num_agents, area_size, max_step, dt, params)
        self.num_goals = 1
        self._goal_table = None

    def goal_offsets(self) -> np.ndarray:
        th = np.linspace(0, 2 * np.pi, self.num_agents + 1).astype(F)[:-1]
        return (F(self._params["comm_radius"]) * np.stack([np.cos(th).astype(F), np.sin(th).astype(F)], -1)).astype(F)

    def _goal_table_ptr(self):
        if self._goal_table is None:
            self._goal_table = dev_f32(self.goal_offsets(), require_cuda())
        return self._goal_table.data_ptr()

    def landmark2goal(self, landmarks: torch.Tensor, R: Optional[float] = None) -> torch.Tensor:
        off = torch.as_tensor(self.goal_offsets(), device=landmarks.device)
        return landmarks[..., 0:1, :] + off


class MPEConnectSpread(MPESpread):
    """dgppo/env/mpe/mpe_connect_spread.py: MPESpread with one large obstacle, agents and goals sampled as
    connected groups on either side of it, the y range doubled, obstacle edges always on and a THIRD cost:
    connectivity (max over agents of nearest-neighbour distance - connect_radius, :116-118).  Env kind 9."""
    PARAMS = {"car_radius": 0.05, "comm_radius": 0.5, "default_area_size": 1.0, "dist2goal": 0.01,
              "n_obs": 1, "obs_radius": 0.25, "connect_radius": 0.45}
    KIND = 9

    def __init__(self, num_agents, area_size=None, max_step=128, dt=0.03, params=None):
        params = dict(type(self).PARAMS if params is None else params)
        super().__init__(num_agents, area_size, max_step, dt, params)
        if self._params["n_obs"] != 1:                                   # mpe_connect_spread.py:38-40
            self._params["n_obs"] = 1
            print("WARNING: n_obs is set to 1 for MPEConnectSpread.")

    @property
    def n_cost(self) -> int:
        return 3

    @property
    def cost_components(self) -> Tuple[str, ...]:
        return "agent collisions", "obs collisions", "connectivity"

    def state_lim(self, state=None):
        A = self.area_size
        return torch.tensor([0., 0., -1., -1.]), torch.tensor([A, A * 2, 1., 1.])


class MPECorridor(MPE):
    """dgppo/env/mpe/mpe_corridor.py: MPESpread with two fixed obstacles that leave a corridor of
    `corridor_width` between them, agents sampled below it and goals above it, the y range doubled and
    the agent-obstacle edges always on.  Same kernels, env kind 5."""
    PARAMS = {"car_radius": 0.05, "comm_radius": 0.5, "default_area_size": 1.0, "dist2goal": 0.01,
              "n_obs": 2, "corridor_width": 0.2}
    KIND = 5

    def __init__(self, num_agents, area_size=None, max_step=128, dt=0.03, params=None):
        params = dict(type(self).PARAMS if params is None else params)
        super().__init__(num_agents, area_size, max_step, dt, params)
        if self._params["n_obs"] != 2:                                   # mpe_corridor.py:33-35
            self._params["n_obs"] = 2
            print("WARNING: n_obs is set to 2 for MPECorridor.")
        # solve for the radius of the obstacles (mpe_corridor.py:36-37)
        self._params["obs_radius"] = (self.area_size - self._params["corridor_width"]) / 4

    def state_lim(self, state=None):
        A = self.area_size
        return torch.tensor([0., 0., -1., -1.]), torch.tensor([A, A * 2, 1., 1.])
