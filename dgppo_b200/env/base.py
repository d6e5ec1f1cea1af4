"""MultiAgentEnv: the environment interface of the reference
(dgppo/env/base.py:30-150), natively batched.

Differences from the reference that follow from the B200 design:
  * every method accepts and returns tensors with optional leading batch axes
    (the reference is single-env code batched by jax.vmap; here the CUDA
    kernels are batched over environments themselves);
  * arrays are torch CUDA tensors (fp32 / int32); the arithmetic runs in
    libdgppo_b200.so - there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from abc import ABC, abstractmethod
from typing import NamedTuple, Optional, Tuple

import torch

from .. import _lib
from ..utils.graph import GraphsTuple


class StepResult(NamedTuple):
    graph: GraphsTuple
    reward: torch.Tensor
    cost: torch.Tensor
    done: torch.Tensor
    info: dict


def require_cuda() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("dgppo_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def dev_f32(x, device) -> torch.Tensor:
    return torch.as_tensor(x, dtype=torch.float32, device=device).contiguous()


def ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


class MultiAgentEnv(ABC):

    PARAMS = {}
    KIND = -1

    def __init__(self, num_agents: int, area_size: float, max_step: int = 256, dt: float = 0.03,
                 params: Optional[dict] = None):
        super().__init__()
        self._num_agents = num_agents
        self._dt = dt
        if params is None:
            params = self.PARAMS
        self._params = params
        self._t = 0
        self._max_step = max_step
        self._area_size = area_size

    # ---- properties mirrored from env/base.py:50-105
    @property
    def params(self) -> dict:
        return self._params

    @property
    def num_agents(self) -> int:
        return self._num_agents

    @property
    def area_size(self) -> float:
        return self._area_size

    @property
    def dt(self) -> float:
        return self._dt

    @property
    def max_episode_steps(self) -> int:
        return self._max_step

    @property
    @abstractmethod
    def n_cost(self) -> int: ...

    @property
    @abstractmethod
    def cost_components(self) -> Tuple[str, ...]: ...

    @property
    @abstractmethod
    def state_dim(self) -> int: ...

    @property
    @abstractmethod
    def node_dim(self) -> int: ...

    @property
    @abstractmethod
    def edge_dim(self) -> int: ...

    @property
    @abstractmethod
    def action_dim(self) -> int: ...

    def clip_state(self, state: torch.Tensor) -> torch.Tensor:
        lo, hi = self.state_lim(state)
        return torch.minimum(torch.maximum(state, lo.to(state.device)), hi.to(state.device))

    def clip_action(self, action: torch.Tensor) -> torch.Tensor:
        lo, hi = self.action_lim()
        return torch.minimum(torch.maximum(action, lo.to(action.device)), hi.to(action.device))

    # ---- kernel-side description
    def env_cfg(self) -> _lib.DgppoEnvCfg:
        p = self._params
        return _lib.DgppoEnvCfg(
            self.KIND, self._num_agents, int(p.get("n_obs", 0)), int(p.get("n_rays", 32)),
            int(p.get("top_k_rays", 8)), 0, float(p["comm_radius"]), float(p["car_radius"]),
            float(p.get("obs_radius", 0.05)), float(self._area_size), float(self._dt),
            float(p["dist2goal"]), float(p.get("connect_radius", 0.0)), self._goal_table_ptr())

    def _goal_table_ptr(self):
        """Device table the kernels need beside the scalars (MPEFormation: goal offsets); None otherwise."""
        return None

    def graph_dims(self) -> _lib.DgppoGraphDims:
        d = _lib.DgppoGraphDims()
        cfg = self.env_cfg()
        _lib.check(_lib.lib().dgppo_graph_dims(C.byref(cfg), C.byref(d)), "dgppo_graph_dims")
        return d

    @abstractmethod
    def reset(self, key) -> GraphsTuple: ...

    @abstractmethod
    def step(self, graph: GraphsTuple, action: torch.Tensor, get_eval_info: bool = False) -> StepResult: ...

    @abstractmethod
    def get_cost(self, graph: GraphsTuple) -> torch.Tensor: ...

    @abstractmethod
    def state_lim(self, state: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]: ...

    @abstractmethod
    def action_lim(self) -> Tuple[torch.Tensor, torch.Tensor]: ...

    @abstractmethod
    def get_graph(self, state, lidar_data=None) -> GraphsTuple: ...

    def render_video(self, *args, **kwargs) -> None:
        raise NotImplementedError("rendering is outside the rollout hot path (SURVEY.md 2, row 11)")
