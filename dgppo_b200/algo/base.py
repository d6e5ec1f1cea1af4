"""What the Trainer and user code expect from an algorithm (the reference's dgppo/algo/base.py:10-99,
restated): sizes, `act` / `step` for one decision, `collect` for a batch of rollouts, `update` for one
training iteration, `save` / `load`.  Tensors are torch CUDA tensors; graphs are `GraphsTuple`s, single
or batched over environments."""
from abc import ABC, abstractmethod
from typing import NamedTuple

from ..env.base import MultiAgentEnv
from ..trainer.data import Rollout
from ..utils.graph import GraphsTuple


class _Sizes(NamedTuple):
    node_dim: int
    edge_dim: int
    action_dim: int
    n_agents: int


class Algorithm(ABC):
    def __init__(self, env: MultiAgentEnv, node_dim: int, edge_dim: int, action_dim: int, n_agents: int):
        self._env = env
        self._sizes = _Sizes(node_dim, edge_dim, action_dim, n_agents)
        self.init_rnn_state = None          # set by the subclass: (rnn_layers, n_agents, n_carries, 64)

    # read-only sizes under the reference's names
    node_dim = property(lambda self: self._sizes.node_dim)
    edge_dim = property(lambda self: self._sizes.edge_dim)
    action_dim = property(lambda self: self._sizes.action_dim)
    n_agents = property(lambda self: self._sizes.n_agents)

    # ---- description
    @property
    @abstractmethod
    def config(self) -> dict:
        """Hyper-parameters as written to config.yaml."""

    @property
    @abstractmethod
    def params(self):
        """Parameter pytrees by network name."""

    # ---- one decision
    @abstractmethod
    def act(self, graph: GraphsTuple, rnn_state, params=None):
        """Deterministic action (distribution mode) -> (action, new rnn_state)."""

    @abstractmethod
    def step(self, graph: GraphsTuple, rnn_state, key, params=None):
        """Sampled action -> (action, log_pi, new rnn_state)."""

    # ---- one training iteration
    @abstractmethod
    def collect(self, params, key) -> Rollout:
        """One rollout per key, batched."""

    @abstractmethod
    def update(self, rollout: Rollout, step: int) -> dict:
        """Consume a batch of rollouts; returns logging scalars."""

    # ---- checkpoints
    @abstractmethod
    def save(self, save_dir: str, step: int):
        """Write the parameter pytrees under save_dir/step/."""

    @abstractmethod
    def load(self, load_dir: str, step: int):
        """Read them back."""
