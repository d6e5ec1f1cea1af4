"""The rollout record handed from `algo.collect` to `algo.update`.

Field names and order are the reference's (dgppo/trainer/data.py:8-32) because user code unpacks and
indexes them; everything else here is this repo's: torch tensors, natively batched over environments.

    field        shape                          notes
    graph        GraphsTuple of (b, T, ...)     view [:, :T] of the (b, T+1, ...) record (trainer/rollout.py)
    actions      (b, T, n, 2)                   as sampled (unclipped)
    rnn_states   (b, T, 1, n, 1, 64)            policy GRU carry BEFORE step t (`rollout`) / AFTER it (`test_rollout`)
    rewards      (b, T)
    costs        (b, T, n, n_cost)
    dones        (b, T) bool                    always False (lidar_env/base.py:172)
    log_pis      (b, T, n) | None               None for the deterministic rollout
    next_graph   GraphsTuple of (b, T, ...)     view [:, 1:] of the same record: next_graph[t] is graph[t+1]
"""
from typing import NamedTuple, Optional, Tuple

import torch

from ..utils.graph import GraphsTuple


class Rollout(NamedTuple):
    graph: GraphsTuple
    actions: torch.Tensor
    rnn_states: torch.Tensor
    rewards: torch.Tensor
    costs: torch.Tensor
    dones: torch.Tensor
    log_pis: Optional[torch.Tensor]
    next_graph: GraphsTuple

    def _sizes(self) -> Tuple[int, ...]:
        """Leading axes of the reward array: (envs, T) for a batch, (T,) for a single trajectory."""
        return tuple(self.rewards.shape)

    # the reference's size helpers (data.py:18-32) read axes 0 / 1 / 2 of `rewards`
    length = property(lambda self: self._sizes()[0])
    time_horizon = property(lambda self: self._sizes()[1])
    num_agents = property(lambda self: self._sizes()[2])
    n_data = property(lambda self: self._sizes()[0] * self._sizes()[1])

    def nbytes(self) -> int:
        """Bytes of the per-step arrays (graphs excluded: they are views of the shared record)."""
        ts = [self.actions, self.rnn_states, self.rewards, self.costs, self.dones, self.log_pis]
        return sum(t.numel() * t.element_size() for t in ts if t is not None)
