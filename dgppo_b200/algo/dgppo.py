"""DGPPO on the B200 kernels: act / step / collect / det rollout / Vh / GAE /
CBF advantage - the rollout side of the reference's class chain
DGPPO <- InforMARLLagr <- InforMARL <- Algorithm
(dgppo/algo/dgppo.py:25-321, informarl.py:28-472).

Constructor arguments, `config`, `params`, `act`, `step`, `collect`,
`get_Vh`, `update`, `save`, `load` keep the reference's names and meaning.
Parameters are held as reference-shaped pytrees (NumPy leaves) and as packed
fp32 device buffers (algo/params.py).  `update` runs the pre-pass of `DGPPO.update` /
`update_inner` (deterministic rollout, Vl scan, Vh over all (b, T), Dec-OCP
GAE, CBF-residual advantage: dgppo.py:136-273) on the kernels, then the PPO
minibatch scan (dgppo.py:276-321, informarl.py:357-457: update_Vl, update_Vh,
update_policy with gradient clipping and Adam) in `algo/update.py`, whose
gradients come from torch autograd and are mean-all-reduced over the ranks.
"""
from __future__ import annotations

import ctypes as C
import os
import pickle
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from .. import _lib
from ..env.base import MultiAgentEnv, ptr, require_cuda, stream_ptr
from ..env.envs import check_reset
from ..trainer import distributed as D
from ..trainer.data import Rollout
from ..trainer.rollout import RNN_DIM, LazyGraphsTuple, RolloutRecord, run_rollout, run_rollout_chunked
from ..utils.graph import GraphsTuple
from . import params as P
from . import update as U
from .base import Algorithm


def piecewise_constant_schedule(init_value: float, boundaries_and_scales: Dict[int, float]):
    """optax.piecewise_constant_schedule (dgppo.py:74-80)."""
    items = sorted(boundaries_and_scales.items())

    def fn(step: int) -> float:
        v = init_value
        for bnd, sc in items:
            if step >= bnd:
                v *= sc
        return v
    return fn


class DGPPO(Algorithm):

    def __init__(self, env: MultiAgentEnv, node_dim: int, edge_dim: int, state_dim: int, action_dim: int,
                 n_agents: int, actor_gnn_layers: int = 2, Vl_gnn_layers: int = 2, Vh_gnn_layers: int = 1,
                 gamma: float = 0.99, lr_actor: float = 3e-4, lr_Vl: float = 1e-3, lr_Vh: float = 1e-3,
                 batch_size: int = 8192, epoch_ppo: int = 1, clip_eps: float = 0.25, gae_lambda: float = 0.95,
                 coef_ent: float = 1e-2, max_grad_norm: float = 2.0, seed: int = 0, use_rnn: bool = True,
                 rnn_layers: int = 1, rnn_step: int = 16, use_lstm: bool = False, alpha: float = 10.0,
                 cbf_eps: float = 1e-2, cbf_weight: float = 1.0, train_steps: int = 1e5,
                 cbf_schedule: bool = True, **kwargs):
        super().__init__(env=env, node_dim=node_dim, edge_dim=edge_dim, action_dim=action_dim, n_agents=n_agents)
        if not use_rnn or use_lstm or rnn_layers != 1:
            raise NotImplementedError("the B200 kernels cover the default recurrent policy: GRU, 1 layer "
                                      "(--no-rnn / --use-lstm / --rnn-layers>1 are outside this path)")
        self.cost_weight = kwargs.get("cost_weight", 0.)
        self.actor_gnn_layers, self.Vl_gnn_layers, self.Vh_gnn_layers = actor_gnn_layers, Vl_gnn_layers, Vh_gnn_layers
        self.gamma, self.lr_actor, self.lr_Vl, self.lr_Vh = gamma, lr_actor, lr_Vl, lr_Vh
        self.batch_size, self.epoch_ppo, self.clip_eps, self.gae_lambda = batch_size, epoch_ppo, clip_eps, gae_lambda
        self.coef_ent, self.max_grad_norm, self.seed = coef_ent, max_grad_norm, seed
        self.use_rnn, self.rnn_layers, self.rnn_step, self.use_lstm = use_rnn, rnn_layers, rnn_step, use_lstm
        self.cost_schedule = kwargs.get("cost_schedule", False)
        self.state_dim = state_dim
        self.alpha, self.cbf_eps, self.cbf_weight, self.cbf_schedule = alpha, cbf_eps, cbf_weight, cbf_schedule
        if self.cbf_schedule:
            self.cbf_schedule_fn = piecewise_constant_schedule(
                cbf_weight, {int(train_steps * 0.5): 2, int(train_steps * 0.75): 2})

        self.device = require_cuda()
        # rnn carry: (rnn_layers, n_agents, n_carries, 64), zeros (GRUCell.initialize_carry; informarl.py:114-124)
        self.init_rnn_state = torch.zeros((rnn_layers, n_agents, 1, RNN_DIM), device=self.device)
        self.init_Vl_rnn_state = torch.zeros((rnn_layers, 1, 1, RNN_DIM), device=self.device)

        self.policy_cfg = P.net_cfg(_lib.NET_POLICY, node_dim, edge_dim, actor_gnn_layers, action_dim)
        self.Vl_cfg = P.net_cfg(_lib.NET_VL, node_dim, edge_dim, Vl_gnn_layers, 1)
        self.Vh_cfg = P.net_cfg(_lib.NET_VH, node_dim, edge_dim, Vh_gnn_layers, env.n_cost)
        self._trees = {
            "policy": P.init_policy_params(node_dim, edge_dim, action_dim, actor_gnn_layers, seed=seed),
            "Vl": P.init_value_params(node_dim, edge_dim, 1, Vl_gnn_layers, seed=seed + 1),
            "Vh": P.init_value_params(node_dim, edge_dim, env.n_cost, Vh_gnn_layers, seed=seed + 2),
        }
        self._cfgs = {"policy": self.policy_cfg, "Vl": self.Vl_cfg, "Vh": self.Vh_cfg}
        self._packed: Dict[str, tuple] = {}
        self._gen = torch.Generator(device=self.device)
        self._gen.manual_seed(seed)
        # The N(0,1) draw behind the one-sample entropy estimate.  TanhTransformedDistribution.entropy seeds it
        # from Python (distribution.py:40-42: np.random.randint -> jr.PRNGKey), i.e. ONCE, when update_inner is
        # traced by jax.jit: every graph of every minibatch of every update shares one (n_agents, action_dim)
        # sample.  Same here: one draw per algorithm instance (checked against the reference's own closure in
        # tests/test_update_reference.py).
        self.entropy_eps = torch.randn((n_agents, action_dim), generator=self._gen, device=self.device)
        # the deterministic-rollout keys and the minibatch shuffle are per-rank streams (each rank owns its envs)
        self._np_rng = np.random.default_rng([seed, D.world()[0]])
        self._train: Dict[str, dict] = {}          # per net: torch leaves + Adam state (algo/update.py)
        self._lrs = {"policy": lr_actor, "Vl": lr_Vl, "Vh": lr_Vh}
        self.last_prepass: Optional[dict] = None
        self._workspaces: dict = {}
        # independent env groups run on separate streams so env kernels overlap policy kernels
        # compact rollout records: K3's inputs per slot instead of the graph arrays; the value
        # kernels and the update read them directly, GraphsTuple views are built on demand
        self.compact_record = bool(kwargs.get("compact_record", os.environ.get("DGPPO_COMPACT", "1") == "1"))
        # (DGPPO_ROLLOUT_CHUNKS overrides; default by batch size, _n_chunks)
        self.rollout_chunks = int(os.environ.get("DGPPO_ROLLOUT_CHUNKS", "0"))

    # ------------------------------------------------------------ config / params
    @property
    def config(self) -> dict:
        return {
            'cost_weight': self.cost_weight, 'actor_gnn_layers': self.actor_gnn_layers,
            'Vl_gnn_layers': self.Vl_gnn_layers, 'gamma': self.gamma, 'lr_actor': self.lr_actor,
            'lr_Vl': self.lr_Vl, 'batch_size': self.batch_size, 'epoch_ppo': self.epoch_ppo,
            'clip_eps': self.clip_eps, 'gae_lambda': self.gae_lambda, 'coef_ent': self.coef_ent,
            'max_grad_norm': self.max_grad_norm, 'seed': self.seed, 'use_rnn': self.use_rnn,
            'rnn_layers': self.rnn_layers, 'rnn_step': self.rnn_step, 'use_lstm': self.use_lstm,
            'cost_schedule': self.cost_schedule, 'lr_Vh': self.lr_Vh, 'Vh_gnn_layers': self.Vh_gnn_layers,
            'alpha': self.alpha, 'cbf_eps': self.cbf_eps, 'cbf_weight': self.cbf_weight,
            'cbf_schedule': self.cbf_schedule,
        }

    @property
    def params(self) -> dict:
        return {"policy": self._trees["policy"], "Vl": self._trees["Vl"], "Vh": self._trees["Vh"]}

    def set_params(self, name: str, tree: dict) -> None:
        """Replace the pytree of net `name`; its packed device copy is rebuilt on next use and the training
        copy takes the new values (the Adam moments are kept, as TrainState.replace(params=...) keeps them)."""
        self._trees[name] = tree
        self._packed.pop(name, None)
        st = self._train.get(name)
        if st is not None:
            st.load(tree)

    def invalidate(self, name: Optional[str] = None) -> None:
        """Drop the packed device copy of `name` (all nets when None).  Needed only after editing the NumPy
        leaves of a pytree IN PLACE: the cache is keyed on pytree identity and cannot see such edits."""
        if name is None:
            self._packed.clear()
        else:
            self._packed.pop(name, None)

    def packed(self, name: str, params: Optional[dict] = None) -> torch.Tensor:
        """Device buffer of net `name` for the given params pytree.  One cached entry per net, holding a
        reference to the pytree it was packed from and compared with `is` (an `id()` key could be reused by
        a later pytree once the earlier one is freed)."""
        tree = (self.params if params is None else params)[name]
        ent = self._packed.get(name)
        if ent is None or ent[0] is not tree:
            ent = (tree, torch.from_numpy(P.pack_params(tree, self._cfgs[name])).to(self.device))
            self._packed[name] = ent
        return ent[1]

    def _n_chunks(self, b: int) -> int:
        """Env groups (streams) of a rollout.  Measured on B200, C3, captured rollouts: 4096 envs - 4 groups
        best; 512 envs - 1 group 7.4 ms, 2: 8.1, 4: 8.2, 8: 8.6 (kernels too small to share the SMs usefully)."""
        if self.rollout_chunks > 0:
            return self.rollout_chunks
        return 4 if b >= 1024 else 1

    # ------------------------------------------------------------------ helpers
    def _cached(self, name: str, key, make):
        """Workspace reused across update() calls (multi-GB buffers: allocating them per call costs more
        than the kernels that fill them).  One entry per (name, shape key): the train and the evaluation
        shapes each keep theirs, so alternating between them neither reallocates nor lets one overwrite the
        other.  NOTE the returned tensors alias the workspace: a later call with the same shape key
        overwrites them (copy what must outlive the next update / evaluation)."""
        ent = self._workspaces.get((name, key))
        if ent is None:
            if sum(1 for k in self._workspaces if k[0] == name) >= 3:        # bound the number of live shapes
                for k in [k for k in self._workspaces if k[0] == name]:
                    del self._workspaces[k]
            ent = make()
            self._workspaces[(name, key)] = ent
        return ent

    def _shape_key(self, b: int):
        """Everything a record / workspace size depends on: batch, horizon and the env's graph dimensions."""
        d = self._env.graph_dims()
        return (b, self._env.max_episode_steps, self.n_agents, d.n_nodes, d.n_edges, d.node_dim, d.state_dim,
                self._env.n_cost)

    def _eps_from_key(self, key, shape) -> torch.Tensor:
        g = torch.Generator(device=self.device)
        k = np.asarray(key.detach().cpu() if isinstance(key, torch.Tensor) else key).astype(np.uint64).ravel()
        g.manual_seed(int(np.bitwise_xor.reduce(k * np.uint64(0x9E3779B97F4A7C15) + np.uint64(1)) % (2 ** 63)))
        return torch.randn(shape, generator=g, device=self.device, dtype=torch.float32)

    def _policy_call(self, graph: GraphsTuple, rnn_state: torch.Tensor, eps, params):
        n = self.n_agents
        single = graph.is_single
        g = graph.map_arrays(lambda t: t.unsqueeze(0)) if single else graph
        b = g.nodes.shape[0]
        rnn_in = rnn_state.reshape(b, n, RNN_DIM).contiguous().float()
        rnn_out = torch.empty_like(rnn_in)
        action = torch.empty((b, n, self.action_dim), dtype=torch.float32, device=rnn_in.device)
        log_pi = torch.empty((b, n), dtype=torch.float32, device=rnn_in.device) if eps is not None else None
        if eps is not None:
            eps = eps.reshape(b, n, self.action_dim).contiguous()
        cfg = self._env.env_cfg()
        nodes, edges = g.nodes.contiguous(), g.edges.contiguous()
        recv, send = g.receivers.contiguous(), g.senders.contiguous()
        _lib.check(_lib.lib().dgppo_gnn_policy(
            stream_ptr(), C.byref(cfg), C.byref(self.policy_cfg), ptr(self.packed("policy", params)),
            ptr(nodes), ptr(edges), ptr(recv), ptr(send), 1, ptr(rnn_in), ptr(rnn_out), 1,
            ptr(eps), 1, ptr(action), ptr(log_pi), 1, b), "dgppo_gnn_policy")
        new_rnn = rnn_out.reshape(rnn_state.shape)
        if single:
            return action[0], (None if log_pi is None else log_pi[0]), new_rnn
        return action, log_pi, new_rnn

    # ---------------------------------------------------------------- act / step
    def act(self, graph: GraphsTuple, rnn_state: torch.Tensor, params: Optional[dict] = None):
        """InforMARL.act (informarl.py:230-239): deterministic mode tanh(mean)."""
        action, _, rnn_state = self._policy_call(graph, rnn_state, None, params)
        return action, rnn_state

    def step(self, graph: GraphsTuple, rnn_state: torch.Tensor, key, params: Optional[dict] = None):
        """InforMARL.step (informarl.py:241-252): sample tanh(mean + std eps), log_pi.
        `key` seeds the N(0,1) draw (or pass a ready-made tensor of draws)."""
        shape = graph.nodes.shape[:-2] + (self.n_agents, self.action_dim)
        eps = key if (isinstance(key, torch.Tensor) and key.dtype == torch.float32 and tuple(key.shape) == tuple(shape)) \
            else self._eps_from_key(key, shape)
        action, log_pi, rnn_state = self._policy_call(graph, rnn_state, eps, params)
        assert action.shape[-2:] == (self.n_agents, self.action_dim)
        return action, log_pi, rnn_state

    # ------------------------------------------------------------------ rollouts
    def collect(self, params: dict, b_key, eps: Optional[torch.Tensor] = None, graph0: Optional[GraphsTuple] = None,
                record=None, prof=None) -> Rollout:
        """InforMARL.collect (informarl.py:254-256): jit(vmap(rollout)) over the
        env keys == one batched rollout.  `b_key`: one key per environment."""
        fresh = graph0 is None
        if fresh:
            graph0 = self._env.reset(b_key, defer_check=True)      # the feasibility flag is read back below
        b, T = graph0.nodes.shape[0], self._env.max_episode_steps
        if eps is None:
            eps = self._eps_from_key(b_key, (b, T, self.n_agents, self.action_dim))
        if record is None:       # one stochastic record per shape, reused: the returned Rollout aliases it (see _cached)
            record = self._cached("record", self._shape_key(b) + (self.compact_record,),
                                  lambda: RolloutRecord(self._env, b, T, graph0.nodes.device, stochastic=True,
                                                        compact=self.compact_record))
        ro = run_rollout_chunked(self._env, self.policy_cfg, self.packed("policy", params), graph0, eps, T,
                                 self.init_rnn_state, record=record, n_chunks=self._n_chunks(b), prof=prof)
        if fresh:
            check_reset(self._env)                                  # after the rollout is enqueued
        return ro

    def det_rollout_fn(self, params: dict, b_key, graph0: Optional[GraphsTuple] = None, record=None) -> Rollout:
        """DGPPO.det_rollout_fn (dgppo.py:108-117): test_rollout with algo.act."""
        fresh = graph0 is None
        if fresh:
            graph0 = self._env.reset(b_key, defer_check=True)
        if record is None:       # update() calls this every step: keep one deterministic record instead of re-allocating GBs
            record = self._cached("det_record", self._shape_key(graph0.nodes.shape[0]) + (self.compact_record,),
                                  lambda: RolloutRecord(self._env, graph0.nodes.shape[0], self._env.max_episode_steps,
                                                        graph0.nodes.device, stochastic=False,
                                                        compact=self.compact_record))
        ro = run_rollout_chunked(self._env, self.policy_cfg, self.packed("policy", params), graph0, None,
                                 self._env.max_episode_steps, self.init_rnn_state, record=record, test_mode=True,
                                 n_chunks=self._n_chunks(graph0.nodes.shape[0]))
        if fresh:
            check_reset(self._env)
        return ro

    @staticmethod
    def _compact(rollout: Rollout):
        """The compact record behind a Rollout, or None for a full (graph) record."""
        g = rollout.graph
        return g.record if isinstance(g, LazyGraphsTuple) else None

    def _state_record(self, rec) -> "_lib.DgppoStateRecord":
        obs = rec.hits_rec if rec.hits_rec is not None else rec._obstacles
        return _lib.DgppoStateRecord(ptr(rec.agent_rec), ptr(obs), ptr(rec._goal))

    @staticmethod
    def _record_arrays(rollout: Rollout):
        """(b, T+1, ...) graph arrays behind a Rollout: the shared record when the
        Rollout came from run_rollout (graph / next_graph are views of it), else a
        concatenation of graph and next_graph[:, -1]."""
        g, ng = rollout.graph, rollout.next_graph
        T = rollout.rewards.shape[1]
        if isinstance(g, LazyGraphsTuple):          # compact record: build every graph (callers that can, avoid this)
            m = g.record.materialize()
            return [m[k] for k in ("nodes", "edges", "receivers", "senders")]
        out = []
        for name in ("nodes", "edges", "receivers", "senders"):
            a, an = getattr(g, name), getattr(ng, name)
            base = a._base
            if base is not None and base.shape[1] == T + 1 and base.shape[2:] == a.shape[2:] and \
                    a.data_ptr() == base.data_ptr() and base.is_contiguous():
                out.append(base)
            else:
                out.append(torch.cat([a, an[:, -1:]], dim=1).contiguous())
        return out

    # -------------------------------------------------------------------- values
    def _value_record(self, which: str, rollout: Rollout, params) -> torch.Tensor:
        """Vh over all (b, T) graphs + the final one (dgppo.py:219-228): the
        graphs of `rollout.graph` and `rollout.next_graph[:, -1]` are slots
        0..T of one record; the carry fed to the GRU is `rollout.rnn_states`
        for t < T and the policy's post-step carry for the final graph."""
        b, T = rollout.rewards.shape
        n = self.n_agents
        crec = self._compact(rollout)
        dev = rollout.rewards.device
        # carries: t < T as stored; final: act(next_graph[-1], rnn_states[-1]) -> its new carry
        rnn_rec = self._cached("vh_rnn_rec", self._shape_key(b),
                               lambda: torch.empty((b, T + 1, n, RNN_DIM), dtype=torch.float32, device=dev))
        rnn_rec[:, :T] = rollout.rnn_states.reshape(b, T, n, RNN_DIM)
        nc = self._env.n_cost
        Vh = torch.empty((b, T + 1, n, nc), dtype=torch.float32, device=dev)
        # rnn_out: the kernels' scratch rows (the new carry is unused for Vh)
        scratch = self._cached("vh_scratch", self._shape_key(b), lambda: torch.empty_like(rnn_rec))
        cfg = self._env.env_cfg()
        last_in = rollout.rnn_states[:, -1].reshape(b, n, RNN_DIM).contiguous()
        if crec is not None:       # compact record: both forwards straight from the stored states
            st_last = _lib.DgppoStateRecord(
                ptr(crec.agent_rec[:, T]), ptr(crec.hits_rec[:, T] if crec.hits_rec is not None else crec._obstacles),
                ptr(crec._goal))
            act = torch.empty((b, n, self.action_dim), dtype=torch.float32, device=dev)
            _lib.check(_lib.lib().dgppo_gnn_policy_from_state(
                stream_ptr(), C.byref(cfg), C.byref(self.policy_cfg), ptr(self.packed("policy", params)),
                C.byref(st_last), T + 1, ptr(rnn_rec[:, T - 1]), ptr(rnn_rec[:, T]), T + 1, None, 1, ptr(act), None, 1, b),
                "dgppo_gnn_policy_from_state")
            st = self._state_record(crec)
            _lib.check(_lib.lib().dgppo_gnn_value_from_state(
                stream_ptr(), C.byref(cfg), C.byref(self.Vh_cfg), ptr(self.packed("Vh", params)), C.byref(st), T + 1,
                ptr(rnn_rec), ptr(scratch), T + 1, ptr(Vh), T + 1, T + 1, b), "dgppo_gnn_value_from_state")
            return Vh
        nodes, edges, recv, send = self._record_arrays(rollout)
        last = GraphsTuple(*[t[:, -1] if isinstance(t, torch.Tensor) else None for t in rollout.next_graph])
        _, _, final_carry = self._policy_call(last, last_in, None, params)
        rnn_rec[:, T] = final_carry
        _lib.check(_lib.lib().dgppo_gnn_value(
            stream_ptr(), C.byref(cfg), C.byref(self.Vh_cfg), ptr(self.packed("Vh", params)),
            ptr(nodes), ptr(edges), ptr(recv), ptr(send), T + 1, ptr(rnn_rec), ptr(scratch), T + 1,
            ptr(Vh), T + 1, T + 1, b), "dgppo_gnn_value")
        return Vh

    def get_Vh(self, graph: GraphsTuple, rnn_state: torch.Tensor, params: Optional[dict] = None) -> torch.Tensor:
        """DGPPO.get_Vh (dgppo.py:128-134)."""
        n = self.n_agents
        single = graph.is_single
        g = graph.map_arrays(lambda t: t.unsqueeze(0)) if single else graph
        b = g.nodes.shape[0]
        rnn_in = rnn_state.reshape(b, n, RNN_DIM).contiguous().float()
        Vh = torch.empty((b, n, self._env.n_cost), dtype=torch.float32, device=rnn_in.device)
        scratch = torch.empty_like(rnn_in)
        cfg = self._env.env_cfg()
        nodes, edges = g.nodes.contiguous(), g.edges.contiguous()
        recv, send = g.receivers.contiguous(), g.senders.contiguous()
        _lib.check(_lib.lib().dgppo_gnn_value(
            stream_ptr(), C.byref(cfg), C.byref(self.Vh_cfg), ptr(self.packed("Vh", params)),
            ptr(nodes), ptr(edges), ptr(recv), ptr(send), 1, ptr(rnn_in), ptr(scratch), 1, ptr(Vh), 1, 1, b),
            "dgppo_gnn_value")
        return Vh[0] if single else Vh

    def scan_Vl(self, rollout: Rollout, params: Optional[dict] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """InforMARL.scan_Vl + final Vl (informarl.py:281-293, dgppo.py:204-216):
        the centralised value is recurrent over T only through its GRU, so the
        GNN layers of all slots run at once and the head runs T+1 times
        (dgppo_vl_scan).  -> Vl (b, T+1), carries (b, T+1, 64)."""
        b, T = rollout.rewards.shape
        dev = rollout.rewards.device
        carry = torch.zeros((b, T + 2, RNN_DIM), dtype=torch.float32, device=dev)
        carry[:, 0] = self.init_Vl_rnn_state.reshape(RNN_DIM)
        Vl = torch.empty((b, T + 1), dtype=torch.float32, device=dev)
        cfg, pv = self._env.env_cfg(), self.packed("Vl", params)
        crec = self._compact(rollout)
        if crec is not None:
            st = self._state_record(crec)
            _lib.check(_lib.lib().dgppo_vl_scan_from_state(
                stream_ptr(), C.byref(cfg), C.byref(self.Vl_cfg), ptr(pv), C.byref(st), T + 1,
                ptr(carry), T + 2, ptr(Vl), T + 1, T + 1, b), "dgppo_vl_scan_from_state")
            return Vl, carry[:, :T + 1]
        nodes, edges, recv, send = self._record_arrays(rollout)
        # one launch of the GNN layers over all b * (T + 1) graphs (no recurrence there), then the
        # GRU head slot by slot inside the library
        _lib.check(_lib.lib().dgppo_vl_scan(
            stream_ptr(), C.byref(cfg), C.byref(self.Vl_cfg), ptr(pv),
            ptr(nodes), ptr(edges), ptr(recv), ptr(send), T + 1,
            ptr(carry), T + 2, ptr(Vl), T + 1, T + 1, b), "dgppo_vl_scan")
        return Vl, carry[:, :T + 1]

    def gae(self, costs, neg_rewards, Vh, Vl) -> Tuple[torch.Tensor, torch.Tensor]:
        """vmap(compute_dec_ocp_gae) (dgppo.py:232-237; algo/utils.py:11-79)."""
        b, T, n, nh = costs.shape
        Qh = torch.empty_like(costs)
        Ql = torch.empty((b, T), dtype=torch.float32, device=costs.device)
        costs, neg_rewards, Vh, Vl = costs.contiguous(), neg_rewards.contiguous(), Vh.contiguous(), Vl.contiguous()
        _lib.check(_lib.lib().dgppo_gae(stream_ptr(), ptr(costs), ptr(neg_rewards), ptr(Vh), ptr(Vl),
                                         self.gamma, self.gae_lambda, ptr(Qh), ptr(Ql), b, T, n, nh), "dgppo_gae")
        return Qh, Ql

    def cbf_advantage(self, Ql, Vl, Vh, step: int):
        """Advantage merge (dgppo.py:239-259) -> A (b,T,n), cbf_deriv, Acbf, is_safe."""
        b, Tp1, n, nh = Vh.shape
        T = Tp1 - 1
        dev = Vh.device
        A = torch.empty((b, T, n), dtype=torch.float32, device=dev)
        deriv = torch.empty((b, T, n, nh), dtype=torch.float32, device=dev)
        acbf = torch.empty_like(deriv)
        safe = torch.empty((b, T, n), dtype=torch.uint8, device=dev)
        w = self.cbf_schedule_fn(step) if self.cbf_schedule else self.cbf_weight
        Ql, Vl, Vh = Ql.contiguous(), Vl.contiguous(), Vh.contiguous()
        _lib.check(_lib.lib().dgppo_cbf_advantage(
            stream_ptr(), ptr(Ql), ptr(Vl), ptr(Vh),
            float(self._env.dt), float(self.alpha), float(self.cbf_eps), float(w),
            ptr(A), ptr(deriv), ptr(acbf), ptr(safe), b, T, n, nh), "dgppo_cbf_advantage")
        return A, deriv, acbf, safe.bool()

    # -------------------------------------------------------------------- update
    def _train_state(self, name: str) -> "U.NetTrainState":
        st = self._train.get(name)
        if st is None:
            st = U.NetTrainState(self._trees[name], self.device, self._lrs[name])
            self._train[name] = st
        return st

    def _sync_params_from_training(self) -> None:
        """Training parameters -> the NumPy pytrees `params` exposes (and the kernels' packed copies)."""
        for name, st in self._train.items():
            self._trees[name] = st.numpy_tree()
            self._packed.pop(name, None)

    def prepass(self, rollout: Rollout, step: int) -> dict:
        """Everything `update_inner` computes before its minibatch scan (dgppo.py:204-273), on the kernels:
        deterministic rollout, Vl scan, Vh on both records, both GAE passes, the CBF advantage merge."""
        b = rollout.dones.shape[0]
        key = self._np_rng.integers(0, 2 ** 31 - 1, size=b)
        det_rollout = self.det_rollout_fn(self.params, key)
        Vl, Vl_carries = self.scan_Vl(rollout)
        Vh = self._value_record("Vh", rollout, None)
        Qh, Ql = self.gae(rollout.costs, -rollout.rewards, Vh, Vl)
        A, deriv, acbf, is_safe = self.cbf_advantage(Ql, Vl, Vh, step)
        Vh_det = self._value_record("Vh", det_rollout, None)
        Qh_det, _ = self.gae(det_rollout.costs, -det_rollout.rewards, Vh_det, Vl)
        T = rollout.dones.shape[1]
        self.last_prepass = dict(det_rollout=det_rollout, bTp1_Vl=Vl, bT_Vl_rnn_states=Vl_carries[:, :T],
                                 bTp1ah_Vh=Vh, bTah_Qh=Qh, bT_Ql=Ql, bTa_A=A, bTah_cbf_deriv=deriv,
                                 bTah_Acbf=acbf, bTa_is_safe=is_safe, bTp1ah_Vh_det=Vh_det,
                                 bTah_Qh_det=Qh_det)
        return self.last_prepass

    def update(self, rollout: Rollout, step: int) -> dict:
        """DGPPO.update (dgppo.py:136-170): pre-pass, then `epoch_ppo` passes of the minibatch scan
        `update_fn` (dgppo.py:276-289): update_Vl -> update_Vh -> update_policy per minibatch."""
        b, T = rollout.dones.shape
        rank, world = D.world()
        mb_envs = (self.batch_size // world) // T        # this rank's share of every minibatch
        if mb_envs < 1 or b * T * world < self.batch_size:
            raise ValueError(f"batch_size {self.batch_size} needs at least {self.batch_size // T} environments "
                             f"in total and T * world_size = {T * world} steps per minibatch; got {b} envs "
                             f"per rank on {world} rank(s) (dgppo.py:153,159)")
        if T % self.rnn_step:
            raise ValueError(f"rnn_step {self.rnn_step} must divide the horizon {T} (dgppo.py:157-158)")
        pp = self.prepass(rollout, step)
        det = pp["det_rollout"]
        gi = U.GraphIndex(self.n_agents, self._env.graph_dims().n_ag, self._env.graph_dims().n_ao,
                          self._env.graph_dims().n_nodes, self.device)
        crec, cdet = self._compact(rollout), self._compact(det)
        arrays = None if crec is not None else self._record_arrays(rollout)
        det_arrays = None if cdet is not None else self._record_arrays(det)

        def mb_graphs(rec, arrs, ix):
            """Graphs of one minibatch, (mb * T, ...): slices of the full record, or (compact record) built by
            K3 for just these environments."""
            if rec is not None:
                m = rec.graphs_of(ix, rec._goal, rec._obstacles)
                arrs, ix = [m[k] for k in ("nodes", "edges", "receivers", "senders")], slice(None)
            return U.chunk_graphs(arrs, ix, T, gi, torch.float32)
        n = self.n_agents
        rnn_det = det.rnn_states.reshape(b, T, n, RNN_DIM)
        st_pi, st_Vl, st_Vh = self._train_state("policy"), self._train_state("Vl"), self._train_state("Vh")

        def minibatch_step(nodes, edge_feat, sidx, mask, d_nodes, d_edge_feat, d_sidx, d_mask, Ql, rnn_d, Qh_d, act, lp_old,
                           adv, eps):
            """update_fn body (dgppo.py:278-287): Vl, Vh, policy - each: loss, gradient, all-reduce, clip, Adam."""
            g = {"nodes": nodes, "edge_feat": edge_feat, "sidx": sidx, "mask": mask}
            gd = {"nodes": d_nodes, "edge_feat": d_edge_feat, "sidx": d_sidx, "mask": d_mask}
            out = {}
            loss = U.loss_Vl(st_Vl.tree(), g, Ql, gi, self.Vl_gnn_layers, self.rnn_step)
            r = st_Vl.step(loss, self.max_grad_norm)
            out.update({"Vl/loss": loss.detach(), "Vl/grad_norm": r["grad_norm"], "Vl/has_nan": r["has_nan"]})
            loss = U.loss_Vh(st_Vh.tree(), gd, rnn_d, Qh_d, gi, self.Vh_gnn_layers)
            r = st_Vh.step(loss, self.max_grad_norm)
            out.update({"Vh/loss_Vh": loss.detach(), "Vh/grad_Vh_norm": r["grad_norm"], "Vh/grad_Vh_has_nan": r["has_nan"]})
            loss, pinfo = U.loss_policy(st_pi.tree(), g, act, lp_old, adv, eps, gi, self.actor_gnn_layers, self.rnn_step,
                                        self.clip_eps, self.coef_ent)
            r = st_pi.step(loss, self.max_grad_norm)
            out.update({"policy/loss": loss.detach(), "policy/grad_norm": r["grad_norm"], "policy/has_nan": r["has_nan"],
                        **{k: v.detach() for k, v in pinfo.items()}})
            return out

        # CUDA-graph replay of the minibatch step: single rank only by default.  With several ranks the step contains
        # the NCCL all-reduce; capturing it works, but tearing the process group down with such graphs alive hung
        # (8 ranks, torch 2.11 / NCCL 2.28), so multi-rank runs take the eager step unless DGPPO_UPDATE_GRAPH=2.
        ug = os.environ.get("DGPPO_UPDATE_GRAPH", "1")
        use_graph = ug == "2" or (ug == "1" and world == 1)
        # DGPPO_UPDATE_TF32=1: let the update's GEMMs run as TF32 (XLA's default precision for f32 dots on NVIDIA
        # GPUs, i.e. what the reference itself trains with); off by default - the update then is fp32 throughout
        tf32_prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = os.environ.get("DGPPO_UPDATE_TF32", "0") == "1"
        info = {}
        try:
            for _ in range(self.epoch_ppo):
                idx = np.arange(b)
                self._np_rng.shuffle(idx)
                for mb in np.array_split(idx, max(1, b // mb_envs)):
                    ix = torch.as_tensor(mb, dtype=torch.long, device=self.device)
                    g = mb_graphs(crec, arrays, ix)
                    gd = mb_graphs(cdet, det_arrays, ix)
                    eps = self.entropy_eps.expand(len(mb), T, n, self.action_dim)
                    inputs = (g["nodes"], g["edge_feat"], g["sidx"], g["mask"], gd["nodes"], gd["edge_feat"],
                              gd["sidx"], gd["mask"], pp["bT_Ql"][ix], rnn_det[ix], pp["bTah_Qh_det"][ix],
                              rollout.actions[ix], rollout.log_pis[ix], pp["bTa_A"][ix], eps)
                    if use_graph:
                        # the whole minibatch step replayed as one CUDA graph (captured once per shape and mode)
                        key = ("update_graph", torch.backends.cuda.matmul.allow_tf32, os.environ.get("DGPPO_UPDATE_GNN"),
                               os.environ.get("DGPPO_UPDATE_SPLITK"), tuple((tuple(x.shape), x.dtype) for x in inputs))
                        gs = self._workspaces.get(key)
                        if gs is None:
                            states = st_Vl.state_tensors() + st_Vh.state_tensors() + st_pi.state_tensors()
                            gs = U.GraphedStep(minibatch_step, inputs, states)
                            self._workspaces[key] = gs
                        info = dict(gs(*inputs))
                    else:
                        info = minibatch_step(*inputs)
        finally:
            torch.backends.cuda.matmul.allow_tf32 = tf32_prev
        self._sync_params_from_training()
        info["policy/log_pi_min"] = rollout.log_pis.min()
        info["Vl/max_target"], info["Vl/min_target"] = pp["bT_Ql"].max(), pp["bT_Ql"].min()
        info["eval/safe_data"] = pp["bTa_is_safe"].float().mean()
        return {k: float(v.detach()) if isinstance(v, torch.Tensor) else float(v) for k, v in info.items()}

    # ---------------------------------------------------------------- save / load
    def save(self, save_dir: str, step: int):
        """Pickle the parameter pytrees (informarl_lagr.py:311-317)."""
        model_dir = os.path.join(save_dir, str(step))
        os.makedirs(model_dir, exist_ok=True)
        for name, fn in (("policy", "actor.pkl"), ("Vl", "Vl.pkl"), ("Vh", "Vh.pkl")):
            with open(os.path.join(model_dir, fn), "wb") as f:
                pickle.dump(self._trees[name], f)

    def load(self, load_dir: str, step: int):
        """Load pickled parameter pytrees (informarl_lagr.py:319-327)."""
        path = os.path.join(load_dir, str(step))
        for name, fn in (("policy", "actor.pkl"), ("Vl", "Vl.pkl"), ("Vh", "Vh.pkl")):
            with open(os.path.join(path, fn), "rb") as f:
                tree = pickle.load(f)
            self.set_params(name, _to_numpy_tree(tree))


def _to_numpy_tree(t):
    """Any mapping-shaped pytree (dict, flax FrozenDict, ...) with array-like leaves (NumPy, jax) -> nested
    dicts of fp32 NumPy arrays."""
    from collections.abc import Mapping
    if isinstance(t, Mapping):
        return {k: _to_numpy_tree(v) for k, v in t.items()}
    return np.asarray(t, np.float32)
