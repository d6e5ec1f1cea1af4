"""Trainer loop (dgppo/trainer/trainer.py:20-141) with the rollout sharded
across ranks.  Evaluation printing / wandb logging are host glue outside the
hot path; what is kept is the call structure: key split -> algo.collect ->
algo.update (trainer.py:131-139), and the evaluation rollout
(test_rollout with algo.act, trainer.py:85-125)."""
from __future__ import annotations

import os
from time import time

import numpy as np
import torch

from ..algo.base import Algorithm
from ..env.base import MultiAgentEnv
from . import distributed as D


class Trainer:

    def __init__(self, env: MultiAgentEnv, env_test: MultiAgentEnv, algo: Algorithm, n_env_train: int,
                 n_env_test: int, log_dir: str, seed: int, params: dict, save_log: bool = True):
        self.env, self.env_test, self.algo = env, env_test, algo
        self.n_env_train, self.n_env_test = n_env_train, n_env_test
        self.log_dir, self.seed = log_dir, seed
        assert set(params) >= {"run_name", "training_steps", "eval_interval", "eval_epi", "save_interval"}
        self.params = params
        self.steps = params["training_steps"]
        self.eval_interval, self.save_interval = params["eval_interval"], params["save_interval"]
        self.save_log = save_log and D.world()[0] == 0
        self.model_dir = os.path.join(log_dir, "models")
        if self.save_log:
            os.makedirs(self.model_dir, exist_ok=True)
        self.rng = np.random.default_rng(seed)       # the same stream on every rank: keys are drawn globally, then sharded
        self.update_steps = 0
        rank, world = D.world()
        if n_env_train % world:
            raise ValueError(f"n_env_train = {n_env_train} must be a multiple of the number of ranks ({world}): the "
                             "gradient mean over ranks equals the single-device mean only for equal shards")
        T = env.max_episode_steps
        bs = getattr(algo, "batch_size", None)
        if bs is not None and (bs % (world * T) or (n_env_train // world) * T * world < bs):
            raise ValueError(f"batch_size = {bs} must be a multiple of world_size * T = {world * T} and at most "
                             f"n_env_train * T = {n_env_train * T} (dgppo.py:153-159)")

    def evaluate(self, step: int, start_time: float) -> dict:
        """trainer.py:105-125: deterministic rollouts on n_env_test fixed keys."""
        keys = np.arange(1000)[:self.n_env_test] + 1_000_000 * self.seed
        ro = self.algo.det_rollout_fn(self.algo.params, keys) if self.env_test is self.env else \
            self._det_rollout_on(self.env_test, keys)
        total_reward = ro.rewards.sum(dim=-1)
        cost = torch.clamp(ro.costs, min=0.0).amax(dim=-1).amax(dim=-1).sum(dim=-1).mean()
        unsafe_frac = (ro.costs.amax(dim=-1).amax(dim=-2) >= 1e-6).float().mean()
        info = {"eval/reward": float(total_reward.mean()), "eval/reward_final": float(ro.rewards[:, -1].mean()),
                "eval/cost": float(cost), "eval/unsafe_frac": float(unsafe_frac)}
        if D.world()[0] == 0:
            print(f"step: {step:3}, time: {time() - start_time:5.0f}s, reward: {info['eval/reward']:9.4f}, "
                  f"min/max reward: {float(total_reward.min()):7.2f}/{float(total_reward.max()):7.2f}, "
                  f"cost: {info['eval/cost']:8.4f}, unsafe_frac: {info['eval/unsafe_frac']:6.2f}")
        return info

    def _det_rollout_on(self, env, keys):
        saved = self.algo._env
        try:
            self.algo._env = env
            return self.algo.det_rollout_fn(self.algo.params, keys)
        finally:
            self.algo._env = saved

    def train(self):
        start_time = time()
        for step in range(0, self.steps + 1):
            if step % self.eval_interval == 0:
                self.evaluate(step, start_time)
            if self.save_log and step % self.save_interval == 0:
                self.algo.save(self.model_dir, step)
            # collect rollouts: one key per environment, sharded over the ranks (trainer.py:131-134)
            keys = self.rng.integers(0, 2 ** 31 - 1, size=self.n_env_train)
            rollouts = self.algo.collect(self.algo.params, D.shard_keys(keys))
            # update the algorithm (trainer.py:137)
            update_info = self.algo.update(rollouts, step)
            self.update_steps += 1
            if D.world()[0] == 0 and step % self.eval_interval == 0 and update_info:
                keys_ = ("policy/loss", "Vl/loss", "Vh/loss_Vh", "policy/entropy", "eval/safe_data")
                print("        " + ", ".join(f"{k}: {update_info[k]:.4g}" for k in keys_ if k in update_info))
