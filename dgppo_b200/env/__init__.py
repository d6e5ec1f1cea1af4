"""Environment registry and factory (dgppo/env/__init__.py:9-53).

Every MPE and LidarEnv task of the reference's registry runs through the same kernels (SURVEY.md 8 and
8f.4): the environments on BASELINE.json's configs plus MPETarget, MPECorridor, the landmark families
(LidarLine, MPELine, MPEFormation) and MPEConnectSpread with its third cost.  VMAS is outside this path.
"""
from typing import Optional

from .base import MultiAgentEnv, StepResult
from .envs import (LidarBicycleTarget, LidarEnv, LidarEnvState, LidarLine, LidarSpread, LidarTarget, MPE,
                   MPEConnectSpread, MPECorridor, MPEEnvState, MPEFormation, MPELine, MPESpread, MPETarget,
                   Rectangle)

# name -> class; the names are the reference's `--env` values (env/__init__.py:9-23: every MPE and LidarEnv task)
ENV = {cls.__name__: cls for cls in (MPETarget, MPESpread, MPELine, MPEFormation, MPECorridor, MPEConnectSpread,
                                     LidarSpread, LidarTarget, LidarLine, LidarBicycleTarget)}

DEFAULT_MAX_STEP = 128


def make_env(env_id: str, num_agents: int, max_step: int = None, full_observation: bool = False,
             num_obs: Optional[int] = None, n_rays: Optional[int] = None) -> MultiAgentEnv:
    """Build an environment the way the reference's factory does (dgppo/env/__init__.py:29-53):
    `num_obs` / `n_rays` override the class defaults, `full_observation` widens the communication radius to
    ten times the arena.  The overrides go into a COPY of the class-level PARAMS (the reference edits the
    class dict in place, env/__init__.py:38-46), so two differently configured envs can coexist."""
    if env_id not in ENV:
        raise AssertionError(f"Environment {env_id} not implemented.")
    cls = ENV[env_id]
    overrides = {"n_obs": num_obs, "n_rays": n_rays}
    params = {**cls.PARAMS, **{k: v for k, v in overrides.items() if v is not None}}
    if full_observation:
        params["comm_radius"] = params["default_area_size"] * 10
    return cls(num_agents=num_agents, area_size=None, dt=0.03, params=params,
               max_step=DEFAULT_MAX_STEP if max_step is None else max_step)
