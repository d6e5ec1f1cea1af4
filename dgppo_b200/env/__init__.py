"""Environment registry and factory (dgppo/env/__init__.py:9-53).

The environments on BASELINE.json's configs are registered plus
MPETarget and MPECorridor, the first of the "other env families through the same
kernels" row; the remaining MPE / Lidar tasks and VMAS are outside this path.
"""
from typing import Optional

from .base import MultiAgentEnv, StepResult
from .envs import (LidarBicycleTarget, LidarEnv, LidarEnvState, LidarSpread, LidarTarget, MPE,
                   MPECorridor, MPEEnvState, MPESpread, MPETarget, Rectangle)

ENV = {
    "MPETarget": MPETarget,
    "MPECorridor": MPECorridor,
    "MPESpread": MPESpread,
    "LidarSpread": LidarSpread,
    "LidarTarget": LidarTarget,
    "LidarBicycleTarget": LidarBicycleTarget,
}

DEFAULT_MAX_STEP = 128


def make_env(env_id: str, num_agents: int, max_step: int = None, full_observation: bool = False,
             num_obs: Optional[int] = None, n_rays: Optional[int] = None) -> MultiAgentEnv:
    """make_env (dgppo/env/__init__.py:29-53).  The reference overrides the
    class-level PARAMS dict in place (env/__init__.py:38-46); a copy is used
    here so two envs with different `num_obs` can coexist."""
    assert env_id in ENV.keys(), f"Environment {env_id} not implemented."
    params = dict(ENV[env_id].PARAMS)
    max_step = DEFAULT_MAX_STEP if max_step is None else max_step
    if num_obs is not None:
        params["n_obs"] = num_obs
    if n_rays is not None:
        params["n_rays"] = n_rays
    if full_observation:
        area_size = params["default_area_size"]
        params["comm_radius"] = area_size * 10
    return ENV[env_id](num_agents=num_agents, area_size=None, max_step=max_step, dt=0.03, params=params)
