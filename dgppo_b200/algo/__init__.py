"""Algorithm factory (dgppo/algo/__init__.py:8-18).  Only DGPPO - the
north-star path - is provided; the baselines (informarl, informarl_lagr,
hcbfcrpo) share the same rollout kernels and are outside this path."""
from .base import Algorithm
from .dgppo import DGPPO


def make_algo(algo: str, **kwargs) -> Algorithm:
    if algo == 'dgppo':
        return DGPPO(**kwargs)
    raise ValueError(f'Unknown algorithm: {algo}')
