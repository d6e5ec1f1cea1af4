"""dgppo_b200: B200-native (sm_100a) implementation of DGPPO's data-parallel
rollout hot path behind the reference's env / algo API.

    from dgppo_b200.env import make_env
    from dgppo_b200.algo import make_algo

The CUDA kernels live in libdgppo_b200.so (C ABI: include/dgppo_abi.h),
built in-tree by `__graft_entry__.build()` / dgppo_b200/csrc/build.sh.
"""
__version__ = "0.1.0"
