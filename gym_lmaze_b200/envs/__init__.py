"""Env classes.  The reference's envs/__init__.py imports a module that does not
exist (lmaze_env_v7, reference gym_lmaze/envs/__init__.py:8); only what is built is exported here."""
from .lmaze_vec_cuda import LmazeVecCuda, shard_range, allreduce_stats, INVALID_ACTION  # noqa: F401
from .lmaze_hier_cuda import LmazeHierCuda  # noqa: F401,E402
from .lmaze_compat import LmazeEnv, LmazeEnv_v2, LmazeEnv_v3, LmazeEnv_v4, LmazeEnv_v5, LmazeEnv_v6  # noqa: F401,E402
