"""gym_lmaze_b200 -- B200-native batched LMaze step/reset (drop-in for the hot path of gkm2708/gym-lmaze).

Keeps the reference's registration ids (reference gym_lmaze/__init__.py:3-38) for the
variants that are built and adds vectorised ids.  `make(id, num_envs=...)` works without
gym; when gym or gymnasium is importable the ids are registered there as well.
"""
from .envs import LmazeVecCuda, shard_range, allreduce_stats, INVALID_ACTION  # noqa: F401
from .envs import LmazeEnv, LmazeEnv_v2, LmazeEnv_v3, LmazeEnv_v4, LmazeEnv_v5, LmazeEnv_v6  # noqa: F401
from .envs import LmazeHierCuda  # noqa: F401

__version__ = "0.1.0"

# id -> (variant, default kwargs).  'lmaze-vK' are the reference's ids; 'lmaze-vec-vK' the batched ones.
_REGISTRY = {}


def register(id, variant, **defaults):
    _REGISTRY[id] = (variant, defaults)


_SINGLE = {"v0": LmazeEnv, "v2": LmazeEnv_v2, "v3": LmazeEnv_v3, "v4": LmazeEnv_v4, "v5": LmazeEnv_v5, "v6": LmazeEnv_v6}


def make(id, **kwargs):
    """'lmaze-vK' -> the reference's single-maze surface (numpy obs, float reward, bool done, action);
    'lmaze-vec-vK' (or num_envs=...) -> the batched LmazeVecCuda."""
    if id not in _REGISTRY:
        raise KeyError("unknown env id %r; built ids: %s" % (id, ", ".join(sorted(_REGISTRY))))
    variant, defaults = _REGISTRY[id]
    kw = dict(defaults)
    kw.update(kwargs)
    if "-vec-" not in id and "num_envs" not in kwargs:
        kw.pop("num_envs", None)
        return _SINGLE[variant](**kw)
    if variant in ("v5", "v6"):
        return LmazeHierCuda(variant=variant, **kw)
    return LmazeVecCuda(variant=variant, **kw)


def registered_ids():
    return sorted(_REGISTRY)


register("lmaze-v0", "v0", num_envs=1)
register("lmaze-v2", "v2", num_envs=1)
register("lmaze-v3", "v3", num_envs=1)
register("lmaze-v4", "v4", num_envs=1)
register("lmaze-v5", "v5", num_envs=1)
register("lmaze-v6", "v6", num_envs=1)
register("lmaze-vec-v5", "v5", num_envs=4096)
register("lmaze-vec-v6", "v6", num_envs=4096)
register("lmaze-vec-v0", "v0", num_envs=4096)
register("lmaze-vec-v2", "v2", num_envs=4096)
register("lmaze-vec-v3", "v3", num_envs=4096)
register("lmaze-vec-v4", "v4", num_envs=4096)


def _register_with_gym():
    for modname in ("gymnasium", "gym"):
        try:
            mod = __import__(modname + ".envs.registration", fromlist=["register"])
            if not hasattr(__import__(modname), "__version__"):
                continue
            for env_id, (variant, defaults) in _REGISTRY.items():
                try:
                    mod.register(id=env_id, entry_point="gym_lmaze_b200:LmazeHierCuda" if variant in ("v5", "v6")
                                 else "gym_lmaze_b200:LmazeVecCuda",
                                 kwargs=dict(defaults, variant=variant))
                except Exception:
                    pass
        except Exception:
            continue


_register_with_gym()
