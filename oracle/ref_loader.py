"""Load the UNMODIFIED reference env modules by file path (test infrastructure).

Only usable where the reference checkout exists (this build container:
/root/reference).  It is used to (a) generate the golden fixtures committed under
tests/golden/ and (b) cross-check the C restatement (oracle/lmaze_oracle.c)
live in the CPU test tier.  Nothing on the product path imports this file, and
nothing that runs on the GPU box may depend on it.

The reference imports `gym` (absent from this image) only for the Env base class
and two space constructors (reference gym_lmaze/envs/lmaze_env.py:1-3,11,16,20),
so tests/_gymstub provides a stand-in.  Modules are loaded one by one with
importlib because `import gym_lmaze.envs` itself fails (envs/__init__.py:8
imports a non-existent lmaze_env_v7, and v1 imports matplotlib).
"""
import contextlib
import importlib.util
import io
import os
import sys

REFERENCE_ROOT = os.environ.get("LMAZE_REFERENCE_ROOT", "/root/reference")
_STUB_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "_gymstub")

_FILES = {
    "v0": ("gym_lmaze/envs/lmaze_env.py", "LmazeEnv"),
    "v2": ("gym_lmaze/envs/lmaze_env_v2.py", "LmazeEnv_v2"),
    "v3": ("gym_lmaze/envs/lmaze_env_v3.py", "LmazeEnv_v3"),
    "v4": ("gym_lmaze/envs/lmaze_env_v4.py", "LmazeEnv_v4"),
    "v5": ("gym_lmaze/envs/lmaze_env_v5.py", "LmazeEnv_v5"),
    "v6": ("gym_lmaze/envs/lmaze_env_v6.py", "LmazeEnv_v6"),
}


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, _FILES["v0"][0]))


def load_reference_module(variant="v0"):
    """Exec the reference file for `variant` and return the module object."""
    rel, _ = _FILES[variant]
    path = os.path.join(REFERENCE_ROOT, rel)
    if not os.path.isfile(path):
        raise FileNotFoundError(path)
    stub = os.path.abspath(_STUB_DIR)
    if "gym" not in sys.modules:
        sys.path.insert(0, stub)
        try:
            import gym  # noqa: F401  (the stub, or a real gym if one is ever installed)
        finally:
            sys.path.remove(stub)
    name = "_lmaze_reference_" + variant
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class ScriptedRandom(object):
    """Stands in for the stdlib `random` module inside a loaded reference module.

    The reference draws spawn / goal coordinates with `random.randint(1, G-2)` in a
    rejection loop (lmaze_env.py:73-75, lmaze_env_v3.py:148-151,157-160).  Parity
    runs must start both implementations from the same cell, so the test queues
    the coordinates it wants; an accepted pair ends the loop after one iteration.
    """

    def __init__(self):
        self.queue = []
        self.calls = 0

    def push(self, *values):
        self.queue.extend(int(v) for v in values)

    def randint(self, a, b):
        self.calls += 1
        if not self.queue:
            raise RuntimeError("ScriptedRandom exhausted")
        v = self.queue.pop(0)
        assert a <= v <= b, (a, v, b)
        return v


def make_reference_env(variant="v0", first_draws=(1, 1, 1, 2)):
    """Instantiate the reference env with its `random` replaced by ScriptedRandom.

    Returns (env, scripted, module).  The constructor itself calls reset() once
    (lmaze_env.py:52, lmaze_env_v3.py:132), which consumes `first_draws`
    (v0: one x,y pair; v3: goal pair then ball pair).
    """
    mod = load_reference_module(variant)
    scripted = ScriptedRandom()
    mod.random = scripted
    scripted.push(*first_draws)
    cls = getattr(mod, _FILES[variant][1])
    with contextlib.redirect_stdout(io.StringIO()):  # v0 prints "init-init"/"init-end"
        env = cls()
    scripted.queue.clear()
    return env, scripted, mod
