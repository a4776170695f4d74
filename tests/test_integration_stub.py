"""The ctypes stub printed in INTEGRATION.md must stay a working binding of the C ABI."""
import ctypes
import os
import re

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def load_stub():
    from gym_lmaze_b200 import build, _abi
    build.build()
    md = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    code = re.search(r"```python\n(# gym_lmaze/envs/lmaze_env_cuda\.py.*?)```", md, re.S).group(1)
    code = code.replace('ctypes.CDLL("liblmaze_b200.so")', "ctypes.CDLL(%r)" % _abi.LIB_PATH)
    ns = {}
    exec(compile(code, "INTEGRATION.md", "exec"), ns)
    return ns, _abi


def test_stub_struct_matches_header():
    ns, _abi = load_stub()
    assert ctypes.sizeof(ns["_Cfg"]) == ctypes.sizeof(_abi.LmzConfig)
    assert [f[0] for f in ns["_Cfg"]._fields_] == [f[0] for f in _abi.LmzConfig._fields_]
    cfg = ns["_Cfg"]()
    ns["_lib"].lmz_default_config(ctypes.byref(cfg))
    assert cfg.struct_size == ctypes.sizeof(ns["_Cfg"])


@pytest.mark.gpu
def test_stub_runs_and_matches_the_package():
    import torch
    import gym_lmaze_b200 as lmz
    ns, _ = load_stub()
    a = ns["LmazeEnvCuda"](300, device=0, seed=11)
    b = lmz.LmazeVecCuda(300, "v0", device="cuda:0", seed=11)
    assert torch.equal(a.reset(), b.reset())
    acts = torch.randint(0, 4, (300,), device="cuda", dtype=torch.uint8)
    oa, ra, da, ia = a.step(acts)
    ob, rb, db, _ = b.step(acts)
    assert torch.equal(oa, ob) and torch.equal(ra.view(torch.int32), rb.view(torch.int32)) and torch.equal(da, db)
    assert ia is acts
    a.close(); b.close()
