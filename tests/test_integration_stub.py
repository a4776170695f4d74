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
    hier = re.search(r"```python\n(# gym_lmaze/envs/lmaze_env_v5_cuda\.py.*?)```", md, re.S).group(1)
    exec(compile(hier, "INTEGRATION.md (v5)", "exec"), ns)
    host = re.search(r"```python\n(# gym_lmaze/envs/lmaze_env_cuda_host\.py.*?)```", md, re.S).group(1)
    exec(compile(host, "INTEGRATION.md (host pipeline)", "exec"), ns)
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


@pytest.mark.gpu
def test_hier_stub_runs_and_matches_the_package():
    import torch
    import gym_lmaze_b200 as lmz
    ns, _ = load_stub()
    a = ns["LmazeHierEnvCuda"](300, device=0, seed=11)
    b = lmz.LmazeHierCuda(300, "v5", device="cuda:0", seed=11)
    assert torch.equal(a.reset(), b.reset())
    gen = torch.Generator(device="cuda").manual_seed(3)
    for _ in range(30):
        goals = torch.randint(0, 25, (300,), generator=gen, device="cuda", dtype=torch.uint8)
        acts = torch.randint(0, 4, (300,), generator=gen, device="cuda", dtype=torch.uint8)
        assert torch.equal(a.plannerStep(goals), b.plannerStep(goals, mask="auto"))
        ta, tb = a.step(acts), b.step(acts)
        for x, y in zip(ta[:7], tb[:7]):
            assert torch.equal(x, y)
        assert ta[7] is acts
    a.close(); b.close()


@pytest.mark.gpu
def test_host_pipeline_stub_runs_and_matches_the_package():
    """The host-consumer binding of INTEGRATION.md: raw pointers, bit-packed observations, lmz_step_host_async / _wait."""
    import torch
    import gym_lmaze_b200 as lmz
    ns, _ = load_stub()
    n = 5000
    a = ns["LmazeHostPipe"](n, device=0, seed=4)
    b = lmz.LmazeVecCuda(n, "v0", device="cuda:0", seed=4)
    b.reset()
    gen = torch.Generator().manual_seed(2)
    acts = [torch.randint(0, 5, (n,), generator=gen, dtype=torch.uint8).pin_memory() for _ in range(12)]
    want = []
    for k, act in enumerate(acts):
        obs, rew, done, _ = b.step(act)
        want.append((obs.cpu().clone(), rew.cpu().clone(), done.cpu().clone()))
        got = a.submit(act)
        assert (got is None) == (k == 0)
        if got is not None:
            o, r, d = got
            wo, wr, wd = want[k - 1]
            assert torch.equal(a.image(o), wo) and torch.equal(r.view(torch.int32), wr.view(torch.int32)) and torch.equal(d.bool(), wd)
    o, r, d = a.drain()
    assert torch.equal(a.image(o), want[-1][0]) and torch.equal(d.bool(), want[-1][2])
    a.close(); b.close()
