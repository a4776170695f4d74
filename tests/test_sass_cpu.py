"""CPU tier: the built library really is sm_100a code that uses the Blackwell paths DESIGN.md claims -- checked on
the SASS of the in-tree .so with cuobjdump (no GPU needed)."""
import os
import re
import shutil
import subprocess

import pytest

CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
pytestmark = pytest.mark.skipif(not os.path.isfile(CUOBJDUMP), reason="cuobjdump not installed")


@pytest.fixture(scope="module")
def sass():
    from gym_lmaze_b200 import build, _abi
    build.build()
    out = subprocess.run([CUOBJDUMP, "-sass", _abi.LIB_PATH], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                         timeout=600).stdout
    funcs = {}
    name = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            funcs[name] = []
        elif name:
            funcs[name].append(line)
    return out, {k: "\n".join(v) for k, v in funcs.items()}


def test_built_for_sm_100a_only(sass):
    out, _ = sass
    archs = set(re.findall(r"arch = (sm_\w+)", out))
    assert archs == {"sm_100a"}, archs


def _one(funcs, *needles):
    hits = [k for k in funcs if all(n in k for n in needles)]
    assert hits, needles
    return funcs[hits[0]]


def test_tma_and_vector_store_paths_are_in_the_sass(sass):
    _, funcs = sass
    tma = _one(funcs, "lmz_env_tma_kernel", "2V0")
    assert "UBLKCP.S.G" in tma            # cp.async.bulk global -> shared: the maze template staged by the TMA engine
    assert "UBLKCP.G.S" in tma            # cp.async.bulk shared -> global: observation segments written by the TMA engine
    assert "SYNCS" in tma                 # mbarrier
    st = _one(funcs, "lmz_env_st_kernel", "2V0")
    assert "STG.E.EF.128" in st           # st.global.cs.v4: 128-bit streaming stores
    fov = _one(funcs, "lmz_env_fov_kernel", "2V5")
    assert "UBLKCP.S.G" in fov and "STG.E.EF.128" in fov     # tables + visit layers by bulk load, float4 render
    assert "DADD" not in fov and "DMUL" not in fov and "DFMA" not in fov     # the visit average runs in float32
    roll = _one(funcs, "lmz_rollout_kernel", "2V0")
    assert "VOTE" in roll or "VOTEU" in roll                  # warp ballots feed the statistics
    assert "HMMA" not in "".join(funcs.values()) and "UTCHMMA" not in "".join(funcs.values())   # no tensor-core use


def _instructions(body):
    return [l for l in body.splitlines() if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l)]


def test_visit_fold_is_a_rolled_ffma2_loop_without_local_memory(sass):
    """Two regressions this guards: the visit fold fully unrolled was 150 KB of code per warp and tile and
    the compact v4 / v5 kernels waited for instruction FETCH (0.85 instead of 4+ G env-steps/s); a store into the history
    words at a runtime index sends them to local memory.  The fold runs on packed FFMA2 pairs."""
    _, funcs = sass
    for needles in (("lmz_fov_small_kernel", "FovILi4ELi7", "Li128"), ("lmz_fov_small_kernel", "2V5", "Li128")):
        body = _one(funcs, *needles)
        n = len(_instructions(body))
        assert n < 8000, (needles, n)                            # rolled: ~4,600 (v4) / ~5,300 (v5); unrolled: ~15,000
        assert len(re.findall(r"\bFFMA2\b", body)) >= 100, needles   # 25 packed fmas per entry x 4 entries per trip
        assert not re.search(r"\b(LDL|STL)\b", body) or needles[1] == "2V5", needles   # v4: no local-memory traffic at all
    fov = _one(funcs, "lmz_env_fov_kernel", "FovILi4ELi7", "Li128")
    assert len(re.findall(r"\bFFMA2\b", fov)) >= 100 and len(_instructions(fov)) < 10000
