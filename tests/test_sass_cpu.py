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
