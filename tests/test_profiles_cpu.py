"""CPU tier: the committed profile evidence belongs to the CURRENT kernel sources ("make staleness
impossible").  profiles/traffic.json (dram bytes per launch, read by bench.py for roofline.traffic) and
profiles/sass_summary.txt (per-kernel opcode histogram) both record gym_lmaze_b200.build.source_hash() -- the sha256
over csrc/ + include/ -- of the sources they were made from; a kernel change without a re-profile fails here."""
import json
import os
import re

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def test_traffic_json_was_captured_on_the_current_kernel_sources():
    from gym_lmaze_b200.build import source_hash
    doc = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    assert doc.get("csrc_hash") == source_hash(), (
        "profiles/traffic.json was captured on other kernel sources: re-run tools/profile_final.py under ncu on the GPU box "
        "and regenerate it with tools/ncu_summary.py")
    for key in ("v0_tma", "v3_tma", "v2_tma", "v4_tma", "v5_tma"):
        rec = doc[key]
        assert rec["bytes"] > 0 and rec["envs"] > 0 and "launch" in rec["source"]
    # the headline kernel writes what the algorithm says and reads next to nothing: traffic / algorithmic within 1 %
    v0 = doc["v0_tma"]
    assert abs(v0["bytes"] / (v0["envs"] * 112910.0) - 1.0) < 0.01


def test_sass_summary_matches_the_current_kernel_sources():
    from gym_lmaze_b200.build import source_hash
    text = open(os.path.join(ROOT, "profiles", "sass_summary.txt")).read()
    first = text.splitlines()[0]
    assert re.search(r"csrc_hash (\w+)", first).group(1) == source_hash(), (
        "profiles/sass_summary.txt is stale: python tools/sass_summary.py > profiles/sass_summary.txt")
    assert "archs sm_100a " in first
    rows = {l.split("(")[0].strip(): l for l in text.splitlines()[2:]}
    tma = next(l for k, l in rows.items() if k.startswith("lmz_env_tma_kernel<V0, 32>"))
    cols = text.splitlines()[1].split()
    def count(line, op):
        return int(line.split()[-(len(cols) - cols.index(op)):][0])
    assert count(tma, "UBLKCP.S.G") >= 1 and count(tma, "UBLKCP.G.S") >= 1 and count(tma, "SYNCS") >= 1
    for l in rows.values():
        assert count(l, "HMMA") == 0 and count(l, "DFMA") == 0 and count(l, "DADD") == 0


def test_source_hash_covers_every_kernel_source():
    """Every file under csrc/ and include/ takes part in the hash (and in the build's staleness check)."""
    from gym_lmaze_b200 import build
    listed = {os.path.abspath(f) for f in build.SOURCES + build.HEADERS}
    on_disk = set()
    for d in (os.path.join(ROOT, "gym_lmaze_b200", "csrc"), os.path.join(ROOT, "include")):
        on_disk |= {os.path.abspath(os.path.join(d, f)) for f in os.listdir(d) if f.endswith((".cu", ".cuh", ".h"))}
    assert on_disk == listed, sorted(on_disk ^ listed)
