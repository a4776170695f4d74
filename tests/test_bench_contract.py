"""CPU tier: bench.py's JSON contract on the arm that runs without a GPU (`--impl reference`, the CPU oracle port)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


@pytest.mark.parametrize("variant", ["v0", "v5"])
def test_reference_arm_prints_one_contract_line(variant):
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                          "--variant", variant], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1                                   # ONE JSON line on stdout, nothing else
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "env_steps_per_sec" and d["unit"] == "env-steps/s"
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 1 and d["n_gpus"] == 1 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_other_ranks_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "1"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=120, env=env)
    assert res.returncode == 0 and res.stdout.strip() == ""


def test_default_warmup_puts_the_foveal_variants_in_steady_state(monkeypatch):
    """bench.py's own arm warms v2 / v4 / v5 for 1,000 steps unless told otherwise (their episodes start together and need
    that long to de-synchronise); everything else, and the CPU arm, keeps the short default."""
    sys.path.insert(0, ROOT)
    import bench

    def parsed(*argv):
        monkeypatch.setattr(sys, "argv", ["bench.py", *argv])
        return bench.parse_args()
    assert parsed().warmup == 5 and parsed("--variant", "v3").warmup == 5
    for v in ("v2", "v4", "v5"):
        assert parsed("--variant", v).warmup == 1000
        assert parsed("--variant", v, "--impl", "reference").warmup == 5
        assert parsed("--variant", v, "--warmup", "7").warmup == 7
    assert "steady_state" in bench.workload_config(parsed("--variant", "v4"))
    assert "steady_state" not in bench.workload_config(parsed())
