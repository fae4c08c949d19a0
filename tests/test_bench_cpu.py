"""CPU checks of bench.py's contract: the reference arm (the unmodified reference from baseline/_ref when it is
importable, else the C restatement, on the host cores) runs without a GPU and prints one JSON line with the keys the
driver reads."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REQUIRED = ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e")


@pytest.mark.parametrize("workload,extra", [("resolve", ["--port"]), ("skytem", []), ("tempest", [])])
def test_reference_arm_json_line(workload, extra):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", workload,
                          "--steps", "1", "--warmup", "1"] + extra, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.strip().splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in REQUIRED:
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "evals/s" and d["higher_is_better"] is True
    assert d["value"] > 100.0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert "workload" in d["config"] and d["vs_baseline"] is None


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "1"], capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_reference_arm_runs_the_unmodified_reference_when_installed():
    """With baseline/_ref present (the offline pip --target install of the reference) the arm times
    the reference's own Inference1D loop; without it the same command falls back to the C port."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads([ln for ln in out.stdout.strip().splitlines() if ln.startswith("{")][0])
    for k in REQUIRED:
        assert k in d, k
    have = os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "geobipy", "src")) or os.path.isdir("/root/reference/geobipy/src")
    assert d["cpu_baseline"]["kind"] == ("reference" if have else "port")
    if have:   # the Python reference runs ~100-130 iterations per second per core; the port rides along
        assert 20.0 < d["value"] / d["cpu_baseline"]["cores"] < 2000.0 and d["cpu_port"]["kind"] == "port"
        assert d["cpu_port"]["value"] > d["value"]
