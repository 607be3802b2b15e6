"""bench.py's reference arm (CPU, no GPU needed) prints ONE JSON line with the contract's keys; under a multi-rank
launch only rank 0 works."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra, *args):
    env = dict(os.environ, **env_extra)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", *args], capture_output=True,
                       text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    return [ln for ln in r.stdout.splitlines() if ln.startswith("{")]


def test_reference_arm_line():
    lines = _run({}, "--workload", "fft1d_2p20", "--steps", "2", "--warmup", "1")
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "GFLOP/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["e2e"]["value"] == d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and "jt_ref.c" in cb["sample"] and cb["pocketfft"]["value"] > 0
    assert d["config"]["workload"].startswith("DoubleFFT_1D")


def test_reference_arm_other_ranks_idle():
    lines = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, "--workload", "fft1d_2p20", "--gpus", "2")
    assert lines == []
