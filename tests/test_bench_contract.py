"""bench.py's JSON contract on the arm that runs without a GPU (--impl reference)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("arm", ["auto", "port"])
def test_reference_arm_json_line(arm):
    from oracle import ref_cython

    # OMP_NUM_THREADS=1 is what torchrun exports; it must not cap the arm
    env = dict(os.environ, OMP_NUM_THREADS="1", TJB_BENCH_CPU_ARM=arm)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                        "--steps", "1", "--warmup", "1", "--log2-ref-step", "9"],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step",
              "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config", "e2e",
              "cpu_baseline", "impl"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "samples/s" and d["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    kind = "reference" if (arm == "auto" and ref_cython.available()) else "port"
    assert cb["kind"] == kind and cb["value"] == d["value"] and cb["cores"] >= 1
    assert cb["cores"] == len(os.sched_getaffinity(0))
    assert "workload" in d["config"] and d["vs_baseline"] is None


def test_reference_arm_nonzero_rank_is_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                        "--gpus", "2", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0 and r.stdout.strip() == ""
