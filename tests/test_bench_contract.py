"""bench.py contract checks that run without a GPU: the reference arm's JSON line (the CPU restatement timed on the host
cores) carries the keys the driver reads, and only rank 0 prints it."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(extra_env=None):
    env = dict(os.environ)
    env.update(extra_env or {})
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1"], capture_output=True, text=True, env=env, timeout=600, check=True)
    return [ln for ln in out.stdout.splitlines() if ln.strip()]


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    lines = run_bench({"OMP_NUM_THREADS": "1"})  # what torchrun exports; the arm must still use every core
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "evals/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("keyframe-pair Jacobian+JtJ evals/sec")
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["steps"] >= 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"]
    assert d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))
    assert d["e2e"] == {"value": d["value"], "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["gpu_launches"] == 0 and d["vs_baseline"] is None


def test_reference_arm_is_silent_on_other_ranks():
    assert run_bench({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}) == []
