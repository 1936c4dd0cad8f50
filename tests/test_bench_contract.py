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
    # oracle/_ref (the reference's own headers, compiled here) when present, else the oracle port
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["value"] == d["value"]
    host = d["cpu_baseline"]["host"]
    assert host["affinity"] == len(os.sched_getaffinity(0)) and 1 <= host["usable"] <= host["affinity"]
    # throughput mode: the fastest of the swept thread counts, never more than the usable cores, and OMP_NUM_THREADS=1
    # (what torchrun exports) must not pin it to one thread when more cores are usable
    assert 1 <= d["cpu_baseline"]["cores"] <= host["usable"]
    assert d["cpu_baseline"]["threads_used"] == d["cpu_baseline"]["cores"] == d["config"]["evals_per_step"]
    assert "1" in d["cpu_baseline"]["thread_sweep_evals_per_s"] and d["cpu_baseline"]["single_thread"]["value"] > 0
    if host["usable"] >= 4:
        assert d["cpu_baseline"]["cores"] > 1
    assert d["e2e"] == {"value": d["value"], "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["gpu_launches"] == 0 and d["vs_baseline"] is None


def test_reference_arm_is_silent_on_other_ranks():
    assert run_bench({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}) == []
