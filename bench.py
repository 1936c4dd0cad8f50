#!/usr/bin/env python
"""bench.py -- keyframe-pair Jacobian+JtJ evaluations per second (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (libdfk.so)
  python bench.py --impl reference --steps K --warmup W    # the reference's CPU path (oracle port, all host threads)

Workload (BASELINE.json configs[1]): one evaluation = SfmAligner::RunStep over the 4-level pyramid
(640x480 ... 80x60, 408 000 px) of one keyframe/frame pair at code size 32, synthetic data
(deepfactors_b200/synth.py), fp32.  One STEP = `--pairs-per-step` (default 8) such evaluations of DISTINCT
pairs submitted as one persistent launch (8 pairs = 496 MB of inputs > the 126 MB L2, so every step streams its
inputs from HBM); at N > 1 every rank evaluates its own pairs (weak scaling: pairs shard across GPUs with no
data-path collective) and the per-pair normal equations are summed into the window's Hessian buffer with one NCCL
all-reduce per step.  `value` = evaluations of all ranks / max-over-ranks device time.

Keys beyond the base contract: `roofline` (dominant kernel = sfm_step kernel; achieved = algorithmic bytes per
launch / CUDA-event time of the kernel launches in the timed region; peak = MEASURED_PEAKS.json hbm_gbs),
`cpu_baseline` (oracle port, OpenMP, all host cores, bounded sample), `e2e` (same metric through the synchronous
C-ABI call with every input uploaded from pinned host memory and the result read back, each step),
`single_launch` (one pair per launch: the latency-bound regime of the per-factor API).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "keyframe-pair Jacobian+JtJ evals/sec (640x480, C=32)"
W0, H0, CS, LEVELS = 640, 480, 32, 4
PIXELS = sum((W0 >> l) * (H0 >> l) for l in range(LEVELS))  # 408000
BYTES_PER_PX = 24 + 4 * CS
BYTES_PER_EVAL = PIXELS * BYTES_PER_PX                      # 62.02 MB


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs-per-step", type=int, default=8)
    ap.add_argument("--gram", default="auto", choices=["auto", "fp32", "tf32x3"])
    ap.add_argument("--e2e-steps", type=int, default=20)
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--code-sigma", type=float, default=0.0,
                    help="std of the latent code used to decode dpt0 (0 = the reference test's zero code, smooth depth; "
                         ">0 adds per-pixel depth noise through the iid synthetic code Jacobian)")
    ap.add_argument("--fused-depth", action="store_true",
                    help="decode dpt0 from prx_orig + code inside the launch (UpdateDepth + RunStep in one pass; "
                         "28+4C algorithmic bytes per pixel instead of 24+4C, and no separate UpdateDepth pass)")
    ap.add_argument("--identity-pose", action="store_true", help="100%% inliers (worst-case work) instead of the ~60%% of the reference test poses")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region (B200_PROFILING.md clocks line).

    The timed region of the default run is tens of milliseconds, shorter than one `nvidia-smi -lms` period, so the
    samples come from NVML directly (nvidia_ml_py), polled from a thread about every millisecond between start()
    and stop(); `nvidia-smi --query-gpu` is only the fallback when NVML cannot be loaded."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index: int):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        try:
            ids = [int(x) for x in vis.split(",")] if vis else []
            self.index = ids[index] if index < len(ids) else index
        except ValueError:
            self.index = index
        self.samples, self.masks = [], []
        self.max_mhz = None
        self.nvml = self.handle = self.thread = None
        self.running = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nvml = None

    def _poll_once(self):
        n = self.nvml
        self.samples.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
        try:
            fn = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
            self.masks.append(int(fn(self.handle)))
        except Exception:
            pass

    def _loop(self):
        while self.running:
            try:
                self._poll_once()
            except Exception:
                break
            time.sleep(0.001)

    def start(self):
        if self.nvml is None:
            return
        self.running = True
        self.thread = threading.Thread(target=self._loop, daemon=True)
        self.thread.start()

    def stop(self):
        if self.nvml is None:
            return self._smi_fallback()
        self.running = False
        if self.thread is not None:
            self.thread.join(timeout=2)
        sm = sorted(self.samples)
        reasons = set()
        for m in self.masks:
            for bit, name in self.REASONS.items():
                if m & bit:
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(reasons),
                "samples": len(sm), "source": "NVML polled during the timed region"}

    def _smi_fallback(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i",
                                  str(self.index)], capture_output=True, text=True, timeout=10).stdout.strip()
            f = [x.strip() for x in out.split(",")]
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            reasons = [n for n, v in zip(names, f[2:6]) if v.lower().startswith("active")]
            return {"sm_mhz": float(f[0]), "sm_max_mhz": float(f[1]), "reasons": reasons, "samples": 1,
                    "source": "nvidia-smi one-shot right after the timed region (NVML unavailable)"}
        except Exception:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}


# ------------------------------------------------------------------------------------------------ CPU arm
def host_threads() -> int:
    """every host core this process may run on -- NOT omp_get_max_threads(): torchrun exports OMP_NUM_THREADS=1"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_eval_seconds(orc, pair, threads):
    t0 = time.perf_counter()
    for L in pair.levels:
        orc.sfm_run_step(pair.pose0, pair.pose1, L.cam, L.img0, L.img1, L.dpt0, None, L.prx_jac, L.grad1,
                         omp_threads=threads)
    return time.perf_counter() - t0


def cpu_baseline(budget_s: float, code_sigma: float = 0.0, identity_pose: bool = False):
    """oracle port (fp32, row-major OpenMP over all host cores) on a bounded sample of the same workload"""
    from deepfactors_b200 import synth
    from oracle import oracle as orc
    orc.build()
    pair = synth.make_pair(W0, H0, CS, LEVELS, seed=0, code_sigma=code_sigma, identity_pose=identity_pose)
    threads = host_threads()
    cpu_eval_seconds(orc, pair, threads)  # warm-up
    ts = []
    t_end = time.perf_counter() + budget_s
    while len(ts) < 3 or (time.perf_counter() < t_end and len(ts) < 5000):  # bounded by wall time, not by count
        ts.append(cpu_eval_seconds(orc, pair, threads))
    ts.sort()
    reps = len(ts)
    med = ts[len(ts) // 2]
    # also the reference test's own structure -- one thread, x outer / y inner (ut_sfmaligner.cpp:303-315)
    t0 = time.perf_counter()
    for L in pair.levels:
        orc.sfm_run_step(pair.pose0, pair.pose1, L.cam, L.img0, L.img1, L.dpt0, None, L.prx_jac, L.grad1, loop_order=0)
    single = 1.0 / (time.perf_counter() - t0)
    return {"value": 1.0 / med, "unit": "evals/s", "cores": threads, "kind": "port",
            "single_thread_reference_loop_order": {"value": single, "unit": "evals/s", "cores": 1,
                                                   "sample": "1 evaluation, x outer / y inner as ut_sfmaligner.cpp:303-315"},
            "sample": f"{reps} evaluations of one 640x480 4-level C=32 pair (median of {reps}, "
                      f"{sum(ts):.1f} s of CPU work), oracle fp32 OpenMP row-major"}, pair


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from deepfactors_b200 import synth
    from oracle import oracle as orc
    orc.build()
    pair = synth.make_pair(W0, H0, CS, LEVELS, seed=0, code_sigma=args.code_sigma, identity_pose=args.identity_pose)
    threads = host_threads()
    for _ in range(max(1, min(args.warmup, 3))):
        cpu_eval_seconds(orc, pair, threads)
    steps = max(1, args.steps)
    # each step = one evaluation (a bounded sample of the GPU arm's step); cap the total at a few minutes
    t1 = cpu_eval_seconds(orc, pair, threads)
    steps = min(steps, max(3, int(150.0 / max(t1, 1e-3))))
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_eval_seconds(orc, pair, threads)
    dt = time.perf_counter() - t0
    val = steps / dt
    out = {"impl": "reference", "metric": METRIC, "value": val, "unit": "evals/s", "n_gpus": args.gpus,
           "steps": steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": "single pair 640x480 4-level pyramid, code dim 32 (BASELINE configs[1])",
                      "evals_per_step": 1, "note": "reference CPU path = oracle port of df::DenseSfm "
                      "(the reference itself cannot be compiled: Eigen/Sophus/VisionCore absent)"},
           "cpu_baseline": {"value": val, "unit": "evals/s", "cores": threads, "kind": "port",
                            "sample": f"{steps} evaluations, one per step"},
           "e2e": {"value": val, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out))


# ------------------------------------------------------------------------------------------------ GPU arm
def main():
    args = parse()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    from deepfactors_b200 import _lib, synth
    from deepfactors_b200.aligners import SfmAligner

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"  # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
        dist.init_process_group("nccl", device_id=dev)
    n_gpus = world

    P = args.pairs_per_step
    # ---- synthetic window: P distinct pairs resident in HBM -------------------------------------------
    base = synth.make_pair(W0, H0, CS, LEVELS, seed=rank, code_sigma=args.code_sigma, identity_pose=args.identity_pose)
    host_levels = []
    for L in base.levels:
        host_levels.append(dict(img0=L.img0, img1=L.img1, dpt0=L.dpt0, prx0_jac=L.prx_jac, grad1=L.grad1))
        if args.fused_depth:
            host_levels[-1]["prx_orig"] = L.prx_orig
    pairs_dev = []
    for p in range(P):
        lv = []
        for L, hl in zip(base.levels, host_levels):
            d = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in hl.items()}
            if p > 0:  # distinct contents per pair (device-side variation of the host-generated base pair)
                d["prx0_jac"] = torch.roll(d["prx0_jac"], shifts=(3 * p, 5 * p), dims=(0, 1)).contiguous()
                d["img0"] = (d["img0"] * (1.0 - 0.01 * p)).contiguous()
            d["valid0"] = torch.zeros_like(d["img0"])
            d["cam"] = L.cam
            lv.append(d)
        pairs_dev.append(lv)

    al = SfmAligner(CS, gram_mode=args.gram)
    items = []
    for lv in pairs_dev:
        for d in lv:
            items.append(dict(pose0=base.pose0, pose1=base.pose1, cam=d["cam"], img0=d["img0"], img1=d["img1"],
                              dpt0=d["dpt0"], valid0=d["valid0"], prx0_jac=d["prx0_jac"], grad1=d["grad1"]))
            if args.fused_depth:
                items[-1].update(prx_orig=d["prx_orig"], code=base.code)
    work = al.make_work_items(items)
    rec_floats = _lib.record_floats(CS)
    # window Hessian buffer: every rank owns P*LEVELS rows; all-reduce(sum) assembles the window
    hess = torch.zeros((world * P * LEVELS, rec_floats), dtype=torch.float32, device=dev)
    my_rows = hess[rank * P * LEVELS:(rank + 1) * P * LEVELS]

    def step():
        al.RunStepBatch(work, my_rows)
        if world > 1:
            dist.all_reduce(hess)

    lib = _lib.lib()
    import ctypes as C

    def read_profile():
        ms, n, tot = C.c_double(0), C.c_uint64(0), C.c_uint64(0)
        _lib.check(al.handle, lib.dfk_get_profile(al.handle, C.byref(ms), C.byref(n), C.byref(tot)))
        return ms.value, n.value, tot.value

    for _ in range(max(3, args.warmup)):
        step()
    torch.cuda.synchronize()
    read_profile()
    _lib.check(al.handle, lib.dfk_set_profiling(al.handle, 1))

    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    clocks = sampler.stop() if rank == 0 else None
    total_ms = float(ms.item())
    kern_ms, kern_n, launches = read_profile()
    _lib.check(al.handle, lib.dfk_set_profiling(al.handle, 0))

    evals = world * P * args.steps
    value = evals / (total_ms * 1e-3)

    # ---- single pair per launch (latency-bound regime), rotating over the P pairs ----------------------
    single = None
    works1 = [al.make_work_items(items[p * LEVELS:(p + 1) * LEVELS]) for p in range(P)]
    recs1 = torch.empty((LEVELS, rec_floats), dtype=torch.float32, device=dev)
    for p in range(P):
        al.RunStepBatch(works1[p], recs1)
    torch.cuda.synchronize()
    n1 = max(50, 4 * args.steps)
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for i in range(n1):
        al.RunStepBatch(works1[i % P], recs1)
    s1.record()
    torch.cuda.synchronize()
    single_ms = s0.elapsed_time(s1) / n1
    single = {"pairs_per_launch": 1, "value": 1e3 / single_ms, "unit": "evals/s", "ms_per_eval": single_ms,
              "frac_of_hbm_roofline": None}

    # ---- e2e: synchronous reference-facing calls, every input uploaded from pinned host memory ----------
    # the inputs of one evaluation live in ONE pinned host allocation and travel as ONE copy into one device allocation
    # (the per-level tensors are 256-byte aligned views of it): 20 separate cudaMemcpyAsync calls cost ~40 % of the PCIe rate
    in_keys = [k for k in host_levels[0].keys() if not (args.fused_depth and k == "dpt0")]  # fused: dpt0 is an output
    layout, total = [], 0
    for hl in host_levels:
        ent = {}
        for k in in_keys:
            n = int(np.asarray(hl[k]).size)
            ent[k] = (total, n, tuple(np.asarray(hl[k]).shape))
            total += (n + 63) // 64 * 64
        layout.append(ent)
    host_blob = torch.empty(total, dtype=torch.float32).pin_memory()
    dev_blob = torch.empty(total, dtype=torch.float32, device=dev)
    stage = []
    for hl, ent in zip(host_levels, layout):
        sd = {}
        for k, (off, n, shape) in ent.items():
            host_blob[off:off + n].copy_(torch.from_numpy(np.ascontiguousarray(hl[k], dtype=np.float32)).reshape(-1))
            sd[k] = dev_blob[off:off + n].view(shape)
        if "dpt0" not in sd:
            sd["dpt0"] = torch.empty(tuple(np.asarray(hl["dpt0"]).shape), dtype=torch.float32, device=dev)
        sd["valid0"] = torch.zeros_like(sd["img0"])
        stage.append(sd)
    h2d = sum(n * 4 for ent in layout for (_, n, _) in ent.values())
    d2h = LEVELS * rec_floats * 4

    e2e_items = [dict(pose0=base.pose0, pose1=base.pose1, cam=L.cam, img0=sd["img0"], img1=sd["img1"], dpt0=sd["dpt0"],
                      valid0=sd["valid0"], prx0_jac=sd["prx0_jac"], grad1=sd["grad1"],
                      **(dict(prx_orig=sd["prx_orig"], code=base.code) if args.fused_depth else {}))
                 for L, sd in zip(base.levels, stage)]
    e2e_work = al.make_work_items(e2e_items)
    e2e_rec_dev = torch.empty((LEVELS, rec_floats), dtype=torch.float32, device=dev)
    e2e_rec_host = torch.empty((LEVELS, rec_floats), dtype=torch.float32).pin_memory()

    def e2e_step():
        # every input of the evaluation travels host -> device (pinned, async on the launch stream), one batched
        # C-ABI launch evaluates the 4 levels, the 4 result records travel back and the host waits for them
        dev_blob.copy_(host_blob, non_blocking=True)
        al.RunStepBatch(e2e_work, e2e_rec_dev)
        e2e_rec_host.copy_(e2e_rec_dev, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return e2e_rec_host

    e2e_step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_dt = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        dist.all_reduce(e2e_dt, op=dist.ReduceOp.MAX)
    e2e_value = world * args.e2e_steps / float(e2e_dt.item())

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak = float(json.load(open(peaks_path))["hbm_gbs"])
            peak_src = "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
        else:
            peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
        bytes_per_eval = PIXELS * (BYTES_PER_PX + (4 if args.fused_depth else 0))
        bytes_per_launch = P * bytes_per_eval
        traffic = None  # dram__bytes_read+write of one step-kernel launch, from the committed ncu --set full capture
        tpath = os.path.join(ROOT, "profiles", "r01_traffic.json")
        if os.path.exists(tpath) and P == 8 and args.gram in ("auto", "tf32x3") and not args.fused_depth:
            traffic = json.load(open(tpath))["traffic_bytes_per_launch"]
        kern_avg_ms = kern_ms / max(kern_n, 1)
        achieved = bytes_per_launch / (kern_avg_ms * 1e-3) / 1e9 if kern_n else None
        single["frac_of_hbm_roofline"] = (bytes_per_eval / (single_ms * 1e-3) / 1e9) / peak
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            cpu, _ = cpu_baseline(args.cpu_seconds, args.code_sigma, args.identity_pose)
        out = {
            "metric": METRIC, "value": value, "unit": "evals/s", "n_gpus": n_gpus, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "single pair 640x480 4-level pyramid, code dim 32 (BASELINE configs[1]); "
                                   f"{P} distinct pairs per step in one persistent launch",
                       "evals_per_step_per_gpu": P, "pixels_per_eval": PIXELS,
                       "algorithmic_bytes_per_eval": bytes_per_eval, "gram": args.gram,
                       "fused_depth_decode": bool(args.fused_depth),
                       "poses": "identity (100% inliers)" if args.identity_pose else
                                "tests/ut_sfmaligner.cpp:254-264 (~60% inliers)",
                       "code_sigma": args.code_sigma,
                       "l2": f"inputs larger than L2: each step streams {P * BYTES_PER_EVAL / 1e6:.0f} MB of distinct "
                             "pair data (> 126 MB L2)",
                       "parallelism": f"pairs sharded over {n_gpus} GPU(s)" + (
                           "; one NCCL all-reduce of the window's normal equations per step" if world > 1 else "")},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": (achieved / peak) if achieved else None, "traffic": traffic,
                         "kernel": "sfm_step kernel (per-tile warp + Gram)", "launches_timed": kern_n,
                         "avg_launch_ms": kern_avg_ms, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": bytes_per_launch},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": "evals/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": args.e2e_steps, "note": "per step: every level's img0/img1/dpt0/prx_jac/grad1 copied "
                    "from pinned host memory (one packed allocation, one async copy on the launch stream), one "
                    "dfk_sfm_run_step_batch call for the 4 levels, the 4 result records copied back and waited for"},
            "single_launch": single,
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
