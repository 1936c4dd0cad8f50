#!/usr/bin/env python
"""Secondary measurements of SURVEY.md 8(d) -- everything on the path that is NOT bench.py's headline line.

  python tools/bench_secondary.py [--reps 20] > gpurun_out/secondary.jsonl

One JSON line per case:
  * SfmAligner::RunStep at the other BASELINE configs: 160x120 C=8 (configs[0]), 640x480 level 0 at C=64 and
    C=128 (config v), batched so that one launch streams more than the 126 MB L2; device time by CUDA events.
  * the synchronous per-call API (what a PhotometricFactor / CameraTracker pays per call, including the result
    read-back): SfmAligner::RunStep, EvaluateError, SE3Aligner::RunStep, UpdateDepth, SobelGradients,
    GaussianBlurDown at 640x480 -- wall clock per call.
Peak for the roofline fraction: MEASURED_PEAKS.json hbm_gbs (fallback 6650 GB/s).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=20)
    args = ap.parse_args()
    import numpy as np
    import torch

    from deepfactors_b200 import synth
    from deepfactors_b200.aligners import (GaussianBlurDown, SE3Aligner, SfmAligner, SobelGradients, UpdateDepth)

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    ppath = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = float(json.load(open(ppath))["hbm_gbs"]) if os.path.exists(ppath) else 6650.0

    def upload(L):
        d = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in dict(
            img0=L.img0, img1=L.img1, dpt0=L.dpt0, prx0_jac=L.prx_jac, grad1=L.grad1, prx_orig=L.prx_orig).items()}
        d["valid0"] = torch.zeros_like(d["img0"])
        d["cam"] = L.cam
        return d

    def batched(w, h, cs, copies, gram="auto"):
        pair = synth.make_pair(w, h, cs, 1, seed=7)
        L = pair.levels[0]
        al = SfmAligner(cs, gram_mode=gram)
        items = []
        keep = []
        for c in range(copies):
            d = upload(L)
            if c:
                d["prx0_jac"] = torch.roll(d["prx0_jac"], shifts=(3 * c, 5 * c), dims=(0, 1)).contiguous()
            keep.append(d)
            items.append(dict(pose0=pair.pose0, pose1=pair.pose1, cam=L.cam, **{k: d[k] for k in (
                "img0", "img1", "dpt0", "valid0", "prx0_jac", "grad1")}))
        work = al.make_work_items(items)
        rec = al.RunStepBatch(work).clone()
        for _ in range(3):
            al.RunStepBatch(work, rec)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.reps):
            al.RunStepBatch(work, rec)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.reps
        bytes_per_launch = copies * w * h * (24 + 4 * cs)
        gbs = bytes_per_launch / (ms * 1e-3) / 1e9
        inl = al.unpack(rec)[0].inliers
        print(json.dumps({"case": f"SfmAligner::RunStep {w}x{h} C={cs} x{copies} items/launch gram={gram}",
                          "ms_per_launch": ms, "evals_per_s": copies / (ms * 1e-3),
                          "algorithmic_MB_per_launch": bytes_per_launch / 1e6, "GBps": gbs, "frac_of_hbm_peak": gbs / peak,
                          "inlier_fraction": inl / (w * h), "timing": "cuda events, includes the finalize kernel"}),
              flush=True)

    batched(160, 120, 8, 256)
    batched(640, 480, 32, 8, "fp32")
    batched(640, 480, 32, 8, "tf32x3")
    batched(640, 480, 64, 4)
    batched(640, 480, 128, 3)

    # ---- synchronous per-call API at 640x480 ---------------------------------------------------------------
    pair = synth.make_pair(640, 480, 32, 1, seed=7)
    L = pair.levels[0]
    d = upload(L)
    al = SfmAligner(32)
    se = SE3Aligner()
    code = pair.code
    dpt_out = torch.empty_like(d["dpt0"])
    grad_out = torch.empty_like(d["grad1"])
    half = torch.empty((240, 320), dtype=torch.float32, device=dev)

    def timed(name, fn, bytes_per_call):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.reps * 5):
            fn()
        torch.cuda.synchronize()
        us = (time.perf_counter() - t0) / (args.reps * 5) * 1e6
        print(json.dumps({"case": name + " 640x480 (synchronous call, wall clock)", "us_per_call": us,
                          "algorithmic_MB_per_call": bytes_per_call / 1e6,
                          "GBps": bytes_per_call / (us * 1e-6) / 1e9}), flush=True)

    px = 640 * 480
    timed("SfmAligner::RunStep C=32", lambda: al.RunStep(pair.pose0, pair.pose1, code, L.cam, d["img0"], d["img1"],
                                                         d["dpt0"], None, d["valid0"], d["prx0_jac"], d["grad1"]),
          px * 152)
    timed("SfmAligner::EvaluateError", lambda: al.EvaluateError(pair.pose0, pair.pose1, L.cam, d["img0"], d["img1"],
                                                                d["dpt0"], None, d["grad1"]), px * 12)
    timed("SE3Aligner::RunStep", lambda: se.RunStep(pair.pose1, L.cam, d["img0"], d["img1"], d["dpt0"], d["grad1"]),
          px * 20)
    timed("UpdateDepth C=32", lambda: UpdateDepth(code, d["prx_orig"], d["prx0_jac"], 2.0, dpt_out), px * (8 + 128))
    timed("SobelGradients", lambda: SobelGradients(d["img1"], grad_out), px * 12)
    timed("GaussianBlurDown", lambda: GaussianBlurDown(d["img1"], half), px * 5)

    # ---- CameraTracker::TrackFrame: 3 levels x (10, 5, 4) iterations (data/flags defaults: 19 SE3 steps per frame) --------
    from deepfactors_b200 import se3
    from deepfactors_b200.aligners import CameraTracker, TrackerConfig
    tp = synth.make_pair(640, 480, 8, 3, seed=9)
    lv = [upload(L) for L in tp.levels]
    cams = [L.cam for L in tp.levels]
    iters = (10, 5, 4)
    trk = CameraTracker(cams, TrackerConfig(pyramid_levels=3, iterations_per_level=iters, huber_delta=0.1))
    trk.SetKeyframe([x["img0"] for x in lv], [x["dpt0"] for x in lv])
    img1s, grads = [x["img1"] for x in lv], [x["grad1"] for x in lv]

    def device_loop():
        trk.Reset()
        trk.TrackFrame(img1s, grads)

    def host_loop():  # what the reference does: synchronous step, host solve, next step (camera_tracker.cpp:48-63)
        pose = se3.identity(np.float64)
        for level in (2, 1, 0):
            for _ in range(iters[level]):
                r = se.RunStep(pose.astype(np.float32), cams[level], lv[level]["img0"], lv[level]["img1"], lv[level]["dpt0"],
                               lv[level]["grad1"])
                pose = se3.se3_solve_and_update(r.toDenseMatrix(), r.Jtr, pose)
        return pose

    bytes_track = sum(iters[l] * (640 >> l) * (480 >> l) * 20 for l in range(3))
    timed("CameraTracker::TrackFrame 19 iterations, device-side loop (dfk_se3_track)", device_loop, bytes_track)
    timed("CameraTracker::TrackFrame 19 iterations, per-step API + host solve (reference structure)", host_loop, bytes_track)


if __name__ == "__main__":
    main()
