"""ctypes front-end of oracle/_ref/libdfk_ref.so: the REFERENCE's own per-pixel headers (dense_sfm.h, warping.h,
pinhole_camera_impl.h, lucas_kanade_se3.h, m_estimators.h, reduction_items.h) compiled unmodified from /root/reference
against the stand-in Eigen / Sophus / VisionCore headers of oracle/shim/ (oracle/ref_driver.cpp, `make -C oracle ref`).

TEST INFRASTRUCTURE ONLY: tests/test_oracle_ref.py pins the hand-written oracle against it, bench.py may time it as the
CPU baseline ("kind": "reference").  The library is built in the dev container (where /root/reference exists) and
travels to the GPU box as a prebuilt file; nothing here reads /root/reference at run time.

Same array conventions as oracle/oracle.py.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from .oracle import Camera, SfmParams, StepResult, _cam, _f32, _pitch, _ptr, default_params, make_levels

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libdfk_ref.so")
REFERENCE_ROOT = "/root/reference"

_lib = None


def available() -> bool:
    return os.path.exists(LIB_PATH) or os.path.isdir(os.path.join(REFERENCE_ROOT, "sources"))


def build(force: bool = False) -> str:
    """Compile the reference headers where they lie (needs /root/reference: the dev container only)."""
    srcs = [os.path.join(_HERE, "ref_driver.cpp"), os.path.join(_HERE, "Makefile")]
    stale = (not os.path.exists(LIB_PATH)) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs)
    if (force or stale) and os.path.isdir(os.path.join(REFERENCE_ROOT, "sources")):
        subprocess.run(["make", "-C", _HERE, "-s", "ref"], check=True)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        if not os.path.exists(LIB_PATH):
            raise FileNotFoundError(f"{LIB_PATH} is missing and {REFERENCE_ROOT} is not present to build it from")
        _lib = C.CDLL(LIB_PATH)
        _lib.dfkr_huber_weight_f.restype = C.c_float
        _lib.dfkr_huber_weight_f.argtypes = [C.c_float, C.c_float]
        _lib.dfkr_sfm_run_step_f.restype = C.c_int
        _lib.dfkr_update_depth_f.restype = C.c_int
    return _lib


def relative_pose(a, b, dtype=np.float64):
    a = np.ascontiguousarray(a, dtype=dtype)
    b = np.ascontiguousarray(b, dtype=dtype)
    ab, ja, jb = np.zeros(7, dtype=dtype), np.zeros(36, dtype=dtype), np.zeros(36, dtype=dtype)
    ct = C.c_double if dtype == np.float64 else C.c_float
    fn = lib().dfkr_relative_pose_d if dtype == np.float64 else lib().dfkr_relative_pose_f
    fn(_ptr(a, ct), _ptr(b, ct), _ptr(ab, ct), _ptr(ja, ct), _ptr(jb, ct))
    return ab, ja.reshape(6, 6), jb.reshape(6, 6)


def probe_pixel(x, y, dpt, cam, pose, border=1, min_dpt=0.0, avg_dpt=2.0):
    pose = np.ascontiguousarray(pose, dtype=np.float64)
    out = np.zeros(17, dtype=np.float64)
    c = _cam(cam)
    lib().dfkr_probe_pixel_d(C.c_size_t(int(x)), C.c_size_t(int(y)), C.c_double(dpt), C.byref(c), _ptr(pose, C.c_double),
                             C.c_int(border), C.c_double(min_dpt), C.c_double(avg_dpt), _ptr(out, C.c_double))
    return dict(valid=bool(out[0]), pix1=out[1:3].copy(), J_pose=out[3:15].reshape(2, 6).copy(), J_prx=out[15:17].copy())


def huber_weight(x, delta):
    return float(lib().dfkr_huber_weight_f(C.c_float(x), C.c_float(delta)))


def sfm_run_step(pose0, pose1, cam, img0, img1, dpt0, valid0, prx0_jac, grad1, params=None) -> StepResult:
    """tests/ut_sfmaligner.cpp:297-315: RelativePose + the x-outer / y-inner loop over df::DenseSfm, fp32."""
    params = params or default_params()
    img0, img1, dpt0, prx0_jac, grad1 = map(_f32, (img0, img1, dpt0, prx0_jac, grad1))
    H, W = img0.shape
    Cs = prx0_jac.shape[2]
    NP = 12 + Cs
    pose0 = np.ascontiguousarray(pose0, dtype=np.float32)
    pose1 = np.ascontiguousarray(pose1, dtype=np.float32)
    if valid0 is None:
        valid0 = np.zeros((H, W), dtype=np.float32)
    c = _cam(cam)
    inl, res = C.c_uint64(0), C.c_float(0)
    JtJ = np.zeros(NP * (NP + 1) // 2, dtype=np.float32)
    Jtr = np.zeros(NP, dtype=np.float32)
    rc = lib().dfkr_sfm_run_step_f(_ptr(pose0), _ptr(pose1), C.c_int(Cs), C.byref(c), C.c_int(W), C.c_int(H), _ptr(img0),
                                   _pitch(img0), _ptr(img1), _pitch(img1), _ptr(dpt0), _pitch(dpt0), _ptr(_f32(valid0)),
                                   _pitch(valid0), _ptr(prx0_jac), _pitch(prx0_jac), _ptr(grad1), _pitch(grad1),
                                   C.byref(params), _ptr(JtJ), _ptr(Jtr), C.byref(res), C.byref(inl))
    if rc != 0:
        raise ValueError(f"code size {Cs} is not instantiated in oracle/ref_driver.cpp")
    return StepResult(JtJ, Jtr, float(res.value), int(inl.value))


def sfm_evaluate_error(pose0, pose1, cam, img0, img1, dpt0, grad1, params=None):
    params = params or default_params()
    img0, img1, dpt0, grad1 = map(_f32, (img0, img1, dpt0, grad1))
    H, W = img0.shape
    pose0 = np.ascontiguousarray(pose0, dtype=np.float32)
    pose1 = np.ascontiguousarray(pose1, dtype=np.float32)
    c = _cam(cam)
    inl, res = C.c_uint64(0), C.c_float(0)
    lib().dfkr_sfm_evaluate_error_f(_ptr(pose0), _ptr(pose1), C.byref(c), C.c_int(W), C.c_int(H), _ptr(img0), _pitch(img0),
                                    _ptr(img1), _pitch(img1), _ptr(dpt0), _pitch(dpt0), _ptr(grad1), _pitch(grad1),
                                    C.byref(params), C.byref(res), C.byref(inl))
    return float(res.value), int(inl.value)


def se3_run_step(se3, cam, img0, img1, dpt0, grad1, huber_delta=0.1) -> StepResult:
    img0, img1, dpt0, grad1 = map(_f32, (img0, img1, dpt0, grad1))
    H, W = img0.shape
    se3 = np.ascontiguousarray(se3, dtype=np.float32)
    c = _cam(cam)
    inl, res = C.c_uint64(0), C.c_float(0)
    JtJ, Jtr = np.zeros(21, dtype=np.float32), np.zeros(6, dtype=np.float32)
    lib().dfkr_se3_run_step_f(_ptr(se3), C.byref(c), C.c_int(W), C.c_int(H), _ptr(img0), _pitch(img0), _ptr(img1),
                              _pitch(img1), _ptr(dpt0), _pitch(dpt0), _ptr(grad1), _pitch(grad1), C.c_float(huber_delta),
                              _ptr(JtJ), _ptr(Jtr), C.byref(res), C.byref(inl))
    return StepResult(JtJ, Jtr, float(res.value), int(inl.value))


def depth_run_step(code, target_dpt, prx_orig, prx_jac, avg_dpt=2.0) -> StepResult:
    code = np.ascontiguousarray(code, dtype=np.float32)
    tgt, prx_orig, prx_jac = _f32(target_dpt), _f32(prx_orig), _f32(prx_jac)
    H, W = tgt.shape
    Cs = code.shape[0]
    inl, res = C.c_uint64(0), C.c_float(0)
    JtJ, Jtr = np.zeros(Cs * (Cs + 1) // 2, dtype=np.float32), np.zeros(Cs, dtype=np.float32)
    fn = lib().dfkr_depth_run_step_f
    fn.restype = C.c_int
    rc = fn(_ptr(code), C.c_int(Cs), C.c_int(W), C.c_int(H), _ptr(tgt), _pitch(tgt), _ptr(prx_orig), _pitch(prx_orig),
            _ptr(prx_jac), _pitch(prx_jac), C.c_float(avg_dpt), _ptr(JtJ), _ptr(Jtr), C.byref(res), C.byref(inl))
    if rc != 0:
        raise ValueError("code size not instantiated")
    return StepResult(JtJ, Jtr, float(res.value), int(inl.value))


def reprojection_rows(pose0, pose1, code, cam, prx_orig, prx_jac, query_xy, train_xy, cauchy_delta, sigma, avg_dpt=2.0):
    code = np.ascontiguousarray(code, dtype=np.float32)
    prx_orig, prx_jac = _f32(prx_orig), _f32(prx_jac)
    H, W = prx_orig.shape
    Cs = code.shape[0]
    q = np.ascontiguousarray(query_xy, dtype=np.float32).reshape(-1, 2)
    t = np.ascontiguousarray(train_xy, dtype=np.float32).reshape(-1, 2)
    M = q.shape[0]
    pose0 = np.ascontiguousarray(pose0, dtype=np.float32)
    pose1 = np.ascontiguousarray(pose1, dtype=np.float32)
    c = _cam(cam)
    rows = np.zeros((2 * M, 13 + Cs), dtype=np.float32)
    fn = lib().dfkr_reprojection_rows_f
    fn.restype = C.c_float
    tot = fn(_ptr(pose0), _ptr(pose1), _ptr(code), C.c_int(Cs), C.byref(c), C.c_int(W), C.c_int(H), _ptr(prx_orig),
             _pitch(prx_orig), _ptr(prx_jac), _pitch(prx_jac), C.c_int(M), _ptr(q), _ptr(t), C.c_float(cauchy_delta),
             C.c_float(sigma), C.c_float(avg_dpt), _ptr(rows))
    if tot < 0:
        raise ValueError("code size not instantiated")
    return rows, float(tot)


def sparse_geometric_rows(pose0, pose1, code0, code1, cam, prx0_orig, prx0_jac, prx1_orig, prx1_jac, dpt_grad1, points_xy,
                          huber_delta, avg_dpt=2.0):
    code0 = np.ascontiguousarray(code0, dtype=np.float32)
    code1 = np.ascontiguousarray(code1, dtype=np.float32)
    prx0_orig, prx0_jac, prx1_orig, prx1_jac, dpt_grad1 = map(_f32, (prx0_orig, prx0_jac, prx1_orig, prx1_jac, dpt_grad1))
    H, W = prx0_orig.shape
    Cs = code0.shape[0]
    pts = np.ascontiguousarray(points_xy, dtype=np.int32).reshape(-1, 2)
    M = pts.shape[0]
    pose0 = np.ascontiguousarray(pose0, dtype=np.float32)
    pose1 = np.ascontiguousarray(pose1, dtype=np.float32)
    c = _cam(cam)
    rows = np.zeros((M, 13 + 2 * Cs), dtype=np.float32)
    fn = lib().dfkr_sparse_geometric_rows_f
    fn.restype = C.c_int
    n = fn(_ptr(pose0), _ptr(pose1), _ptr(code0), _ptr(code1), C.c_int(Cs), C.byref(c), C.c_int(W), C.c_int(H),
           _ptr(prx0_orig), _pitch(prx0_orig), _ptr(prx0_jac), _pitch(prx0_jac), _ptr(prx1_orig), _pitch(prx1_orig),
           _ptr(prx1_jac), _pitch(prx1_jac), _ptr(dpt_grad1), _pitch(dpt_grad1), C.c_int(M),
           pts.ctypes.data_as(C.POINTER(C.c_int)), C.c_float(huber_delta), C.c_float(avg_dpt), _ptr(rows))
    if n < 0:
        raise ValueError("code size not instantiated")
    return rows, int(n)


def update_depth(code, prx_orig, prx_jac, avg_dpt=2.0):
    code = np.ascontiguousarray(code, dtype=np.float32)
    prx_orig, prx_jac = _f32(prx_orig), _f32(prx_jac)
    H, W = prx_orig.shape
    out = np.zeros((H, W), dtype=np.float32)
    rc = lib().dfkr_update_depth_f(_ptr(code), C.c_int(code.shape[0]), C.c_int(W), C.c_int(H), _ptr(prx_orig),
                                   _pitch(prx_orig), _ptr(prx_jac), _pitch(prx_jac), C.c_float(avg_dpt), _ptr(out), _pitch(out))
    if rc != 0:
        raise ValueError("code size not instantiated")
    return out


def sfm_throughput(pose0, pose1, levels, nthreads, evals_per_thread, params=None):
    """`nthreads` threads x `evals_per_thread` whole-pyramid evaluations through the reference's own host loop
    (x outer / y inner, ut_sfmaligner.cpp:303-315).  Returns (wall seconds, StepResult of level 0 from thread 0)."""
    params = params or default_params()
    arr, keep = make_levels(levels)
    Cs = keep[0][3].shape[2]
    NP = 12 + Cs
    NH = NP * (NP + 1) // 2
    rec = np.zeros(NH + NP + 2, dtype=np.float32)
    pose0 = np.ascontiguousarray(pose0, dtype=np.float32)
    pose1 = np.ascontiguousarray(pose1, dtype=np.float32)
    fn = lib().dfkr_sfm_throughput_f
    fn.restype = C.c_double
    dt = fn(C.c_int(nthreads), C.c_int(evals_per_thread), _ptr(pose0), _ptr(pose1), C.c_int(Cs), C.c_int(len(levels)), arr,
            C.byref(params), _ptr(rec))
    if dt < 0:
        raise ValueError(f"code size {Cs} is not instantiated in oracle/ref_driver.cpp")
    return float(dt), StepResult(rec[:NH].copy(), rec[NH:NH + NP].copy(), float(rec[NH + NP]), int(rec[NH + NP + 1]))
