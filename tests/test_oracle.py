"""CPU tests that pin the oracle (oracle/) against what the reference's own tests pin.

The reference has no golden vectors for this path ("parity unpinned", oracle/dfk_oracle.h); its
tests are relational.  We reproduce those relations on the restatement:
  * tests/ut_warping.cpp:150-212      RelativePose Jacobians == box-minus finite differences (tol 1e-5)
  * tests/ut_warping.cpp:214-308      FindCorrespondenceJacobianPose / ...Prx == finite differences
  * tests/ut_sfmaligner.cpp:329-487   0.5 * d(residual)/eps == Jtr  (poses and code)
  * tests/ut_se3aligner.cpp:173-211   40 Gauss-Newton iterations on data/testimg/1047->1052 converge
                                      to residual/inliers <= 1e-3
  * tests/ut_cuda_utils.cpp:73-144    Sobel == cv::Sobel(scale 1/8) (1e-4), blur-down == GaussianBlur +
                                      pyrDown (1e-1), interior pixels
and the internal consistency the CUDA kernels rely on (reduced (7+C) Gram expands exactly to the
(12+C) system; fp32 and fp64 flavours agree; loop order / OpenMP variants agree).
"""
import numpy as np
import pytest

from deepfactors_b200 import se3, synth
from helpers import scenenet_inputs


# ------------------------------------------------------------------------------------------ helpers
def so3_log(q):
    q = np.asarray(q, dtype=np.float64)
    n = np.linalg.norm(q[:3])
    if n < 1e-12:
        return 2.0 * q[:3] / q[3]
    return 2.0 * np.arctan2(n, q[3]) * q[:3] / n


def boxminus(p_fwd, p):
    """[t_fwd - t, log(R_fwd * R^-1)] (tests/ut_warping.cpp:188-190)"""
    qinv = np.array([-p[0], -p[1], -p[2], p[3]])
    dq = se3.quat_mul(p_fwd[:4], qinv)
    if dq[3] < 0:
        dq = -dq
    return np.concatenate([p_fwd[4:7] - p[4:7], so3_log(dq)])


def random_pose(rng):
    return se3.make_pose(rng.uniform(-0.5, 0.5, 3), rng.uniform(-1, 1, 3), np.float64)


# ------------------------------------------------------------------------------------------ SE3 algebra
def test_relative_pose_jacobians_match_finite_differences(oracle):
    rng = np.random.default_rng(3)
    for _ in range(5):
        a, b = random_pose(rng), random_pose(rng)
        ab, ja, jb = oracle.relative_pose(a, b)
        ref = se3.compose(se3.inverse(a), b, np.float64)
        assert np.allclose(ab[4:], ref[4:], atol=1e-12)
        assert np.allclose(np.abs(ab[:4] @ ref[:4]), 1.0, atol=1e-12)
        eps, tol = 1e-6, 1e-5  # ut_warping.cpp:170-171
        fd_a = np.zeros((6, 6))
        fd_b = np.zeros((6, 6))
        for i in range(6):
            ab_f, _, _ = oracle.relative_pose(oracle.pose_perturb(a, i, eps), b)
            fd_a[:, i] = boxminus(ab_f, ab) / eps
            ab_f, _, _ = oracle.relative_pose(a, oracle.pose_perturb(b, i, eps))
            fd_b[:, i] = boxminus(ab_f, ab) / eps
        assert np.abs(fd_a - ja).max() < tol
        assert np.abs(fd_b - jb).max() < tol


def test_so3_exp_and_perturb_match_numpy_helpers(oracle):
    rng = np.random.default_rng(4)
    for _ in range(10):
        w = rng.uniform(-1, 1, 3)
        assert np.allclose(oracle.so3_exp(w), se3.so3_exp(w), atol=1e-14)
        p = random_pose(rng)
        for i in range(6):
            assert np.allclose(oracle.pose_perturb(p, i, 1e-3), se3.perturb(p, i, 1e-3, np.float64), atol=1e-14)


def test_correspondence_jacobians_match_finite_differences(oracle):
    """ut_warping.cpp:214-308 (pose) and ut_sfmaligner.cpp:137-216 (prx/code) on the probe pixel."""
    rng = np.random.default_rng(5)
    cam = synth.Camera.scenenet(800, 600)  # ut_warping.cpp:57-59
    eps, tol = 1e-6, 5e-4                  # FindiffParams<double> ut_warping.cpp:39-43
    checked = 0
    for _ in range(50):
        pose = se3.make_pose(rng.uniform(-0.1, 0.1, 3), rng.uniform(-0.3, 0.3, 3), np.float64)
        x, y, d = rng.uniform(100, 700), rng.uniform(100, 500), rng.uniform(3, 10)
        r0 = oracle.probe_pixel(x, y, d, cam, pose, border=1)
        if not r0["valid"]:
            continue
        checked += 1
        fd = np.zeros((2, 6))
        for i in range(6):
            rf = oracle.probe_pixel(x, y, d, cam, oracle.pose_perturb(pose, i, eps), border=1)
            fd[:, i] = (rf["pix1"] - r0["pix1"]) / eps
        assert np.abs(fd - r0["J_pose"]).max() < tol * max(1.0, np.abs(fd).max())
        # prx Jacobian: prx = avg/(avg+d)  =>  d(prx) = avg/prx - avg
        avg = 2.0
        prx = avg / (avg + d)
        eps_p = 1e-7
        d_f = avg / (prx + eps_p) - avg
        rf = oracle.probe_pixel(x, y, d_f, cam, pose, border=1, avg_dpt=avg)
        fdp = (rf["pix1"] - r0["pix1"]) / eps_p
        assert np.abs(fdp - r0["J_prx"]).max() < 1e-3 * max(1.0, np.abs(fdp).max())
    assert checked > 20


# ------------------------------------------------------------------------------------------ DenseSfm
@pytest.fixture(scope="module")
def small_pair():
    pair = synth.make_pair(160, 120, 8, 1, seed=7)
    return pair


def _run(oracle, pair, lvl=0, **kw):
    L = pair.levels[lvl]
    valid0 = kw.pop("valid0", None)
    pose0 = kw.pop("pose0", pair.pose0)
    pose1 = kw.pop("pose1", pair.pose1)
    dpt0 = kw.pop("dpt0", L.dpt0)
    return oracle.sfm_run_step(pose0, pose1, L.cam, L.img0, L.img1, dpt0, valid0, L.prx_jac, L.grad1, **kw)


def test_sfm_fp32_matches_fp64_and_inlier_fraction(oracle, small_pair):
    r32 = _run(oracle, small_pair, precision="f32")
    r64 = _run(oracle, small_pair, precision="f64")
    n = small_pair.levels[0].width * small_pair.levels[0].height
    assert r32.inliers == r64.inliers
    assert 0.3 * n < r64.inliers < 0.95 * n  # test poses give a partial overlap
    H32, H64 = r32.dense().astype(np.float64), r64.dense()
    scale = np.abs(H64).max()
    assert np.abs(H32 - H64).max() / scale < 2e-5
    assert np.abs(r32.Jtr - r64.Jtr).max() / np.abs(r64.Jtr).max() < 2e-4
    assert abs(r32.residual - r64.residual) / r64.residual < 1e-5
    # Hessian is symmetric PSD (it is J^T J)
    ev = np.linalg.eigvalsh(H64)
    assert ev.min() > -1e-9 * ev.max()


def test_sfm_loop_order_and_omp_variants_agree(oracle, small_pair):
    a = _run(oracle, small_pair, precision="f64", loop_order=0)
    b = _run(oracle, small_pair, precision="f64", loop_order=1)
    assert a.inliers == b.inliers
    assert np.allclose(a.JtJ, b.JtJ, rtol=1e-10, atol=1e-10 * np.abs(a.JtJ).max())
    s = _run(oracle, small_pair, precision="f32", loop_order=1)
    for nt in (1, 3, oracle.omp_max_threads()):
        o = _run(oracle, small_pair, precision="f32", omp_threads=nt)
        assert o.inliers == s.inliers
        assert np.abs(o.JtJ - s.JtJ).max() / np.abs(s.JtJ).max() < 1e-5


def test_sfm_valid0_is_only_ever_set(oracle, small_pair):
    L = small_pair.levels[0]
    v = np.full((L.height, L.width), 0.25, dtype=np.float32)
    r = _run(oracle, small_pair, valid0=v)
    assert int((v == 1.0).sum()) == r.inliers
    assert set(np.unique(v)) <= {0.25, 1.0}  # never cleared (dense_sfm.h:161)


def test_reduced_system_expands_to_reference_layout(oracle, small_pair):
    """G = sum m^T m with m = w*[a | e*jc | diff] expands with E = [[P0,P1,0],[0,0,I]]
    to exactly the (12+C) system the reference accumulates.  Checked through linear algebra on the oracle's
    own output: the pose0/pose1 blocks must be congruent images of the same 6x6 block."""
    r = _run(oracle, small_pair, precision="f64")
    H = r.dense()
    _, P1, P0 = oracle.relative_pose(small_pair.pose1.astype(np.float64), small_pair.pose0.astype(np.float64))
    # H00 = P0^T Gaa P0, H11 = P1^T Gaa P1, H01 = P0^T Gaa P1
    Gaa = np.linalg.solve(P0.T, np.linalg.solve(P0.T, H[:6, :6].T).T)
    assert np.allclose(P1.T @ Gaa @ P1, H[6:12, 6:12], rtol=1e-9, atol=1e-9 * np.abs(H).max())
    assert np.allclose(P0.T @ Gaa @ P1, H[:6, 6:12], rtol=1e-9, atol=1e-9 * np.abs(H).max())
    Gac = np.linalg.solve(P0.T, H[:6, 12:])
    assert np.allclose(P1.T @ Gac, H[6:12, 12:], rtol=1e-9, atol=1e-9 * np.abs(H).max())
    ga = np.linalg.solve(P0.T, r.Jtr[:6])
    assert np.allclose(P1.T @ ga, r.Jtr[6:12], rtol=1e-9, atol=1e-9 * np.abs(r.Jtr).max())


def test_sfm_jtr_matches_finite_difference_of_residual(oracle):
    """ut_sfmaligner.cpp:329-487: 0.5 * (res(x+eps) - res(x)) / eps ~= Jtr for both poses and the code.
    Done in fp64 on smooth images (periods >= 70 px, so that the Sobel gradient the Jacobian uses is within
    ~1% of the true derivative of the bilinearly sampled image) with the Huber threshold opened up; a
    perturbation that changes the inlier set (a jump of the cost) is skipped, as the analytic Jacobian cannot
    see it."""
    code = np.zeros(8, dtype=np.float32)
    L = synth.make_level(160, 120, 8, scale=0.1, seed=11, code=code)
    pose0, pose1 = synth.reference_test_poses(np.float64)
    prm = oracle.default_params(huber_delta=1e6)
    b64 = np_dense_sfm(oracle, pose0, pose1, L, prm)
    eps = 1e-6
    scale = np.abs(b64.Jtr[:12]).max()
    checked = 0
    for which in (0, 1):
        for i in range(6):
            p0, p1 = pose0, pose1
            if which == 0:
                p0 = oracle.pose_perturb(p0, i, eps)
            else:
                p1 = oracle.pose_perturb(p1, i, eps)
            r = np_dense_sfm(oracle, p0, p1, L, prm)
            if r.inliers != b64.inliers:
                continue
            checked += 1
            fd = 0.5 * (r.residual - b64.residual) / eps
            assert abs(fd - b64.Jtr[6 * which + i]) < 2e-2 * scale
    assert checked >= 8
    # code: perturb code -> new depth -> residual (ut_sfmaligner.cpp:458-484)
    code_eps = 1e-3  # the reference's own step (ut_sfmaligner.cpp:459)
    cscale = np.abs(b64.Jtr[12:]).max()
    checked = 0
    for k in range(8):
        c = np.zeros(8)
        c[k] += code_eps
        prx = L.prx_orig.astype(np.float64) + L.prx_jac.astype(np.float64) @ c
        dpt = (2.0 / prx - 2.0)
        r = np_dense_sfm(oracle, pose0, pose1, L, prm, dpt0=dpt)
        if r.inliers != b64.inliers:
            continue
        checked += 1
        fd = 0.5 * (r.residual - b64.residual) / code_eps
        assert abs(fd - b64.Jtr[12 + k]) < 3e-2 * cscale
    assert checked >= 5


def np_dense_sfm(oracle, p0, p1, L, prm, dpt0=None):
    p10, P1, P0 = oracle.relative_pose(p1, p0)
    R = se3.quat_to_matrix(p10[:4])
    t = p10[4:]
    cam = L.cam
    H_, W_ = L.height, L.width
    y, x = np.mgrid[0:H_, 0:W_].astype(np.float64)
    d = (L.dpt0 if dpt0 is None else dpt0).astype(np.float64)
    xn, yn = (x - cam.u0) / cam.fx, (y - cam.v0) / cam.fy
    X = np.stack([xn * d, yn * d, d], -1)
    p = X @ R.T
    T = p + t
    Z = T[..., 2]
    with np.errstate(divide="ignore", invalid="ignore"):
        u = cam.fx * T[..., 0] / Z + cam.u0
        v = cam.fy * T[..., 1] / Z + cam.v0
    b = prm.valid_border
    valid = (Z > prm.min_dpt) & (u >= b) & (u < cam.width - b) & (v >= b) & (v < cam.height - b)
    u = np.where(valid, u, 2.0)
    v = np.where(valid, v, 2.0)
    ix, iy = np.floor(u).astype(int), np.floor(v).astype(int)
    fu, fv = u - ix, v - iy

    def bil(img):
        img = img.astype(np.float64)
        q00, q01, q10, q11 = img[iy, ix], img[iy, ix + 1], img[iy + 1, ix], img[iy + 1, ix + 1]
        top = q00 + fu * (q01 - q00)
        bot = q10 + fu * (q11 - q10)
        return top + fv * (bot - top)

    gx, gy, I1 = bil(L.grad1[..., 0]), bil(L.grad1[..., 1]), bil(L.img1)
    Zs = np.where(valid, Z, 1.0)
    c00, c11 = cam.fx / Zs, cam.fy / Zs
    c02, c12 = -cam.fx * T[..., 0] / Zs / Zs, -cam.fy * T[..., 1] / Zs / Zs
    px, py, pz = p[..., 0], p[..., 1], p[..., 2]
    A0 = np.stack([c00, 0 * c00, c02, c02 * py, c00 * pz - c02 * px, -c00 * py], -1)
    A1 = np.stack([0 * c11, c11, c12, c12 * py - c11 * pz, -c12 * px, c11 * px], -1)
    a = -(gx[..., None] * A0 + gy[..., None] * A1)
    ray = np.stack([xn, yn, np.ones_like(xn)], -1) @ R.T
    pJx = c00 * ray[..., 0] + c02 * ray[..., 2]
    pJy = c11 * ray[..., 1] + c12 * ray[..., 2]
    prx = prm.avg_dpt / (prm.avg_dpt + d)
    dJ = -prm.avg_dpt / (prx * prx)
    e = -(gx * pJx + gy * pJy) * dJ
    diff = L.img0.astype(np.float64) - I1
    aa = np.abs(diff)
    with np.errstate(divide="ignore", invalid="ignore"):
        w = np.where(aa <= prm.huber_delta, 1.0, np.sqrt(prm.huber_delta * (2 * aa - prm.huber_delta)) / aa)
    J = np.concatenate([a @ P0, a @ P1, e[..., None] * L.prx_jac.astype(np.float64)], -1) * w[..., None]
    r = diff * w
    J = J[valid]
    r = r[valid]
    res = type("R", (), {})()
    res.residual = float(r @ r)
    res.Jtr = J.T @ r
    res.H = J.T @ J
    res.inliers = int(valid.sum())
    return res


def test_numpy_restatement_matches_c_oracle(oracle, small_pair):
    """two independent restatements (vectorised numpy, scalar C) of the same reference math agree"""
    L = small_pair.levels[0]
    prm = oracle.default_params()
    r_np = np_dense_sfm(oracle, small_pair.pose0.astype(np.float64), small_pair.pose1.astype(np.float64), L, prm)
    r_c = _run(oracle, small_pair, precision="f64")
    assert r_np.inliers == r_c.inliers
    assert np.allclose(r_np.H, r_c.dense(), rtol=1e-9, atol=1e-9 * np.abs(r_np.H).max())
    assert np.allclose(r_np.Jtr, r_c.Jtr, rtol=1e-9, atol=1e-9 * np.abs(r_np.Jtr).max())
    assert abs(r_np.residual - r_c.residual) < 1e-9 * r_c.residual


def test_evaluate_error_uses_border_one_and_matches_step_residual_semantics(oracle, small_pair):
    L = small_pair.levels[0]
    res, inl = oracle.sfm_evaluate_error(small_pair.pose0, small_pair.pose1, L.cam, L.img0, L.img1, L.dpt0,
                                         precision="f64")
    step = _run(oracle, small_pair, precision="f64")
    # border 1 (EvaluateError, dense_sfm.h:91) admits at least the pixels of border 2 (RunStep, :154-155)
    assert inl >= step.inliers
    assert inl - step.inliers < 0.05 * step.inliers
    assert res >= step.residual * 0.999


# ------------------------------------------------------------------------------------------ SE3 KAT
def test_se3_image_alignment_converges_like_reference(oracle, golden):
    """tests/ut_se3aligner.cpp:173-211 ImageAlignmentTest: 40 GN iterations from identity, huber 0.1,
    error = residual / inliers <= 1e-3."""
    cam, img0, img1, dpt0 = scenenet_inputs(golden)
    grad1 = oracle.sobel_gradients(img1)
    pose = se3.identity(np.float64)
    err = None
    first = None
    for _ in range(40):
        r = oracle.se3_run_step(pose.astype(np.float32), cam, img0, img1, dpt0, grad1, 0.1, precision="f32")
        assert r.inliers > 0
        pose = se3.se3_solve_and_update(r.dense(), r.Jtr, pose)
        err = r.residual / r.inliers
        first = first if first is not None else err
    assert err <= 1e-3
    assert err < first


def test_se3_fp32_matches_fp64_and_warp_counts(oracle, golden):
    cam, img0, img1, dpt0 = scenenet_inputs(golden)
    grad1 = oracle.sobel_gradients(img1)
    pose = se3.make_pose([0.01, -0.02, 0.005], [0.02, 0.01, -0.01])
    a = oracle.se3_run_step(pose, cam, img0, img1, dpt0, grad1, 0.1, precision="f32")
    b = oracle.se3_run_step(pose, cam, img0, img1, dpt0, grad1, 0.1, precision="f64")
    assert a.inliers == b.inliers
    assert np.abs(a.JtJ - b.JtJ).max() / np.abs(b.JtJ).max() < 2e-5
    img2, res, inl = oracle.se3_warp(pose, cam, img0, img1, dpt0)
    assert inl == a.inliers  # both use depth > 0 and border 1
    assert (img2 != 0).sum() <= inl


# ------------------------------------------------------------------------------------------ image proc
def test_sobel_matches_opencv_like_reference(oracle, golden):
    img = golden["gray_1047"].astype(np.float32) * np.float32(1 / 255.0)
    g = oracle.sobel_gradients(img)
    eps = 1e-4  # ut_cuda_utils.cpp:69
    assert np.abs(g[1:-1, 1:-1, 0] - golden["ocv_sobel_x_1047"][1:-1, 1:-1]).max() < eps
    assert np.abs(g[1:-1, 1:-1, 1] - golden["ocv_sobel_y_1047"][1:-1, 1:-1]).max() < eps
    assert np.array_equal(g, synth.sobel_np(img)) or np.abs(g - synth.sobel_np(img)).max() < 1e-7


def test_blur_down_matches_opencv_like_reference(oracle, golden):
    img = golden["gray_1047"].astype(np.float32) * np.float32(1 / 255.0)
    out = oracle.gaussian_blur_down(img)
    assert out.shape == (120, 160)
    eps = 1e-1  # ut_cuda_utils.cpp:68
    assert np.abs(out[1:-1, 1:-1] - golden["ocv_blurdown_1047"][1:-1, 1:-1]).max() < eps
    # a constant image stays constant (the kernel is normalised by its running sum)
    c = oracle.gaussian_blur_down(np.full((50, 70), 0.37, dtype=np.float32))
    assert np.abs(c - 0.37).max() < 1e-6


def test_update_depth_and_squared_error(oracle):
    rng = np.random.default_rng(0)
    prx = (0.3 + 0.4 * rng.random((30, 40))).astype(np.float32)
    jac = (rng.standard_normal((30, 40, 8)) * 0.02).astype(np.float32)
    code = rng.standard_normal(8).astype(np.float32)
    d = oracle.update_depth(code, prx, jac, 2.0)
    ref = 2.0 / (prx.astype(np.float64) + jac.astype(np.float64) @ code.astype(np.float64)) - 2.0
    assert np.abs(d - ref).max() < 1e-5
    a = rng.random((30, 40)).astype(np.float32)
    b = rng.random((30, 40)).astype(np.float32)
    ref = float(((a.astype(np.float64) - b) ** 2).sum())
    assert abs(oracle.squared_error(a, b) - ref) < 1e-4 * ref
    assert abs(oracle.squared_error(a, b, "f64") - ref) < 1e-9 * ref


def test_tracker_loop_converges_coarse_to_fine(oracle, golden):
    """CameraTracker::TrackFrame restated (camera_tracker.cpp:42-69) on the reference's own test pair: the 3-level
    10/5/4 schedule of the default TrackerConfig reaches the same minimum as 40 single-level iterations
    (ut_se3aligner.cpp:173-211 bar: mean residual <= 1e-3)"""
    from helpers import tracking_pyramid
    cams, p0, p1, pd, pg = tracking_pyramid(golden, oracle, 3)
    pose3, frac3, err3, hist = oracle.se3_track(se3.identity(np.float64), cams, p0, p1, pd, pg, (10, 5, 4), 0.1)
    pose1, frac1, err1, _ = oracle.se3_track(se3.identity(np.float64), cams[:1], p0[:1], p1[:1], pd[:1], pg[:1], (40,), 0.1)
    assert len(hist) == 19 and err3 <= 1e-3 and err1 <= 1e-3
    assert np.abs(pose3 - pose1).max() <= 1e-4 and abs(frac3 - frac1) <= 1e-3
    # the coarse levels do the work: the error at the first fine-level iteration is already close to the final one
    first_fine = hist[15][0]
    assert first_fine.residual / first_fine.inliers <= 2.0 * err3
