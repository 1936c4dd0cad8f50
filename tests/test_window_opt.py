"""Host logic of the GPU-driven Gauss-Newton / LM window loop (deepfactors_b200/window_opt.py) without a GPU: the
linearisation cache (photometric_factor.cpp:296-328), the retraction (gtsam_traits.h:48-58), the damped solve, and the LM
schedule, driven by a synthetic quadratic `linearise`."""
import numpy as np

from deepfactors_b200 import se3
from deepfactors_b200.factors import WindowBlocks
from deepfactors_b200.window_opt import LMParams, LinearisationCache, WindowOptimizer, apply_update, damped_solve


def test_linearisation_cache_follows_the_reference_rule():
    pairs = [(0, 1), (1, 2), (2, 0)]
    c = LinearisationCache(pairs, 1e-6)
    poses = np.tile(se3.identity(np.float64), (3, 1))
    codes = np.zeros((3, 4))
    assert c.stale(poses, codes) == [0, 1, 2]          # nothing evaluated yet
    c.store([0, 1, 2], poses, codes)
    assert c.stale(poses, codes) == []
    codes2 = codes.copy(); codes2[1, 2] = 5e-7          # below eps: still cached (photometric_factor.cpp:302-316)
    assert c.stale(poses, codes2) == []
    codes2[1, 2] = 1e-3                                  # code of keyframe 1 moved: only the pair whose k0 is 1
    assert c.stale(poses, codes2) == [1]
    poses2 = poses.copy(); poses2[2, 4] += 1e-3          # pose of keyframe 2: pairs (1,2) as pose1 and (2,0) as pose0
    assert c.stale(poses2, codes) == [1, 2]


def test_apply_update_uses_the_reference_retraction():
    poses = np.stack([se3.make_pose([0.1, 0.0, -0.2], [1, 2, 3], np.float64), se3.identity(np.float64)])
    codes = np.zeros((2, 3))
    dx = np.array([0.1, -0.2, 0.3, 0.01, 0.02, -0.03, 1, 2, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0], dtype=np.float64)
    p2, c2 = apply_update(poses, codes, dx, 3)
    assert np.allclose(p2[0], se3.retract(poses[0], dx[:6], np.float64)) and np.allclose(c2[0], [1, 2, 3])
    assert np.allclose(p2[1], poses[1]) and np.allclose(c2[1], 0)


def test_damped_solve_numpy_and_torch_agree_and_hold_fixed_variables():
    import torch
    rng = np.random.default_rng(0)
    A = rng.standard_normal((30, 12))
    H = A.T @ A
    g = rng.standard_normal(12)
    dx = damped_solve(H, g, 1e-3, fixed=range(6))
    assert np.all(dx[:6] == 0)
    Hk = H[6:, 6:] + np.diag(1e-3 * np.diag(H[6:, 6:]))
    assert np.allclose(dx[6:], np.linalg.solve(Hk, g[6:]), rtol=1e-6, atol=1e-9)
    dxt = damped_solve(torch.from_numpy(H), torch.from_numpy(g), 1e-3, fixed=range(6)).numpy()
    assert np.allclose(dxt, dx, rtol=1e-6, atol=1e-9)


def test_lm_loop_converges_on_a_quadratic_and_relinearises_only_what_moved():
    """linearise() of a synthetic least-squares problem in the window's own block-sparse layout: two keyframes, one pair;
    energy = |J x - r|^2 in the 2 * (6 + C) variables around the start point."""
    cs = 2
    pairs = [(0, 1)]
    wb = WindowBlocks(2, cs, pairs)
    rng = np.random.default_rng(1)
    # target: keyframe 1 should move by `goal` in translation, code of keyframe 0 to `cgoal`
    goal = np.array([0.05, -0.02, 0.03])
    cgoal = np.array([0.3, -0.1])
    evals = []

    def linearise(poses, codes, todo):
        evals.append(list(todo))
        # residual rows: t1 - goal (3), code0 - cgoal (2); Jacobian identity on those variables
        NP = 12 + cs
        J = np.zeros((5, NP))
        J[0:3, 6:9] = np.eye(3)        # pose1 translation
        J[3:5, 12:14] = np.eye(2)      # code0
        r = np.concatenate([poses[1][4:7] - goal, codes[0] - cgoal])
        H = J.T @ J
        buf = wb.pack([0], H[None], (J.T @ r)[None], [float(r @ r)], [5], [(1, 5)])  # W*H/inliers == 1
        return buf, None

    poses = np.tile(se3.identity(np.float64), (2, 1))
    codes = np.zeros((2, cs))
    opt = WindowOptimizer(wb, linearise, LMParams(iterations=8, lambda_init=1e-6))
    p, c, tr = opt.run(poses, codes)
    assert tr.energy[0] > 1e-3 and tr.energy[-1] < 1e-10 and all(np.diff(tr.energy) < 0)
    assert np.allclose(p[1][4:7], goal, atol=1e-5) and np.allclose(c[0], cgoal, atol=1e-5)
    assert np.allclose(p[0], se3.identity(np.float64))               # the gauge keyframe did not move
    assert evals[0] == [0] and tr.factors_relinearised[0] == 1
    # once converged (dx ~ 0) the candidate equals the accepted point within eps: the cache answers, nothing is re-evaluated
    assert tr.factors_relinearised[-1] == 0
