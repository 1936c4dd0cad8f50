"""GPU-driven Gauss-Newton / Levenberg-Marquardt over a keyframe window.

The reference runs this loop on the host, one factor at a time: ISAM2::update (sources/core/mapping/mapper.cpp:518-519)
calls PhotometricFactor::linearize for every relinearised factor (sources/core/gtsam/photometric_factor.cpp:84-181),
which consults its linearisation cache (:296-328, keyed on pose0 / pose1 / code0 within 1e-6), re-decodes the keyframe's
depth from its code (UpdateDepthMaps, :229,331-341), runs SfmAligner::RunStep synchronously (:267-274), rescales the
residual (:275-282) and slices the 44x44 system into HessianFactor blocks (:126-161); the solver adds the factors up and
solves.  Here every heavy step of one iteration is ONE device operation over the whole window:

    linearise   one batched RunStep launch over all (pair, level) factors whose variables moved, with the depth decode
                fused in (the code of keyframe k0 rides in the work item)               dfk_sfm_run_step_batch
    assemble    block-sparse normal equations of the window, on the device               dfk_window_assemble
                (+ one all-reduce across ranks when the pairs are sharded)
    solve       damped dense solve on the device (Cholesky, float64)                     torch.linalg
    retract     pose: t += dt, R = exp(w) R (gtsam_traits.h:48-58); code += dc           (tiny, host)
    accept      Levenberg-Marquardt on the rescaled residual energy f (what the factor's error() returns)

The host side (cache, retraction, damping schedule) is plain Python / numpy; nothing here needs GTSAM.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

from . import se3
from .factors import WindowBlocks


@dataclass
class LMParams:
    iterations: int = 10
    lambda_init: float = 1e-4
    lambda_up: float = 10.0
    lambda_down: float = 0.1
    lambda_max: float = 1e6
    fix_first_pose: bool = True       # gauge: the window's first keyframe keeps its pose (the reference adds a pose prior)
    code_prior_weight: float = 0.0    # zero-code prior: adds w * I to every code block and -w * code to its gradient
    cache_eps: float = 1e-6           # photometric_factor.cpp:302-316


@dataclass
class LMTrace:
    energy: List[float] = field(default_factory=list)       # accepted energies, energy[0] = initial
    lam: List[float] = field(default_factory=list)
    accepted: List[bool] = field(default_factory=list)
    factors_relinearised: List[int] = field(default_factory=list)  # per linearisation: pairs that were re-evaluated


def damped_solve(H, g, lam: float, fixed: Sequence[int] = ()):
    """(H + lam * diag(H)) dx = g with the variables in `fixed` held (rows / columns removed).  H, g: torch (any device,
    float64) or numpy.  Returns dx of full dimension (zeros at the fixed variables)."""
    if hasattr(H, "detach"):
        import torch
        n = H.shape[0]
        keep = torch.ones(n, dtype=torch.bool, device=H.device)
        if len(fixed):
            keep[torch.as_tensor(list(fixed), device=H.device)] = False
        idx = torch.nonzero(keep).squeeze(1)
        Hk = H.index_select(0, idx).index_select(1, idx)
        d = torch.diagonal(Hk).clone()
        Hk = Hk + torch.diag(lam * d + 1e-12 * d.abs().max())
        gk = g.index_select(0, idx)
        L, info = torch.linalg.cholesky_ex(Hk)
        if int(info.item()) != 0:  # not positive definite at this damping: least squares keeps the loop alive
            sol = torch.linalg.lstsq(Hk, gk.unsqueeze(1)).solution.squeeze(1)
        else:
            sol = torch.cholesky_solve(gk.unsqueeze(1), L).squeeze(1)
        dx = torch.zeros(n, dtype=H.dtype, device=H.device)
        dx[idx] = sol
        return dx
    Hn = np.asarray(H, dtype=np.float64)
    gn = np.asarray(g, dtype=np.float64)
    n = Hn.shape[0]
    keep = np.ones(n, dtype=bool)
    keep[list(fixed)] = False
    Hk = Hn[np.ix_(keep, keep)]
    d = np.diag(Hk).copy()
    Hk = Hk + np.diag(lam * d + 1e-12 * np.abs(d).max())
    dx = np.zeros(n)
    dx[keep] = np.linalg.lstsq(Hk, gn[keep], rcond=None)[0]
    return dx


class LinearisationCache:
    """photometric_factor.cpp:296-328 GetJacobiansIfNeeded for a whole window: a pair's factors are re-evaluated only when
    pose0, pose1 or code0 moved by more than eps since the evaluation whose records are still in the record buffer."""

    def __init__(self, pairs: Sequence[Tuple[int, int]], eps: float = 1e-6):
        self.pairs = list(pairs)
        self.eps = eps
        self._at = [None] * len(self.pairs)  # (pose0, pose1, code0) the stored records were evaluated at

    def stale(self, poses: np.ndarray, codes: np.ndarray) -> List[int]:
        out = []
        for p, (k0, k1) in enumerate(self.pairs):
            at = self._at[p]
            if at is None or np.abs(at[0] - poses[k0]).max() > self.eps or np.abs(at[1] - poses[k1]).max() > self.eps or \
                    np.abs(at[2] - codes[k0]).max() > self.eps:
                out.append(p)
        return out

    def store(self, done: Sequence[int], poses: np.ndarray, codes: np.ndarray):
        for p in done:
            k0, k1 = self.pairs[p]
            self._at[p] = (poses[k0].copy(), poses[k1].copy(), codes[k0].copy())

    def invalidate(self):
        self._at = [None] * len(self.pairs)


def apply_update(poses: np.ndarray, codes: np.ndarray, dx: np.ndarray, code_size: int):
    """variables [pose_k (6: t, w) | code_k (C)] per keyframe; pose retraction of gtsam_traits.h:48-58"""
    B = 6 + code_size
    new_p, new_c = poses.copy(), codes.copy()
    for k in range(poses.shape[0]):
        d = dx[k * B:(k + 1) * B]
        new_p[k] = se3.retract(poses[k].astype(np.float64), d[:6], np.float64)
        new_c[k] = codes[k] + d[6:]
    return new_p, new_c


class WindowOptimizer:
    """Levenberg-Marquardt over the poses and codes of a keyframe window.

    `linearise(poses, codes, pairs_to_eval) -> (window_buffer, f)` is the device pipeline (see SfmWindowProblem below for
    the one built on SfmAligner / Window); injected so that the host logic is testable without a GPU."""

    def __init__(self, layout: WindowBlocks, linearise: Callable, params: Optional[LMParams] = None):
        self.layout = layout
        self.linearise = linearise
        self.params = params or LMParams()
        self.cache = LinearisationCache(layout.pairs, self.params.cache_eps)

    def _system(self, buf, codes):
        H, g, f, inl = self.layout.to_dense(buf)
        w = self.params.code_prior_weight
        if w > 0:
            B = self.layout.B
            for k in range(self.layout.num_keyframes):
                sl = slice(k * B + 6, (k + 1) * B)
                if hasattr(H, "detach"):
                    import torch
                    H[sl, sl] += w * torch.eye(B - 6, dtype=H.dtype, device=H.device)
                    g[sl] -= w * torch.as_tensor(codes[k], dtype=g.dtype, device=g.device)
                else:
                    H[sl, sl] += w * np.eye(B - 6)
                    g[sl] -= w * codes[k]
            f += 0.5 * w * float((codes ** 2).sum())
        return H, g, f

    def _evaluate(self, poses, codes, trace: LMTrace):
        todo = self.cache.stale(poses, codes)
        buf, _ = self.linearise(poses, codes, todo)
        self.cache.store(todo, poses, codes)
        trace.factors_relinearised.append(len(todo))
        return buf

    def run(self, poses, codes) -> Tuple[np.ndarray, np.ndarray, LMTrace]:
        prm = self.params
        poses = np.asarray(poses, dtype=np.float64).copy()
        codes = np.asarray(codes, dtype=np.float64).copy()
        trace = LMTrace()
        fixed = list(range(6)) if prm.fix_first_pose else []
        lam = prm.lambda_init
        buf = self._evaluate(poses, codes, trace)
        H, g, f = self._system(buf, codes)
        trace.energy.append(f)
        for _ in range(prm.iterations):
            dx = damped_solve(H, g, lam, fixed)
            dxh = dx.detach().cpu().numpy() if hasattr(dx, "detach") else np.asarray(dx)
            cand_p, cand_c = apply_update(poses, codes, dxh, self.layout.code_size)
            cbuf = self._evaluate(cand_p, cand_c, trace)
            cH, cg, cf = self._system(cbuf, cand_c)
            ok = np.isfinite(cf) and cf < f
            trace.accepted.append(bool(ok))
            trace.lam.append(lam)
            if ok:
                poses, codes, H, g, f = cand_p, cand_c, cH, cg, cf
                trace.energy.append(f)
                lam = max(lam * prm.lambda_down, 1e-12)
            else:
                # H, g, f of the accepted point are still at hand; the record buffer (and with it the cache) now describes
                # the rejected candidate, which the next candidate is compared against -- nothing to re-evaluate
                lam = lam * prm.lambda_up
                if lam > prm.lambda_max:
                    break
        return poses, codes, trace


class SfmWindowProblem:
    """The device pipeline of one linearisation, on SfmAligner + Window: keyframes hold their pyramids on the device
    (img, grad, prx_orig, prx_jac per level + the dpt / valid buffers the fused decode writes); `linearise` re-evaluates
    the factors of the given pairs in one launch (depth decode fused in) and re-assembles the window."""

    def __init__(self, aligner, cams, keyframes, pairs, allreduce: Optional[Callable] = None):
        import torch
        from . import _lib
        from .aligners import Window
        self.al = aligner
        self.cams = list(cams)
        self.kf = keyframes          # kf[k][l] = dict(img, grad, prx_orig, prx_jac, dpt, valid) of device tensors
        self.pairs = [tuple(p) for p in pairs]
        self.levels = len(self.cams)
        item_pair, sizes = [], []
        for p in range(len(self.pairs)):
            for l in range(self.levels):
                item_pair.append(p)
                t = self.kf[self.pairs[p][0]][l]["img"]
                sizes.append((int(t.shape[1]), int(t.shape[0])))
        self.window = Window(aligner, len(keyframes), self.pairs, item_pair, sizes)
        self.layout = self.window.layout
        self.records = torch.zeros((len(item_pair), _lib.record_floats(aligner.CS)), dtype=torch.float32, device=self.kf[0][0]["img"].device)
        self.allreduce = allreduce

    def _items(self, poses, codes, todo):
        items = []
        for p in todo:
            k0, k1 = self.pairs[p]
            for l in range(self.levels):
                a, b = self.kf[k0][l], self.kf[k1][l]
                items.append(dict(pose0=poses[k0].astype(np.float32), pose1=poses[k1].astype(np.float32), cam=self.cams[l],
                                  img0=a["img"], img1=b["img"], dpt0=a["dpt"], valid0=a["valid"], prx0_jac=a["prx_jac"],
                                  grad1=b["grad"], prx_orig=a["prx_orig"], code=codes[k0].astype(np.float32)))
        return items

    def linearise(self, poses, codes, todo):
        import torch
        if todo:
            work = self.al.make_work_items(self._items(poses, codes, todo))
            if len(todo) == len(self.pairs):
                self.al.RunStepBatch(work, self.records)
            else:
                part = self.al.RunStepBatch(work)
                rows = torch.as_tensor([p * self.levels + l for p in todo for l in range(self.levels)],
                                       device=self.records.device)
                self.records.index_copy_(0, rows, part)
        buf = self.window.assemble(self.records)
        if self.allreduce is not None:
            self.allreduce(buf)
        return buf, None
