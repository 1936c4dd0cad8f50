"""The consumer side of the hot path: what the factor graph does with an aligner result.

Restates (host logic only, no GTSAM here) the parts of the reference that sit immediately above the
aligners, so that a sharded evaluation can be reduced into one set of normal equations:

  * `photometric_factor_blocks`  PhotometricFactor::linearize + RunAlignmentStep
        (sources/core/gtsam/photometric_factor.cpp:84-181, 223-293): residual rescale
        res/inliers*W*H (:275-282), JtJ to double, Jtr negated (:105-106), slicing into the
        HessianFactor blocks G11 G12 G13 G22 G23 G33 / g1 g2 g3 (:126-161).
  * `WindowLayout` / `assemble_window`   the block-sparse -> dense normal equations of a keyframe window
        (SURVEY section 8e): variables [pose_k (6) | code_k (C)] per keyframe; a pair (k0 -> k1) adds
        its pose0/code0 blocks to keyframe k0's diagonal block, pose1 to k1's and the pose0-pose1 /
        pose1-code0 couplings off the diagonal.
  * `shard_pairs` / `allreduce_window`   pairs shard across ranks with no data-path collective; ONE
        all-reduce (sum) of the window's normal equations per Gauss-Newton step joins them
        (torch.distributed: NCCL over NVLink on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Sequence, Tuple

import numpy as np


def record_layout(code_size: int) -> Tuple[int, int, int]:
    """(NP, NH, record_floats) of a device result record [JtJ packed | Jtr | residual | inliers bits]."""
    n = 12 + code_size
    nh = n * (n + 1) // 2
    return n, nh, nh + n + 2


def unpack_records(records, code_size: int):
    """records: [n, REC] float32 (numpy or torch, host or device) -> dense JtJ [n,NP,NP], Jtr [n,NP], residual [n],
    inliers [n] (int64), computed with the array library the input comes from."""
    n, nh, rec = record_layout(code_size)
    if hasattr(records, "detach"):  # torch
        import torch
        r = records
        iu = torch.triu_indices(n, n, device=r.device)
        H = torch.zeros((r.shape[0], n, n), dtype=r.dtype, device=r.device)
        H[:, iu[0], iu[1]] = r[:, :nh]
        H = H + torch.triu(H, 1).transpose(1, 2)
        inl = r[:, nh + n + 1].contiguous().view(torch.int32).to(torch.int64)
        return H, r[:, nh:nh + n], r[:, nh + n], inl
    r = np.asarray(records, dtype=np.float32)
    iu = np.triu_indices(n)
    H = np.zeros((r.shape[0], n, n), dtype=r.dtype)
    H[:, iu[0], iu[1]] = r[:, :nh]
    H = H + np.transpose(np.triu(H, 1), (0, 2, 1))
    inl = np.ascontiguousarray(r[:, nh + n + 1]).view(np.uint32).astype(np.int64)
    return H, r[:, nh:nh + n], r[:, nh + n], inl


def photometric_factor_blocks(JtJ_dense, Jtr, residual, inliers, width, height, code_size):
    """photometric_factor.cpp:84-181,275-282 for one aligner result.
    Returns (Gs, gs, f): Gs = [G11, G12, G13, G22, G23, G33], gs = [g1, g2, g3] (= -Jtr blocks), f = rescaled
    residual (inf when there is no overlap, :279-282).  float64 like the reference's cast."""
    H = np.asarray(JtJ_dense, dtype=np.float64)
    g = -np.asarray(Jtr, dtype=np.float64)
    c = code_size
    Gs = [H[0:6, 0:6], H[0:6, 6:12], H[0:6, 12:12 + c], H[6:12, 6:12], H[6:12, 12:12 + c], H[12:12 + c, 12:12 + c]]
    gs = [g[0:6], g[6:12], g[12:12 + c]]
    f = float(residual) / float(inliers) * float(width) * float(height) if inliers > 0 else float("inf")
    return Gs, gs, f


@dataclass
class WindowLayout:
    """Variable order of a keyframe window: keyframe k owns [pose (6) | code (C)] at offset k * (6 + C)."""
    num_keyframes: int
    code_size: int

    @property
    def block(self) -> int:
        return 6 + self.code_size

    @property
    def dim(self) -> int:
        return self.num_keyframes * self.block


def assemble_window(layout: WindowLayout, pairs: Sequence[Tuple[int, int]], JtJ, Jtr, residual, inliers, sizes):
    """Scatter-add the per-(pair, level) systems into the window's dense normal equations.

    pairs[i] = (k0, k1): keyframe k0 is warped into frame k1 (pose0/code0 belong to k0, pose1 to k1).
    JtJ [n, NP, NP], Jtr [n, NP] in the aligner's column order [pose0 | pose1 | code0].  sizes[i] = (W, H) of the
    level (for the residual rescale).  Works on numpy arrays or torch tensors (any device).
    Returns (H [dim, dim], g [dim], f) with g = -sum Jtr (photometric_factor.cpp:106) and f = sum of rescaled residuals
    over items with overlap."""
    c, b = layout.code_size, layout.block
    is_torch = hasattr(JtJ, "detach")
    if is_torch:
        import torch
        H = torch.zeros((layout.dim, layout.dim), dtype=torch.float64, device=JtJ.device)
        g = torch.zeros((layout.dim,), dtype=torch.float64, device=JtJ.device)
        J64, r64 = JtJ.to(torch.float64), Jtr.to(torch.float64)
    else:
        H = np.zeros((layout.dim, layout.dim))
        g = np.zeros(layout.dim)
        J64, r64 = np.asarray(JtJ, dtype=np.float64), np.asarray(Jtr, dtype=np.float64)
    f = 0.0
    for i, (k0, k1) in enumerate(pairs):
        p0, c0, p1 = k0 * b, k0 * b + 6, k1 * b
        # local column ranges: pose0 [0,6), pose1 [6,12), code0 [12,12+c)
        loc = [(slice(0, 6), slice(p0, p0 + 6)), (slice(6, 12), slice(p1, p1 + 6)), (slice(12, 12 + c), slice(c0, c0 + c))]
        for la, ga in loc:
            g[ga] -= r64[i, la]
            for lb, gb in loc:
                H[ga, gb] += J64[i, la, lb]
        inl = int(inliers[i])
        if inl > 0:
            f += float(residual[i]) / inl * sizes[i][0] * sizes[i][1]
    return H, g, f


def shard_pairs(num_pairs: int, world_size: int, rank: int) -> range:
    """Contiguous, balanced shard of the pair list for `rank` (sizes differ by at most one)."""
    lo = (num_pairs * rank) // world_size
    hi = (num_pairs * (rank + 1)) // world_size
    return range(lo, hi)


def allreduce_window(H, g, group=None):
    """The one collective of a sharded Gauss-Newton step: sum the window's normal equations over ranks.
    H, g are torch tensors (CUDA with NCCL, CPU with gloo); reduced in place, also returned."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(H, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(g, op=dist.ReduceOp.SUM, group=group)
    return H, g


def gauss_newton_step(H, g, damping: float = 0.0):
    """Solve H dx = g (g already carries the sign flip) with optional Levenberg damping; numpy, float64."""
    Hn = np.asarray(H, dtype=np.float64)
    gn = np.asarray(g, dtype=np.float64)
    if damping > 0:
        Hn = Hn + damping * np.diag(np.diag(Hn))
    return np.linalg.lstsq(Hn, gn, rcond=None)[0]
