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


@dataclass
class WindowBlocks:
    """The packed block-sparse buffer dfk_window_assemble writes (include/dfk.h): K diagonal blocks B x B,
    K gradients B, P coupling blocks B x 6 ([pose0 | code0] of k0 x pose1 of k1), then f and the inlier total."""
    num_keyframes: int
    code_size: int
    pairs: Sequence[Tuple[int, int]]

    @property
    def B(self) -> int:
        return 6 + self.code_size

    @property
    def floats(self) -> int:
        K, P, B = self.num_keyframes, len(self.pairs), self.B
        return K * (B * B + B) + P * 6 * B + 2

    def offsets(self):
        K, P, B = self.num_keyframes, len(self.pairs), self.B
        o_g = K * B * B
        o_c = o_g + K * B
        o_t = o_c + P * 6 * B
        return o_g, o_c, o_t

    def pack(self, item_pair, JtJ, Jtr, residual, inliers, sizes):
        """Host mirror of dfk_window_assemble (numpy, float32 sums in item order): item i belongs to pair item_pair[i];
        JtJ [n, NP, NP] dense, Jtr [n, NP], sizes[i] = (W, H).  Returns the flat buffer."""
        K, B, c = self.num_keyframes, self.B, self.code_size
        out = np.zeros(self.floats, dtype=np.float32)
        o_g, o_c, o_t = self.offsets()
        D = out[:o_g].reshape(K, B, B)
        g = out[o_g:o_c].reshape(K, B)
        O = out[o_c:o_t].reshape(len(self.pairs), B, 6)
        loc0 = np.r_[0:6, 12:12 + c]  # [pose0 | code0] rows of a record
        f = np.float32(0)
        ninl = np.float32(0)
        for i, p in enumerate(item_pair):
            k0, k1 = self.pairs[p]
            H = np.asarray(JtJ[i], dtype=np.float32)
            r = np.asarray(Jtr[i], dtype=np.float32)
            D[k0] += H[np.ix_(loc0, loc0)]
            D[k1][:6, :6] += H[6:12, 6:12]
            g[k0] -= r[loc0]
            g[k1][:6] -= r[6:12]
            O[p] += H[np.ix_(loc0, np.arange(6, 12))]
            inl = int(inliers[i])
            if inl > 0:
                f += np.float32(residual[i]) / np.float32(inl) * np.float32(sizes[i][0] * sizes[i][1])
            ninl += np.float32(inl)
        out[o_t] = f
        out[o_t + 1] = ninl
        return out

    def to_dense(self, buf):
        """(H [dim, dim], g [dim], f, inliers) of the dense normal equations the buffer stands for (numpy float64, or
        torch float64 on the buffer's device)."""
        K, B = self.num_keyframes, self.B
        o_g, o_c, o_t = self.offsets()
        is_torch = hasattr(buf, "detach")
        if is_torch:
            import torch
            b64 = buf.detach().to(torch.float64)
            H = torch.zeros((K * B, K * B), dtype=torch.float64, device=buf.device)
        else:
            b64 = np.asarray(buf, dtype=np.float64)
            H = np.zeros((K * B, K * B))
        D = b64[:o_g].reshape(K, B, B)
        for k in range(K):
            H[k * B:(k + 1) * B, k * B:(k + 1) * B] += D[k]
        O = b64[o_c:o_t].reshape(len(self.pairs), B, 6)
        for p, (k0, k1) in enumerate(self.pairs):
            H[k0 * B:(k0 + 1) * B, k1 * B:k1 * B + 6] += O[p]
            H[k1 * B:k1 * B + 6, k0 * B:(k0 + 1) * B] += O[p].T if not is_torch else O[p].transpose(0, 1)
        g = b64[o_g:o_c].reshape(K * B)
        return H, g, float(b64[o_t]), float(b64[o_t + 1])


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
