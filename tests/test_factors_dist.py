"""CPU tests of the host logic above the aligners: HessianFactor block slicing, window assembly, pair
sharding, and the world_size-2 all-reduce (gloo) that joins the shards -- driven by the oracle's
per-pair systems, so no GPU is needed."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from deepfactors_b200 import factors, synth


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _window_systems(oracle, n_kf=3, cs=8, w=80, h=60):
    """a ring of n_kf keyframes, each paired with its successor: per-pair systems from the oracle"""
    pairs, H, g, res, inl, sizes = [], [], [], [], [], []
    for k in range(n_kf):
        p = synth.make_pair(w, h, cs, 1, seed=30 + k, code_sigma=0.3, phase=0.1 * k)
        L = p.levels[0]
        r = oracle.sfm_run_step(p.pose0, p.pose1, L.cam, L.img0, L.img1, L.dpt0, None, L.prx_jac, L.grad1, precision="f64")
        pairs.append((k, (k + 1) % n_kf))
        H.append(r.dense()); g.append(r.Jtr); res.append(r.residual); inl.append(r.inliers); sizes.append((w, h))
    return pairs, np.stack(H), np.stack(g), np.array(res), np.array(inl), sizes


def test_record_unpack_roundtrip():
    cs = 8
    n, nh, rec = factors.record_layout(cs)
    assert (n, nh, rec) == (20, 210, 232)
    rng = np.random.default_rng(0)
    r = rng.standard_normal((3, rec)).astype(np.float32)
    r[:, nh + n + 1] = np.array([5, 70000, 0], dtype=np.uint32).view(np.float32)
    H, g, res, inl = factors.unpack_records(r, cs)
    assert H.shape == (3, n, n) and np.allclose(H, np.transpose(H, (0, 2, 1)))
    assert np.array_equal(inl, [5, 70000, 0])
    Ht, gt, rest, inlt = factors.unpack_records(torch.from_numpy(r), cs)
    assert np.allclose(Ht.numpy(), H) and np.array_equal(inlt.numpy(), inl)
    iu = np.triu_indices(n)
    assert np.allclose(H[1][iu], r[1, :nh])


def test_photometric_factor_blocks_follow_reference(oracle):
    pairs, H, g, res, inl, sizes = _window_systems(oracle, n_kf=1)
    Gs, gs, f = factors.photometric_factor_blocks(H[0], g[0], res[0], inl[0], 80, 60, 8)
    assert [G.shape for G in Gs] == [(6, 6), (6, 6), (6, 8), (6, 6), (6, 8), (8, 8)]
    assert np.allclose(gs[0], -g[0][:6]) and np.allclose(gs[2], -g[0][12:])        # photometric_factor.cpp:106
    assert np.isclose(f, res[0] / inl[0] * 80 * 60)                                  # :275-278
    assert factors.photometric_factor_blocks(H[0], g[0], 0.0, 0, 80, 60, 8)[2] == float("inf")  # :279-282


def test_window_assembly_is_sum_of_pair_blocks(oracle):
    pairs, H, g, res, inl, sizes = _window_systems(oracle, n_kf=3)
    lay = factors.WindowLayout(3, 8)
    Hw, gw, f = factors.assemble_window(lay, pairs, H, g, res, inl, sizes)
    assert Hw.shape == (42, 42) and np.allclose(Hw, Hw.T)
    assert np.linalg.eigvalsh(Hw).min() > -1e-8 * np.abs(Hw).max()
    b = lay.block
    # keyframe 1's pose block = pose0 block of pair (1,2) + pose1 block of pair (0,1)
    assert np.allclose(Hw[b:b + 6, b:b + 6], H[1][0:6, 0:6] + H[0][6:12, 6:12])
    # coupling pose1(k=1) <-> code0(k=0) comes only from pair (0,1)
    assert np.allclose(Hw[b:b + 6, 6:6 + 8], H[0][6:12, 12:20])
    assert np.allclose(gw[6:14], -(g[0][12:20]))
    # same thing through torch
    Ht, gt, ft = factors.assemble_window(lay, pairs, torch.from_numpy(H), torch.from_numpy(g), res, inl, sizes)
    assert np.allclose(Ht.numpy(), Hw) and np.allclose(gt.numpy(), gw) and np.isclose(ft, f)
    dx = factors.gauss_newton_step(Hw, gw, damping=1e-3)
    assert dx.shape == (42,) and np.all(np.isfinite(dx))


def test_shard_pairs_partitions_everything():
    for n in (0, 1, 7, 200, 2001):
        for world in (1, 2, 3, 8):
            got = [i for r in range(world) for i in factors.shard_pairs(n, world, r)]
            assert got == list(range(n))
            sizes = [len(factors.shard_pairs(n, world, r)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, payload, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pairs, H, g, res, inl, sizes, n_kf, cs = payload
        lay = factors.WindowLayout(n_kf, cs)
        mine = list(factors.shard_pairs(len(pairs), world, rank))
        Hw, gw, f = factors.assemble_window(lay, [pairs[i] for i in mine], torch.from_numpy(H[mine]),
                                            torch.from_numpy(g[mine]), res[mine], inl[mine], [sizes[i] for i in mine])
        factors.allreduce_window(Hw, gw)
        out[rank] = (Hw.numpy().copy(), gw.numpy().copy())
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_allreduce_equals_single_rank_window(oracle):
    pairs, H, g, res, inl, sizes = _window_systems(oracle, n_kf=3)
    lay = factors.WindowLayout(3, 8)
    H_ref, g_ref, _ = factors.assemble_window(lay, pairs, H, g, res, inl, sizes)
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(2, port, (pairs, H, g, res, inl, sizes, 3, 8), out), nprocs=2, join=True)
        for r in range(2):
            Hr, gr = out[r]
            assert np.allclose(Hr, H_ref, rtol=1e-12, atol=1e-9 * np.abs(H_ref).max())
            assert np.allclose(gr, g_ref, rtol=1e-12, atol=1e-9 * np.abs(g_ref).max())


# ---------------------------------------------------------------------------------------------------------------------
# The packed block-sparse window buffer (what dfk_window_assemble writes on the device and ONE all-reduce sums):
# host mirror factors.WindowBlocks vs the dense assembly, and the world_size-2 reduction of the packed buffers.
def test_block_sparse_buffer_expands_to_the_dense_window(oracle):
    pairs, H, g, res, inl, sizes = _window_systems(oracle, n_kf=3)
    wb = factors.WindowBlocks(3, 8, pairs)
    assert wb.floats == 3 * (14 * 14 + 14) + 3 * 6 * 14 + 2
    buf = wb.pack(list(range(len(pairs))), H, g, res, inl, sizes)
    Hd, gd, f, ninl = wb.to_dense(buf)
    H_ref, g_ref, f_ref = factors.assemble_window(factors.WindowLayout(3, 8), pairs, H, g, res, inl, sizes)
    assert np.allclose(Hd, H_ref, rtol=0, atol=2e-6 * np.abs(H_ref).max())   # fp32 buffer vs fp64 dense sums
    assert np.allclose(gd, g_ref, rtol=0, atol=2e-6 * np.abs(g_ref).max())
    assert abs(f - f_ref) <= 1e-5 * abs(f_ref) and ninl == float(inl.sum())
    # two levels per pair: both records of a pair add into the same blocks
    wb2 = factors.WindowBlocks(3, 8, pairs)
    buf2 = wb2.pack([0, 0, 1, 1, 2, 2], np.concatenate([H[[0]], H[[0]], H[[1]], H[[1]], H[[2]], H[[2]]]),
                    np.repeat(g, 2, axis=0), np.repeat(res, 2), np.repeat(inl, 2), [s for s in sizes for _ in range(2)])
    assert np.allclose(buf2[:-2], 2 * buf[:-2], rtol=1e-6, atol=1e-6 * np.abs(buf).max())


def _worker_sparse(rank, world, port, payload, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pairs, H, g, res, inl, sizes, n_kf, cs = payload
        wb = factors.WindowBlocks(n_kf, cs, pairs)          # every rank: the SAME layout (all pairs of the window) ...
        mine = list(factors.shard_pairs(len(pairs), world, rank))
        buf = wb.pack(mine, H[mine], g[mine], res[mine], inl[mine], [sizes[i] for i in mine])  # ... filled from its shard
        t = torch.from_numpy(buf)
        dist.all_reduce(t)                                    # the ONE collective of a sharded Gauss-Newton step
        out[rank] = t.numpy().copy()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_allreduce_of_the_block_sparse_buffer(oracle):
    pairs, H, g, res, inl, sizes = _window_systems(oracle, n_kf=4)
    wb = factors.WindowBlocks(4, 8, pairs)
    ref = wb.pack(list(range(len(pairs))), H, g, res, inl, sizes)
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker_sparse, args=(2, port, (pairs, H, g, res, inl, sizes, 4, 8), out), nprocs=2, join=True)
        for r in range(2):
            assert np.allclose(out[r], ref, rtol=1e-6, atol=1e-6 * np.abs(ref).max())
        assert np.array_equal(out[0], out[1])
    Hd, gd, f, _ = wb.to_dense(ref)
    H_ref, g_ref, _ = factors.assemble_window(factors.WindowLayout(4, 8), pairs, H, g, res, inl, sizes)
    assert np.allclose(Hd, H_ref, rtol=0, atol=2e-6 * np.abs(H_ref).max())
