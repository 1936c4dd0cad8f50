"""include/df/dfk_factor.h (C++ PhotometricFactor block slicing + window assembly) against its Python mirror
deepfactors_b200/factors.py on the same LCG-generated aligner results.  Host-only: runs without a GPU."""
import json
import os
import subprocess

import numpy as np

from deepfactors_b200 import factors

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CS, NP = 8, 20


def lcg_items():
    s = 2024
    def rnd():
        nonlocal s
        s = (s * 1664525 + 1013904223) & 0xFFFFFFFF
        return np.float32(np.float32((s >> 8) & 0xFFFF) / np.float32(65535.0) - np.float32(0.5))
    items = []
    for i in range(5):
        packed = np.array([rnd() for _ in range(NP * (NP + 1) // 2)], dtype=np.float32)
        jtr = np.array([rnd() for _ in range(NP)], dtype=np.float32)
        res = np.float32(rnd() + np.float32(1.0))
        inl = 0 if i == 3 else 1000 + 17 * i
        H = np.zeros((NP, NP), dtype=np.float32)
        H[np.triu_indices(NP)] = packed
        H = H + np.triu(H, 1).T
        items.append((H, jtr, res, inl))
    return items


def test_cpp_factor_header_matches_python_mirror():
    subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "cpp"), "factor_test"], check=True, capture_output=True)
    out = subprocess.run([os.path.join(ROOT, "tests", "cpp", "factor_test")], check=True, capture_output=True, text=True)
    got = json.loads(out.stdout)
    items = lcg_items()
    pairs = [(0, 1), (1, 2), (2, 0), (0, 2), (1, 0)]
    sizes = [(640, 480), (320, 240), (160, 120), (80, 60), (640, 480)]
    lay = factors.WindowLayout(3, CS)
    H, g, f = factors.assemble_window(lay, pairs, np.stack([i[0] for i in items]), np.stack([i[1] for i in items]),
                                      [i[2] for i in items], [i[3] for i in items], sizes)
    assert got["dim"] == lay.dim and got["no_overlap_is_inf"] is True
    assert np.allclose(np.array(got["H"]).reshape(lay.dim, lay.dim), H, rtol=0, atol=1e-12)
    assert np.allclose(np.array(got["g"]), g, rtol=0, atol=1e-12)
    assert abs(got["f"] - f) <= 1e-9 * abs(f)
    Gs, gs, f0 = factors.photometric_factor_blocks(items[0][0], items[0][1], items[0][2], items[0][3], 640, 480, CS)
    for name, ref in zip(["G11", "G12", "G13", "G22", "G23", "G33"], Gs):
        assert np.allclose(np.array(got[name]).reshape(ref.shape), ref, rtol=0, atol=1e-12), name
    for name, ref in zip(["g1", "g2", "g3"], gs):
        assert np.allclose(np.array(got[name]), ref, rtol=0, atol=1e-12), name
    assert abs(got["f0"] - f0) <= 1e-9 * abs(f0)
