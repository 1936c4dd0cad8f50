"""CPU tests of the drop-in boundary: libdfk.so loads and exports every symbol include/dfk.h declares
(no compute calls -- there is no GPU in the build container), the Python binding table is complete, and the
product path refuses to run without the CUDA extension (there is no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "dfk.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dfk_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_entry_points():
    names = declared_symbols()
    for must in ("dfk_sfm_run_step", "dfk_sfm_evaluate_error", "dfk_se3_run_step", "dfk_se3_warp", "dfk_update_depth",
                 "dfk_sfm_run_step_batch", "dfk_create", "dfk_destroy", "dfk_last_error"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from deepfactors_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "libdfk.so not built: run __graft_entry__.build()"
    h = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(h, name), f"libdfk.so does not export {name}"
    # and the binding table covers exactly the header
    assert sorted(_lib.SYMBOLS) == declared_symbols()


def test_non_compute_calls_work_without_a_gpu():
    from deepfactors_b200 import _lib
    L = _lib.lib()
    assert L.dfk_version() == 104  # DFK_VERSION of include/dfk.h
    assert L.dfk_status_string(0) == b"ok"
    assert L.dfk_sfm_supports_code_size(32) == 1
    assert L.dfk_sfm_supports_code_size(64) == 1 and L.dfk_sfm_supports_code_size(128) == 1
    assert L.dfk_sfm_supports_code_size(5) == 0
    assert _lib.record_floats(32) == 990 + 44 + 2


def test_missing_extension_fails_loudly(monkeypatch):
    from deepfactors_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libdfk.so")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.lib()


def test_product_package_never_touches_the_oracle():
    """only tests/, smoke() and bench.py's CPU-baseline legs may use oracle/"""
    pkg = os.path.join(ROOT, "deepfactors_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.lower().replace("# oracle-free", ""), f"{f} mentions the oracle"
    for f in os.listdir(os.path.join(ROOT, "include")):
        p = os.path.join(ROOT, "include", f)
        if os.path.isfile(p):
            assert "oracle" not in open(p).read().lower()
