"""No-GPU checks of the drop-in boundary: libb2y.so builds, loads, and exports every symbol include/b2y.h declares."""
import ctypes
import os
import re

import pytest

from bs_yolo_b200 import _cabi as A

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(A.LIB_PATH):
        from bs_yolo_b200 import build
        build.build()
    return A.lib()


def test_exports_match_header(lib):
    hdr = open(os.path.join(ROOT, "include", "b2y.h")).read()
    declared = sorted(set(re.findall(r"B2Y_API\s+[\w\s\*]+?\b(b2y_\w+)\s*\(", hdr)))
    assert declared == sorted(A.EXPORTS)
    raw = ctypes.CDLL(A.LIB_PATH)
    for name in declared:
        assert hasattr(raw, name), name


def test_abi_version_and_struct_sizes(lib):
    assert lib.b2y_abi_version() == A.ABI_VERSION
    assert ctypes.sizeof(A.Buf) == 24
    assert ctypes.sizeof(A.Op) == 4 * 25 + 4 + 8 * 6 + 4 * 8 + 4 * 4  # 25 int32 + pad + 6 int64 + ip[8] + fp[4]


def test_error_codes_map_to_reference_exception_types(lib):
    # no compute: argument validation happens before any CUDA call
    with pytest.raises(ValueError):
        A.check(lib.b2y_preprocess(None, 1, 8, 8, None, 8, 8, 0, 0, 8, 8, 0, None))
    with pytest.raises(ValueError):
        A.check(lib.b2y_plan_run(None, 0, -1, None))
    assert b"plan_run" in lib.b2y_last_error()


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(A, "_lib", None)
    monkeypatch.setattr(A, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(A.B2YError, match="no CPU fallback"):
        A.lib()
