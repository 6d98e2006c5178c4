"""The C-ABI library must load without a GPU and export every symbol include/impdar_b200.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "impdar_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{]*\)\s*;", src)
    return sorted(set(n for n in names if n.startswith("impdar_") or n == "mig_kirch_loop"))


def test_header_symbols_exported():
    from impdar_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 24
    for s in syms:
        assert hasattr(lib, s), "libimpdar_b200.so does not export %s" % s


def test_python_prototypes_cover_header():
    from impdar_b200 import _lib
    assert sorted(_lib.PROTOTYPES) == declared_symbols()


def test_load_and_errors_without_gpu():
    from impdar_b200 import _lib
    lib = _lib.load()
    assert lib.impdar_b200_version() == 100
    # argument validation happens before any CUDA call
    rc = lib.impdar_hfilt_f32(None, None, 4, 4, 1, 0, 4, None, 0, None)
    assert rc == 1 and "null" in _lib.last_error()
    rc = lib.impdar_stolt_f32(ctypes.c_void_p(8), ctypes.c_void_p(8), 1, 4, 1, 1e-8, 5.0, 1.68e8, 10., 10., 0,
                              None, 0, None)
    assert rc == 1
    assert lib.impdar_stolt_workspace_bytes(2048, 8192, 1) >= 2 * 1025 * 8192 * 8
    assert lib.impdar_filtfilt_workspace_bytes(100, 10, 2, 33, 4) == 2 * 166 * 10 * 4


def test_missing_library_fails_loudly(monkeypatch):
    from impdar_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libimpdar_b200.so")
    try:
        _lib.load()
    except ImportError as e:
        assert "no CPU fallback" in str(e)
    else:
        raise AssertionError("load() must fail when the CUDA library is missing")


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "impdar_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in src.replace("# oracle", ""), "%s references the oracle" % fn
