import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "reference: needs the reference tree at /root/reference (build container only)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_cuda = torch.cuda.is_available()
    except Exception:
        has_cuda = False
    from oracle._refimport import reference_available
    has_ref = reference_available()
    for item in items:
        if "gpu" in item.keywords and not has_cuda:
            item.add_marker(pytest.mark.skip(reason="no CUDA device"))
        if "reference" in item.keywords and not has_ref:
            item.add_marker(pytest.mark.skip(reason="reference tree not present"))


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """Make sure libimpdar_b200.so exists (nvcc cross-compiles without a GPU); never falls back."""
    from impdar_b200 import _build
    if _build.needs_build():
        _build.build()
    yield


def golden_names(prefix=None, contains=None):
    names = sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith(".npz"))
    if prefix:
        names = [n for n in names if n.startswith(prefix)]
    if contains:
        names = [n for n in names if contains in n]
    return names


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    return {k: z[k] for k in z.files}


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    m = np.isfinite(b)
    assert np.array_equal(np.isfinite(a), m), "finite/NaN pattern differs"
    den = np.linalg.norm(b[m])
    num = np.linalg.norm(a[m] - b[m])
    return num / den if den > 0 else num


def max_abs(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    m = np.isfinite(b)
    return float(np.max(np.abs(a[m] - b[m]))) if m.any() else 0.0
