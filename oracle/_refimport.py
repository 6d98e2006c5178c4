"""Import the UNMODIFIED reference (dlilien/ImpDAR) from /root/reference for fixture generation.

TEST INFRASTRUCTURE ONLY.  Used by tests/golden/make_golden.py and by the (container-only) tests that
pin the oracle against the running reference.  Nothing in the product path, the -m gpu tests, smoke()
or bench.py imports this module: /root/reference does not exist on the GPU box.

The reference cannot be imported as-is here: src/impdar/__init__.py imports matplotlib and
lib/ApresData/__init__.py imports h5py; both are absent.  Two empty stub modules make the hot-path
modules importable.
"""
import os
import sys
import types

REFERENCE_SRC = os.environ.get("IMPDAR_REFERENCE_SRC", "/root/reference/src")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_SRC, "impdar"))


def import_reference():
    """Return (impdar.lib.migrationlib.mig_python, RadarData class, NoInitRadarData module)."""
    if not reference_available():
        raise ImportError("reference tree not present at %s" % REFERENCE_SRC)
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib  # noqa: F401
        except ImportError:
            mpl = types.ModuleType("matplotlib")
            mpl.use = lambda *a, **k: None
            plt = types.ModuleType("matplotlib.pyplot")
            mpl.pyplot = plt
            for sub in ("colors", "widgets", "figure", "gridspec", "cm"):
                m = types.ModuleType("matplotlib." + sub)
                setattr(mpl, sub, m)
                sys.modules["matplotlib." + sub] = m
            sys.modules["matplotlib"] = mpl
            sys.modules["matplotlib.pyplot"] = plt
    if "h5py" not in sys.modules:
        try:
            import h5py  # noqa: F401
        except ImportError:
            sys.modules["h5py"] = types.ModuleType("h5py")
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)
    from impdar.lib.migrationlib import mig_python
    from impdar.lib.RadarData import RadarData
    from impdar.lib import NoInitRadarData
    return mig_python, RadarData, NoInitRadarData
