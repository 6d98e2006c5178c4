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


def stub_missing_modules():
    """Empty stand-ins for the two modules the reference imports unconditionally and this image lacks."""
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


def import_reference():
    """Return (impdar.lib.migrationlib.mig_python, RadarData class, NoInitRadarData module)."""
    if not reference_available():
        raise ImportError("reference tree not present at %s" % REFERENCE_SRC)
    stub_missing_modules()
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)
    from impdar.lib.migrationlib import mig_python
    from impdar.lib.RadarData import RadarData
    from impdar.lib import NoInitRadarData
    return mig_python, RadarData, NoInitRadarData


VENDORED = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")


def vendored_available():
    """baseline/_ref: `pip install --target` of the UNMODIFIED reference (git-ignored, travels to the GPU box)."""
    return os.path.isfile(os.path.join(VENDORED, "impdar", "lib", "migrationlib", "mig_python.py"))


def import_vendored_mig_python():
    """The reference's own mig_python module from baseline/_ref, loaded standalone (it imports only numpy/scipy),
    so nothing else of the package - and none of this repo's code - is on the path it runs."""
    import importlib.util
    if not vendored_available():
        raise ImportError("baseline/_ref is not installed")
    name = "_impdar_ref_mig_python"
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(
        name, os.path.join(VENDORED, "impdar", "lib", "migrationlib", "mig_python.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod
