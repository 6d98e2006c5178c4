"""Pin the oracle against the RUNNING reference (build container only: needs /root/reference)."""
import contextlib
import io

import numpy as np
import pytest

from conftest import rel_l2
from oracle import filtering as of
from oracle import migration as om

pytestmark = pytest.mark.reference


def _ref():
    from oracle._refimport import import_reference
    return import_reference()


def _quiet(f, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return f(*a, **k)


def _mk(RD, S, T, seed, tt0=0.0, dtype=np.float64):
    rng = np.random.default_rng(seed)
    d = RD(None)
    d.data = rng.standard_normal((S, T)).astype(dtype)
    d.snum, d.tnum, d.dt = S, T, 1e-8
    d.travel_time = tt0 + np.arange(S) * 0.01
    d.dist = np.arange(T) * 0.005
    d.trace_int = np.ones(T) * 5.0
    return d


@pytest.mark.parametrize("S,T,tt0", [(40, 33, 0.0), (37, 24, 0.02)])
def test_migrations(S, T, tt0):
    mp, RD, _ = _ref()
    for nf in (False, True):
        d = _mk(RD, S, T, 1, tt0); x = d.data.copy()
        _quiet(mp.migrationKirchhoff, d, vel=1.69e8, nearfield=nf)
        assert rel_l2(om.kirchhoff(x, d.travel_time, d.dist, 1.69e8, nf), d.data) < 1e-13
    d = _mk(RD, S, T, 2, tt0); x = d.data.copy()
    _quiet(mp.migrationStolt, d, vel=1.68e8, htaper=4, vtaper=6)
    assert rel_l2(om.stolt(x, d.dt, d.trace_int, d.dist, 1.68e8, 4, 6)[1], d.data) < 1e-13
    d = _mk(RD, S, T, 3, tt0); x = d.data.copy()
    _quiet(mp.migrationPhaseShift, d, vel=1.69e8, htaper=4, vtaper=6)
    assert rel_l2(om.phase_shift(x, d.dt, d.travel_time, d.trace_int, d.dist, 1.69e8, 4, 6)[1], d.data) < 1e-13
    d = _mk(RD, S, T, 4, tt0); x = d.data.copy()
    _quiet(mp.migrationTimeWavenumber, d, htaper=4, vtaper=6)
    assert np.array_equal(om.time_wavenumber(x, 4, 6), d.data)


def test_stolt_int_dtype_truncation():
    mp, RD, _ = _ref()
    d = _mk(RD, 32, 24, 5)
    d.data = (d.data * 10).astype(int); x = d.data.copy()
    _quiet(mp.migrationStolt, d, vel=1.68e8, htaper=4, vtaper=6)
    assert rel_l2(om.stolt(x, d.dt, d.trace_int, d.dist, 1.68e8, 4, 6)[1], d.data) < 1e-13


def test_filters_and_known_answer():
    _, RD, NI = _ref()
    f = NI.NoInitRadarDataFiltering()
    x = f.data.copy()
    assert np.all(of.horizontalfilt(x, f.travel_time, 0, 100) == f.hfilt_target_output)
    d = _mk(RD, 80, 30, 6, tt0=0.001); x = d.data.copy()
    _quiet(d.adaptivehfilt, 7)
    assert rel_l2(of.adaptivehfilt(x, d.travel_time, 7), d.data) < 1e-13
    d = _mk(RD, 150, 10, 7); x = d.data.copy()
    _quiet(d.vertical_band_pass, 2, 10)
    assert np.array_equal(of.vertical_band_pass(x, d.dt, 2, 10), d.data)


def test_reference_own_hot_path_tests_pass_here():
    """The reference's own test modules for the path run green in this container."""
    import os
    import subprocess
    import sys
    env = dict(os.environ)
    code = ("import sys, types; sys.path.insert(0, '/root/repo');"
            "from oracle._refimport import import_reference; import_reference();"
            "import pytest; sys.exit(pytest.main(['-q', '-x', '-p', 'no:cacheprovider',"
            "'/root/reference/test/test_migrationlib.py', '/root/reference/test/test_RadarDataFiltering.py']))")
    r = subprocess.run([sys.executable, "-c", code], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                       text=True, cwd="/tmp")
    assert r.returncode == 0, r.stdout[-2000:]
