"""Parity at the BASELINE.json shapes (the inputs the bench times), plus the reference's own C boundary.

At 2048 x 4096 (config 2), 8192 x 16384 and 8192 x 65536 (config 5) the full float64 image is out of the oracle's
reach, so blocks of output traces x strided output samples are checked with ``oracle.migration.kirchhoff_sparse``
(same arithmetic per pair as the reference loop, mig_python.py:35-60, which itself accepts reduced loop bounds).
The data are white noise: one flipped nearest-sample pick in 1e4 already costs ~1e-2 relative L2, and at config 5 each
output sums up to 2 x 1386 float32 terms, so the accumulation order is under test as well.  Tolerance = north-star:
relative L2 <= 1e-5 against the float64 oracle fed the float64 upcast of the same float32 input.
"""
import ctypes
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, load_golden, max_abs, rel_l2
from util import synthetic_dat

pytestmark = pytest.mark.gpu

TOL = 1e-5
VEL = 1.69e8


def _report(name, got, want):
    r, m = rel_l2(got, want), max_abs(got, want)
    print("%-40s rel-L2 %.3e  max-abs %.3e" % (name, r, m))
    return r


def _noise(S, T, seed):
    import torch
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    return torch.randn((S, T), generator=g, device="cuda", dtype=torch.float32)


def _geometry(S, T, jitter=0.0, seed=0):
    tt = np.arange(S) * 0.01
    dist = np.arange(T) * 0.005
    if jitter:
        dist = dist + jitter * 0.005 * (np.random.default_rng(seed).random(T) - 0.5)
    return tt, dist


def _check_blocks(x, tt, dist, nearfield, blocks, rows, mode, label):
    """Device result of the output-trace blocks `blocks` against the sparse oracle on `rows`."""
    from impdar_b200 import migrationlib as ml
    from oracle import migration as om
    import torch
    ml.set_kirchhoff_mode({"auto": ml.KIRCHHOFF_AUTO, "general": ml.KIRCHHOFF_GENERAL, "table": ml.KIRCHHOFF_TABLE}[mode])
    try:
        worst = 0.0
        for xb, xe in blocks:
            got = ml.kirchhoff_device(x, tt, dist, VEL, nearfield, xb, xe)
            torch.cuda.synchronize()
            path = ml.kirchhoff_last_path()
            want = om.kirchhoff_sparse(x, tt, dist, VEL, nearfield, range(xb, xe), rows)
            worst = max(worst, _report("%s [%d,%d) %s" % (label, xb, xe, path), got[rows].cpu().numpy(), want))
    finally:
        ml.set_kirchhoff_mode(ml.KIRCHHOFF_AUTO)
    return worst, path


@pytest.mark.parametrize("mode", ["table", "general"])
@pytest.mark.parametrize("nearfield", [False, True])
def test_kirchhoff_config2_shape(mode, nearfield):
    """BASELINE config 2: 4096 traces x 2048 samples; first / interior / last blocks of output traces, all rows."""
    S, T = 2048, 4096
    x = _noise(S, T, 21)
    tt, dist = _geometry(S, T)
    rows = np.arange(S)
    blocks = [(0, 4), (2045, 2051), (T - 3, T)] if not nearfield else [(1000, 1003)]
    worst, path = _check_blocks(x, tt, dist, nearfield, blocks, rows, mode, "kirch 2048x4096 nf=%d" % nearfield)
    assert path == mode
    assert worst < TOL


def test_kirchhoff_config2_jittered_positions_general():
    """Irregular trace positions (what a field profile looks like before constant_space) at the config-2 shape."""
    S, T = 2048, 4096
    x = _noise(S, T, 22)
    tt, dist = _geometry(S, T, jitter=0.6, seed=3)
    worst, path = _check_blocks(x, tt, dist, False, [(0, 2), (3001, 3004)], np.arange(S), "auto", "kirch 2048x4096 jitter")
    assert path == "general"
    assert worst < TOL


@pytest.mark.parametrize("mode", ["table", "general"])
def test_kirchhoff_8192x16384(mode):
    S, T = 8192, 16384
    x = _noise(S, T, 23)
    tt, dist = _geometry(S, T)
    rows = np.unique(np.concatenate([np.arange(0, S, 61), [S - 2, S - 1]]))
    blocks = [(0, 2), (8191, 8194)] if mode == "table" else [(8191, 8193)]
    worst, path = _check_blocks(x, tt, dist, False, blocks, rows, mode, "kirch 8192x16384")
    assert path == mode
    assert worst < TOL


def test_kirchhoff_config5_shape():
    """BASELINE config 5: 65536 traces x 8192 samples (2 GiB); up to 2773 float32 terms per output sample."""
    S, T = 8192, 65536
    x = _noise(S, T, 25)
    tt, dist = _geometry(S, T)
    rows = np.unique(np.concatenate([np.arange(0, S, 127), [1, S - 1]]))
    worst, path = _check_blocks(x, tt, dist, False, [(0, 2), (32767, 32770), (T - 2, T)], rows, "auto", "kirch 8192x65536")
    assert path == "table"
    assert worst < TOL


def test_kirchhoff_sharded_equals_unsharded():
    """Eight output-trace ranges (what eight ranks compute), each through the row-chunked entry the exchange pipeline
    uses, assembled == the unsharded image bit for bit."""
    import torch
    from impdar_b200 import migrationlib as ml, parallel
    S, T = 1024, 6000
    x = _noise(S, T, 27)
    tt, dist = _geometry(S, T)
    whole = ml.kirchhoff_device(x, tt, dist, VEL, False)
    ranges = parallel.kirchhoff_output_ranges(T, 8, tt, dist, VEL)
    assert ranges[0][0] == 0 and ranges[-1][1] == T
    parts = []
    for xb, xe in ranges:
        block = torch.empty((S, xe - xb), dtype=torch.float32, device="cuda")
        g_hi = S
        for r0, r1 in reversed(parallel.row_chunks(S, 4)):
            ml.kirchhoff_rows_device(x, tt, dist, VEL, False, xb, xe, r0, r1, g_hi, block)
            g_hi = r0
        parts.append(block)
    assert torch.equal(torch.cat(parts, dim=1), whole)
    one = ml.kirchhoff_device(x, tt, dist, VEL, False, ranges[3][0], ranges[3][1])
    assert torch.equal(one, parts[3])


@pytest.mark.parametrize("T", [6000, 6002])
def test_kirchhoff_column_window_equals_unsharded(T):
    """The unit of the multi-GPU exchange: a rank holds only the input columns its output range can read - as a packed
    copy (what a peer-mapped window buffer is) or as a strided slice of the image (what the holding rank uses) - and
    writes its block through a row stride (its place in the final image).  Bit for bit the unsharded image, for row
    chunks of unequal height, with and without the 128-bit d/dt pass (T = 6002: image rows are not 16-byte aligned)."""
    import torch
    from impdar_b200 import migrationlib as ml, parallel
    S = 1024
    x = _noise(S, T, 29)
    tt, dist = _geometry(S, T)
    whole = ml.kirchhoff_device(x, tt, dist, VEL, False)
    xd = x
    ranges = parallel.kirchhoff_output_ranges(T, 4, tt, dist, VEL)
    assert ranges[0][0] == 0 and ranges[-1][1] == T
    image = torch.full((S, T), float('nan'), dtype=torch.float32, device="cuda")
    for i, (xb, xe) in enumerate(ranges):
        c0, c1 = ml.kirchhoff_input_window(S, tt, dist, VEL, xb, xe)
        assert c0 % 4 == 0 and (c1 % 4 == 0 or c1 == T) and c0 <= xb and c1 >= xe and c1 - c0 < T
        win = xd[:, c0:c1].contiguous() if i % 2 == 0 else xd[:, c0:c1]
        g_hi = S
        for r0, r1 in reversed(parallel.row_chunks(S, parallel.DEFAULT_CHUNKS)):
            ml.kirchhoff_window_device(win, c0, T, tt, dist, VEL, False, xb, xe, image[:, xb:xe], (r0, r1, g_hi))
            g_hi = r0
    assert torch.equal(image, whole)


@pytest.mark.parametrize("layered", [False, True])
def test_phase_shift_4096_samples(layered):
    """S = nt = 4096 (config 3's depth: phases up to ~1e4 rad, 64 re-seeded recurrence blocks) on a trace crop the
    oracle can do."""
    from impdar_b200 import migrationlib as ml, synthetic
    from oracle import migration as om
    S, T = 4096, 96 if layered else 256
    d = synthetic_dat(S, T, seed=31)
    x64 = d.data.astype(np.float64)
    vel = synthetic.layered_velocity(d.travel_time) if layered else VEL
    _, want = om.phase_shift(x64, d.dt, d.travel_time, d.trace_int, d.dist, vel, 10, 10)
    ml.migrationPhaseShift(d, vel=vel, htaper=10, vtaper=10)
    assert d.data.shape == want.shape and d.data.dtype == np.float64
    assert _report("phsh %s 4096x%d" % ("layered" if layered else "const", T), d.data, want) < TOL


# ---------------------------------------------------------------------------------------- the C boundary itself
def _cdll():
    from impdar_b200 import _lib
    return _lib.load()


def _dp(a):
    return a.ctypes.data_as(ctypes.c_void_p)


@pytest.mark.parametrize("name", ["r64x128_kirch_far", "r65x50_kirch_far"])
def test_mig_kirch_loop_prototype(name):
    """The reference's own C prototype (migrationlib/mig_cython.h:11), called with exactly the arguments
    _mig_cython.pyx:36-47 passes: host float64 C-contiguous gradD / dist / zs / zs2 / tt_sec and the output array."""
    g = load_golden(name)
    lib = _cdll()
    data = np.ascontiguousarray(g["data"], dtype=np.float64)
    S, T = data.shape
    tt_sec = np.ascontiguousarray(g["travel_time"] / 1.0e6)                  # _mig_cython.pyx:75
    gradD = np.ascontiguousarray(np.gradient(data, tt_sec, axis=0))          # :70
    vel = float(g["vel"])
    zs = vel * tt_sec / 2.0                                                  # :78-79
    zs2 = zs ** 2.
    dist = np.ascontiguousarray(g["dist"] * 1.0e3)                           # :85
    migdata = np.zeros_like(data)                                            # :72
    lib.mig_kirch_loop(_dp(migdata), T, S, _dp(dist), _dp(zs), _dp(zs2), _dp(tt_sec), vel, _dp(gradD),
                       float(np.max(tt_sec)), 0)
    assert np.all(np.isfinite(migdata))
    assert _report(name + " mig_kirch_loop", migdata, g["out"]) < TOL


@pytest.mark.parametrize("name", ["r64x128_kirch_far", "r64x128_kirch_near"])
def test_kirchhoff_host_f64_entry(name):
    from impdar_b200 import _lib
    g = load_golden(name)
    lib = _cdll()
    data = np.ascontiguousarray(g["data"], dtype=np.float64)
    S, T = data.shape
    tt_sec = np.ascontiguousarray(g["travel_time"] / 1.0e6)
    dist = np.ascontiguousarray(g["dist"] * 1.0e3)
    out = np.empty_like(data)
    _lib.check(lib.impdar_kirchhoff_host_f64(_dp(data), _dp(out), S, T, _dp(dist), _dp(tt_sec), float(g["vel"]),
                                             int(bool(g["nearfield"]))))
    assert _report(name + " host_f64", out, g["out"]) < TOL
    rc = lib.impdar_kirchhoff_host_f64(_dp(data), None, S, T, _dp(dist), _dp(tt_sec), float(g["vel"]), 0)
    assert rc == 1 and b"null" in lib.impdar_b200_last_error()


def test_reference_cython_shim_links_against_the_library():
    """The drop-in claim end to end: the UNMODIFIED reference's compiled shim (impdar.lib.migrationlib.mig_cython,
    built by its own setup.py into baseline/_ref) has one undefined symbol, mig_kirch_loop - the C body is not in the
    reference tree.  With libimpdar_b200.so loaded RTLD_GLOBAL the shim imports and its migrationKirchhoff runs on
    the GPU.  Skipped where baseline/_ref was not installed."""
    ref = os.path.join(ROOT, "baseline", "_ref")
    import glob
    if not glob.glob(os.path.join(ref, "impdar", "lib", "migrationlib", "mig_cython*.so")):
        pytest.skip("baseline/_ref (pip install of the reference) is not present")
    _cdll()                                         # RTLD_GLOBAL: exports mig_kirch_loop to later dlopen()s
    from oracle import _refimport
    _refimport.stub_missing_modules()
    sys.path.insert(0, ref)
    try:
        for k in [k for k in sys.modules if k == "impdar" or k.startswith("impdar.")]:
            del sys.modules[k]
        from impdar.lib.migrationlib import mig_cython
        from impdar.lib.NoInitRadarData import NoInitRadarData
        g = load_golden("r64x128_kirch_far")
        dat = NoInitRadarData(big=True)
        dat.data = np.ascontiguousarray(g["data"], dtype=np.float64)
        dat.snum, dat.tnum = dat.data.shape
        dat.travel_time = g["travel_time"].copy()
        dat.dist = g["dist"].copy()
        dat.dt = float(g["dt"])
        mig_cython.migrationKirchhoff(dat, vel=float(g["vel"]), nearfield=False)
        assert _report("reference mig_cython -> libimpdar_b200", dat.data, g["out"]) < TOL
    finally:
        sys.path.remove(ref)
        for k in [k for k in sys.modules if k == "impdar" or k.startswith("impdar.")]:
            del sys.modules[k]


@pytest.mark.parametrize("S,T", [(256, 64), (300, 50), (1024, 256), (4096, 128), (5000, 40), (70, 33)])
def test_phase_shift_const_tensor_core_vs_simt(S, T):
    """Constant velocity: the tensor-core path (default; tcgen05 kind::tf32 with the 3xTF32 split and per-stage
    accumulator drains) and the SIMT (+w, -w) pair kernel against the float64 oracle - pow-2 and padded nt, odd tnum,
    nt < 64 (pair kernel only) and snum > 4096 (two accumulator passes)."""
    import torch
    from impdar_b200 import migrationlib as ml, _lib
    from oracle import migration as om
    lib = _lib.load()
    d = synthetic_dat(S, T, seed=S + T)
    x64 = d.data.astype(np.float64)
    _, want = om.phase_shift(x64, d.dt, d.travel_time, d.trace_int, d.dist, VEL, 10, 10)
    xd = torch.from_numpy(d.data).cuda()
    for mode, name in ((0, "tensor-core"), (3, "simt pair")):
        lib.impdar_phsh_set_legacy(mode)
        try:
            got = ml.phase_shift_device(xd, d.dt, 5.0, d.travel_time, VEL, 10, 10).double().cpu().numpy()
        finally:
            lib.impdar_phsh_set_legacy(0)
        assert _report("phsh const %dx%d %s" % (S, T, name), got, want) < TOL


def test_kirchhoff_tile_kernel_is_deterministic_and_used():
    """The shared-memory tile kernel is a producer / consumer pipeline over mbarriers (TMA writes, consumer reads,
    slot reuse) - a hazard compute-sanitizer's racecheck cannot model (it reports every TMA write / consumer read pair
    across the mbarriers, profiles/r02g_sanitizer_racecheck.txt).  A real race would show as run-to-run differences:
    five runs of the config-2 shape must agree bit for bit, with each other and with the image assembled from row
    chunks and trace ranges (different CTA decompositions of the same sums)."""
    import torch
    from impdar_b200 import migrationlib as ml
    S, T = 2048, 4096
    x = _noise(S, T, 41)
    tt, dist = _geometry(S, T)
    first = ml.kirchhoff_device(x, tt, dist, VEL, False)
    assert ml.kirchhoff_last_kernel() == "table_tile"
    for _ in range(4):
        assert torch.equal(ml.kirchhoff_device(x, tt, dist, VEL, False), first)
    parts = []
    for xb, xe in ((0, 1000), (1000, 1001), (1001, 3000), (3000, T)):
        block = torch.empty((S, xe - xb), dtype=torch.float32, device="cuda")
        g_hi = S
        for r0, r1 in ((1500, S), (700, 1500), (3, 700), (0, 3)):
            ml.kirchhoff_rows_device(x, tt, dist, VEL, False, xb, xe, r0, r1, g_hi, block)
            g_hi = r0
        parts.append(block)
    assert torch.equal(torch.cat(parts, dim=1), first)
    ml.set_kirchhoff_mode(ml.KIRCHHOFF_TABLE_GATHER)
    try:
        gather = ml.kirchhoff_device(x, tt, dist, VEL, False)
        assert ml.kirchhoff_last_kernel() == "table_gather"
    finally:
        ml.set_kirchhoff_mode(ml.KIRCHHOFF_AUTO)
    # same picks and weights, different summation order: equal to float32 rounding, not bit for bit
    assert float((gather - first).abs().max()) <= 2e-5 * float(first.abs().max())
