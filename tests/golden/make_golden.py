#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference (dlilien/ImpDAR @ /root/reference).

Run in the build container only (the reference does not travel to the GPU box):

    python tests/golden/make_golden.py [--slow]

Each .npz holds the inputs (data, dt, travel_time [us], dist [km], trace_int [m], parameters) and the
reference's own output for one hot-path call.  ``--slow`` adds the 1598x85 tutorial radargram through
Kirchhoff (about 4 minutes of reference time) and phase shift (about 30 s).
"""
import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle._refimport import import_reference  # noqa: E402

mig_python, RadarData, NoInit = import_reference()
REF_ROOT = os.path.dirname(os.environ.get("IMPDAR_REFERENCE_SRC", "/root/reference/src"))


def quiet(f, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return f(*a, **k)


def make_dat(data, dt=1e-8, dx=5.0, tt0_us=0.0, dist_km=None):
    S, T = data.shape
    d = RadarData(None)
    d.data = data.copy()
    d.snum, d.tnum, d.dt = S, T, dt
    d.travel_time = tt0_us + np.arange(S) * dt * 1e6
    d.dist = np.arange(T) * dx / 1e3 if dist_km is None else dist_km
    d.trace_int = np.ones(T) * dx if dist_km is None else np.gradient(dist_km) * 1e3
    return d


def geom(d):
    return dict(dt=d.dt, travel_time=np.asarray(d.travel_time, dtype=np.float64),
                dist=np.asarray(d.dist, dtype=np.float64),
                trace_int=np.asarray(d.trace_int, dtype=np.float64))


def save(name, **kw):
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **kw)
    print("wrote", name)


def layered_velocity(d, n=12):
    zmax = 2.3e8 * d.travel_time[-1] * 1e-6 / 2
    z = np.linspace(0, 1.05 * zmax, n)
    v = 1.69e8 + (2.3e8 - 1.69e8) * np.exp(-z / (0.15 * zmax))
    return np.stack([v, z], 1)


def run_all_migrations(tag, data, nonuniform=False, **mk):
    S, T = data.shape
    dist_km = None
    if nonuniform:
        rng = np.random.default_rng(99)
        dist_km = np.cumsum(5.0 * (0.6 + 0.8 * rng.random(T))) / 1e3
        mk = dict(mk, dist_km=dist_km)
    for nf in (False, True):
        d = make_dat(data, **mk)
        quiet(mig_python.migrationKirchhoff, d, vel=1.69e8, nearfield=nf)
        save("%s_kirch_%s" % (tag, "near" if nf else "far"), data=data, out=d.data, vel=1.69e8,
             nearfield=nf, **geom(d))
    d = make_dat(data, **mk)
    quiet(mig_python.migrationStolt, d, vel=1.68e8, htaper=5, vtaper=7)
    save(tag + "_stolt", data=data, out=d.data, vel=1.68e8, htaper=5, vtaper=7, **geom(d))
    d = make_dat(data, **mk)
    quiet(mig_python.migrationPhaseShift, d, vel=1.69e8, htaper=5, vtaper=7)
    save(tag + "_phsh_const", data=data, out=d.data, vel=1.69e8, htaper=5, vtaper=7, **geom(d))
    if mk.get("tt0_us", 0.0) == 0.0:
        d = make_dat(data, **mk)
        vel = layered_velocity(d)
        vmig = quiet(mig_python.getVelocityProfile, d, vel)
        quiet(mig_python.migrationPhaseShift, d, vel=vel, htaper=5, vtaper=7)
        save(tag + "_phsh_layered", data=data, out=d.data, vel=vel, vmig=vmig, htaper=5, vtaper=7, **geom(d))
    d = make_dat(data, **mk)
    quiet(mig_python.migrationTimeWavenumber, d, vel=1.69e8, htaper=5, vtaper=7)
    save(tag + "_tk", data=data, out=d.data, htaper=5, vtaper=7, **geom(d))


def run_filters(tag, data, **mk):
    T = data.shape[1]
    mk = dict(mk)
    mk["tt0_us"] = mk.get("tt0_us", 0.0) + 0.001
    d = make_dat(data, **mk)
    quiet(d.horizontalfilt, 3, T // 2)
    save(tag + "_hfilt", data=data, out=d.data, ntr1=3, ntr2=T // 2, **geom(d))
    for w in (2, 10, T + 7, 2 * T):
        d = make_dat(data, **mk)
        quiet(d.adaptivehfilt, w)
        save("%s_ahfilt_w%d" % (tag, w), data=data, out=d.data, window_size=w, **geom(d))
    for ft in ("butter", "cheb", "bessel", "fir"):
        d = make_dat(data, **mk)
        kw = dict(order=20) if ft == "fir" else {}
        quiet(d.vertical_band_pass, 2., 10., filttype=ft, **kw)
        save("%s_vbp_%s" % (tag, ft), data=data, out=d.data, low=2., high=10., filttype=ft,
             order=kw.get("order", 5), **geom(d))


def run_siblings(tag, data, **mk):
    """winavg_hfilt, highpass / lowpass / horizontal_band_pass, rangegain, agc."""
    S, T = data.shape
    mk = dict(mk)
    mk["tt0_us"] = mk.get("tt0_us", 0.0) + 0.001
    for w, tp in ((1, 'full'), (7, 'full'), (10, 'pexp'), (T + 5, 'full')):
        d = make_dat(data, **mk)
        quiet(d.winavg_hfilt, w, taper=tp, filtdepth=S // 3)
        save("%s_winavg_w%d_%s" % (tag, w, tp), data=data, out=d.data, avg_win=w, taper=tp, filtdepth=S // 3, **geom(d))
    for name, args in (("highpass", (100.,)), ("lowpass", (60.,)), ("horizontal_band_pass", (40., 200.))):
        d = make_dat(data, **mk)
        d.flags.interp = np.array([1., 5.])
        quiet(getattr(d, name), *args)
        save("%s_%s" % (tag, name), data=data, out=d.data, args=np.array(args), tracespace=5., **geom(d))
    d = make_dat(data, **mk)
    d.trig = 3
    quiet(d.rangegain, 1.0e-2)
    save(tag + "_rangegain_scalar", data=data, out=d.data, trig=3, slope=1.0e-2, **geom(d))
    d = make_dat(data, **mk)
    d.trig = (np.arange(T) % 7).astype(float)
    quiet(d.rangegain, 2.5e-2)
    save(tag + "_rangegain_vector", data=data, out=d.data, trig=d.trig, slope=2.5e-2, **geom(d))
    d = make_dat(data, **mk)
    quiet(d.agc, window=20, scaling_factor=50)
    save(tag + "_agc", data=data, out=d.data, window=20, scaling_factor=50, **geom(d))
    run_denoise(tag, data, **mk)


def run_denoise(tag, data, **mk):
    """denoise(ftype='wiener') (_RadarDataFiltering.py:552-587): default window, a 2-D window, an even x even
    window, a given noise power; float32 input as well (scipy squares in the input precision)."""
    for name, dtype, kw in (("default", np.float64, {}), ("v3h5", np.float64, dict(vert_win=3, hor_win=5)),
                            ("v4h2", np.float64, dict(vert_win=4, hor_win=2)),
                            ("noise", np.float64, dict(vert_win=3, hor_win=7, noise=0.4)),
                            ("f32_v3h5", np.float32, dict(vert_win=3, hor_win=5))):
        d = make_dat(data.astype(dtype) + (3.0 if "f32" in name else 0.0), **mk)
        x = d.data.copy()
        quiet(d.denoise, **kw)
        save("%s_denoise_%s" % (tag, name), data=x, out=d.data, vert_win=kw.get("vert_win", 1),
             hor_win=kw.get("hor_win", 10), noise=kw.get("noise", np.nan), **geom(d))
    # ftype='median' (scipy.ndimage.median_filter): default window, odd and even 2-D windows, float32 and int16 data
    for name, conv, kw in (("v1h10", lambda a: a, {}), ("v3h3", lambda a: a, dict(vert_win=3, hor_win=3)),
                           ("v4h5_f32", lambda a: a.astype(np.float32), dict(vert_win=4, hor_win=5)),
                           ("v5h2_i16", lambda a: np.round(a * 100).astype(np.int16), dict(vert_win=5, hor_win=2))):
        d = make_dat(conv(data), **mk)
        x = d.data.copy()
        quiet(d.denoise, ftype='median', **kw)
        save("%s_median_%s" % (tag, name), data=x, out=d.data, vert_win=kw.get("vert_win", 1),
             hor_win=kw.get("hor_win", 10), **geom(d))


def run_lateral():
    """Laterally varying velocity (Fourier finite-difference branch, mig_python.py:428-432, 466-481, 496-540) with the
    reference's own test/input_data/velocity_lateral.txt table: even and odd shapes, plus the zeros fixture of
    test_migrationlib.py:133-135."""
    import warnings
    warnings.simplefilter("ignore")
    vel = np.genfromtxt(os.path.join(REF_ROOT, "test", "input_data", "velocity_lateral.txt"))
    for tag, shape, seed in [("r32x48", (32, 48), 21), ("r33x50", (33, 50), 22), ("r64x40", (64, 40), 23)]:
        rng = np.random.default_rng(seed)
        data = rng.standard_normal(shape)
        d = make_dat(data)
        vmig = quiet(mig_python.getVelocityProfile, d, vel)
        quiet(mig_python.migrationPhaseShift, d, vel=vel, htaper=5, vtaper=7)
        assert np.isfinite(d.data).all()
        save(tag + "_phsh_lateral", data=data, out=d.data, vel=vel, vmig=vmig, htaper=5, vtaper=7, **geom(d))
    f = NoInit.NoInitRadarData(big=True)
    x = f.data.copy()
    quiet(mig_python.migrationPhaseShift, f, vel_fn=os.path.join(REF_ROOT, "test", "input_data", "velocity_lateral.txt"))
    save("noinit_phsh_lateral", data=x, out=f.data, vel=vel, htaper=100, vtaper=1000, dt=f.dt,
         travel_time=np.asarray(f.travel_time, dtype=np.float64), dist=np.asarray(f.dist, dtype=np.float64),
         trace_int=np.asarray(f.trace_int, dtype=np.float64) * np.ones(f.tnum))


def main():
    slow = "--slow" in sys.argv
    if "--only-lateral" in sys.argv:
        run_lateral()
        return
    if "--only-denoise" in sys.argv:
        rng = np.random.default_rng(17)
        run_denoise("r96x160", rng.standard_normal((96, 160)) + 2.0, tt0_us=0.001)
        return
    if "--only-siblings" in sys.argv:
        rng = np.random.default_rng(17)
        run_siblings("r96x160", rng.standard_normal((96, 160)) + 2.0)
        return
    run_lateral()
    rng = np.random.default_rng(17)
    run_siblings("r96x160", rng.standard_normal((96, 160)) + 2.0)
    # (vi) seeded random shapes
    for tag, shape, seed, mk in [("r64x128", (64, 128), 11, {}),
                                 ("r65x50", (65, 50), 12, dict(tt0_us=0.013)),
                                 ("r128x200", (128, 200), 13, {})]:
        rng = np.random.default_rng(seed)
        run_all_migrations(tag, rng.standard_normal(shape), **mk)
    rng = np.random.default_rng(14)
    run_all_migrations("nu48x40", rng.standard_normal((48, 40)), nonuniform=True)
    rng = np.random.default_rng(15)
    run_filters("r256x96", rng.standard_normal((256, 96)) + 3.0)
    # NaN-padded input (crop/elev_correct leave NaNs; nansum skips them, mig_python.py:53)
    rng = np.random.default_rng(16)
    x = rng.standard_normal((40, 36))
    x[:3, 5:9] = np.nan
    x[-2:, 20] = np.nan
    d = make_dat(x)
    quiet(mig_python.migrationKirchhoff, d, vel=1.69e8, nearfield=True)
    save("nan40x36_kirch_near", data=x, out=d.data, vel=1.69e8, nearfield=True, **geom(d))

    # (i) the reference's own fixtures
    f = NoInit.NoInitRadarDataFiltering()
    x = f.data.copy()
    quiet(f.horizontalfilt, 0, 100)
    assert np.all(f.data == f.hfilt_target_output)
    save("noinit_hfilt", data=x, out=f.data, target=f.hfilt_target_output, ntr1=0, ntr2=100,
         dt=f.dt, travel_time=f.travel_time)

    # (ii) C1: small_data.mat through RadarData.migrate('stolt') (defaults htaper=vtaper=10, vel=1.68e8)
    fn = os.path.join(REF_ROOT, "test", "input_data", "small_data.mat")
    d = RadarData(fn)
    x = d.data.copy()
    quiet(d.migrate, mtype='stolt')
    save("c1_small_data_stolt", data=x, out=d.data, vel=1.68e8, htaper=10, vtaper=10, **geom(d))

    # (iv) tutorial synthetic radargram (gprMax box model, float32)
    fn = os.path.join(REF_ROOT, "doc", "impdar_tutorials", "migration", "data", "synthetic_radargram.mat")
    if os.path.exists(fn):
        d = RadarData(fn)
        x = d.data.copy()
        quiet(d.migrate, mtype='stolt', htaper=10, vtaper=20)
        save("tutorial_stolt", data=x, out=d.data.astype(np.float32), vel=1.68e8, htaper=10, vtaper=20, **geom(d))
        if slow:
            d = RadarData(fn)
            d.data = d.data.astype(np.float64)
            quiet(d.migrate, mtype='phsh', vel=1.69e8, htaper=10, vtaper=20)
            save("tutorial_phsh_const", data=x, out=d.data.astype(np.float32), vel=1.69e8, htaper=10, vtaper=20, **geom(d))
            d = RadarData(fn)
            d.data = d.data.astype(np.float64)
            quiet(d.migrate, mtype='kirch', vel=1.69e8)
            save("tutorial_kirch_far", data=x, out=d.data.astype(np.float32), vel=1.69e8, nearfield=False, **geom(d))


if __name__ == "__main__":
    main()
