#!/usr/bin/env python
"""Generate tests/golden/proc_*.npz: object state before and after the UNMODIFIED reference's index / resampling
methods (RadarData/_RadarDataProcessing.py: reverse, crop, hcrop, restack, nmo, constant_sample_depth_spacing,
constant_space, elev_correct).  Build container only (the reference does not travel to the GPU box):

    python tests/golden/make_golden_processing.py

Each file holds ``in_<attr>`` / ``out_<attr>`` for every attribute in STATE plus ``call`` (method name) and
``kw_<name>`` (arguments); None attributes are left out.
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

STATE = ('data', 'travel_time', 'dist', 'trace_int', 'trace_num', 'trig', 'snum', 'tnum', 'dt', 'lat', 'long',
         'x_coord', 'y_coord', 'elev', 'decday', 'pressure', 'nmo_depth', 'elevation')
FLAGS = ('crop', 'nmo', 'interp', 'restack', 'reverse', 'elev')
RHO_PROFILE = "0,800\n50,900\n51,910\n100,910\n10000,910\n"      # content of the reference's test/input_data/rho_profile.txt


def make_dat(S=96, T=160, seed=11, dtype=np.float64, dt=1e-8, wander=False, integer=False):
    rng = np.random.default_rng(seed)
    d = RadarData(None)
    data = rng.standard_normal((S, T))
    d.data = (np.round(data * 1000).astype(np.int16) if integer else data.astype(dtype))
    d.snum, d.tnum, d.dt = S, T, dt
    d.travel_time = np.arange(S) * dt * 1e6
    steps = 5.0 + (rng.random(T) * 4.0 - 2.0 if wander else np.zeros(T))   # metres between traces
    if wander:
        steps[rng.integers(1, T, 6)] = 0.001                               # stationary shots (< min_movement)
    d.dist = np.cumsum(np.r_[0.0, steps[1:]]) / 1e3
    d.trace_int = np.gradient(d.dist) * 1e3
    d.trace_num = np.arange(T) + 1
    d.trig = np.zeros((T,))
    d.lat = -75.0 + np.cumsum(rng.random(T)) * 1e-4
    d.long = 110.0 + np.cumsum(rng.random(T)) * 1e-4
    d.x_coord = 1000.0 + d.dist * 1e3 + rng.random(T)
    d.y_coord = 2000.0 + rng.random(T) * 3
    d.elev = 800.0 + np.cumsum(rng.standard_normal(T)) * 0.4
    d.decday = 100.0 + np.arange(T) / 86400.
    d.pressure = np.zeros((T,))
    d.picks = None
    d.nmo_depth = None
    return d


def snapshot(d, prefix):
    out = {}
    for name in STATE:
        val = getattr(d, name, None)
        if val is not None:
            out[prefix + name] = np.array(val, copy=True)
    for name in FLAGS:
        out[prefix + 'flag_' + name] = np.array(getattr(d.flags, name), dtype=np.float64, copy=True)
    return out


def quiet(f, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return f(*a, **k)


def record(name, d, call, *args, prep=None, **kwargs):
    """Run d.<call>(*args, **kwargs) on the reference object and store before/after state."""
    if prep is not None:
        quiet(prep, d)
    blob = snapshot(d, 'in_')
    quiet(getattr(d, call), *args, **kwargs)
    blob.update(snapshot(d, 'out_'))
    blob['call'] = np.array(call)
    for i, a in enumerate(args):
        blob['arg_%d' % i] = np.array(a)
    for k, v in kwargs.items():
        blob['kw_' + k] = np.array(v)
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **blob)
    print('wrote', name, blob['out_data'].shape, blob['out_data'].dtype)


def with_rho_file(fn):
    path = os.path.join('/tmp', 'impdar_b200_rho_profile.txt')
    with open(path, 'w') as f:
        f.write(RHO_PROFILE)
    return fn(path)


def main():
    record('proc_reverse_f64', make_dat(), 'reverse')
    record('proc_reverse_i16', make_dat(integer=True), 'reverse')
    record('proc_crop_twtt_top', make_dat(), 'crop', 0.105, 'top', dimension='twtt')
    record('proc_crop_snum_bottom_f32', make_dat(dtype=np.float32), 'crop', 70, 'bottom', dimension='snum')
    record('proc_crop_depth_top', make_dat(), 'crop', 20.0, 'top', dimension='depth', uice=1.69e8, rezero=False)

    def trig_scalar(d):
        d.trig = 7
    record('proc_crop_pretrig_scalar', make_dat(), 'crop', 0, 'top', dimension='pretrig', prep=trig_scalar)

    def trig_vector(d):
        d.trig = np.random.default_rng(5).integers(2, 9, d.tnum)
    record('proc_crop_pretrig_vector', make_dat(), 'crop', 0, 'top', dimension='pretrig', prep=trig_vector)
    record('proc_crop_pretrig_vector_f32', make_dat(dtype=np.float32), 'crop', 0, 'top', dimension='pretrig',
           prep=trig_vector)
    record('proc_hcrop_tnum_left', make_dat(), 'hcrop', 23, 'left', dimension='tnum')
    record('proc_hcrop_dist_right_f32', make_dat(dtype=np.float32), 'hcrop', 0.5, 'right', dimension='dist')
    for n, dtype, tag in ((3, np.float64, 'f64'), (4, np.float32, 'f32'), (9, np.float32, 'f32'),
                          (21, np.float64, 'f64')):
        record('proc_restack_%d_%s' % (n, tag), make_dat(dtype=dtype), 'restack', n)
    record('proc_restack_131_f32', make_dat(S=24, T=1400, dtype=np.float32), 'restack', 131)
    record('proc_restack_301_f64', make_dat(S=24, T=1400), 'restack', 301)
    record('proc_restack_3_i16', make_dat(integer=True), 'restack', 3)
    record('proc_nmo_sep0_f64', make_dat(), 'nmo', 0.0, uice=1.69e8)
    record('proc_nmo_sep60_f64', make_dat(), 'nmo', 60.0, uice=1.69e8)
    record('proc_nmo_sep60_f32', make_dat(dtype=np.float32), 'nmo', 60.0, uice=1.69e8)
    record('proc_nmo_sep25_i16', make_dat(integer=True), 'nmo', 25.0, uice=1.69e8, const_firn_offset=3.0)
    with_rho_file(lambda p: record('proc_nmo_rho_f64', make_dat(), 'nmo', 30.0, rho_profile=p))
    with_rho_file(lambda p: record('proc_nmo_rho_const_sample_f32', make_dat(dtype=np.float32), 'nmo', 30.0,
                                   rho_profile=p, const_sample=True))

    def nmo_rho(d):
        with_rho_file(lambda p: d.nmo(10.0, rho_profile=p))
    record('proc_const_depth_spacing_f64', make_dat(), 'constant_sample_depth_spacing', prep=nmo_rho)
    record('proc_constant_space_f64', make_dat(wander=True), 'constant_space', 4.0)
    record('proc_constant_space_f32', make_dat(wander=True, dtype=np.float32), 'constant_space', 7.5,
           min_movement=3.5)

    def nmo_plain(d):
        d.nmo(0.0, uice=1.69e8)
    record('proc_elev_correct_f64', make_dat(), 'elev_correct', v_avg=1.69e8, prep=nmo_plain)
    record('proc_elev_correct_f32', make_dat(dtype=np.float32), 'elev_correct', prep=nmo_plain)


if __name__ == '__main__':
    main()
