"""impdar_b200.process: the device-resident counterpart of impdar.lib.process.process (lib/process.py:72-197).
CPU part: the reference's argument checks (test/test_process.py of the reference exercises the same errors) and
the order in which steps are issued.  GPU part: results equal the step-by-step drop-in calls and match the oracle."""
import numpy as np
import pytest

import impdar_b200
from impdar_b200 import process as proc
from util import synthetic_dat


def _dat(S=50, T=40):
    return impdar_b200.RadarData(np.ones((S, T)), dt=1e-9, travel_time=0.001 * np.arange(S) + 0.001,
                                 dist=np.arange(T) * 1e-3, trace_int=np.ones(T))


def test_argument_checks_like_reference():
    assert proc.process([_dat()]) is False                      # nothing requested -> False (process.py:195-197)
    with pytest.raises(TypeError):
        proc.process([_dat()], crop=3)                          # 'Crop must be subscriptible'
    with pytest.raises(ValueError):
        proc.process([_dat()], crop=('a', 'top', 'snum'))
    with pytest.raises(TypeError):
        proc.process([_dat()], hcrop=3)
    with pytest.raises(ValueError):
        proc.process([_dat()], denoise=(1.0, 2))
    with pytest.raises(TypeError):
        proc.process([_dat()], vbp=3)
    with pytest.raises(ValueError):
        proc.process([_dat()], interp=('x',))
    with pytest.raises(Exception):                              # interp needs ImpDAR's gpslib and a GPS file
        proc.process([_dat()], interp=(10., 'no_such_gps_file.mat'))


def test_step_order_and_chain_split(monkeypatch):
    """Every device step forms one chain, in the reference's order (process.py:111-193), unless the reference
    orders a host step (denoise, interp) in between."""
    calls = []
    monkeypatch.setattr(proc, 'run_device_chain', lambda dats, steps, n: calls.append([s[0] for s in steps]))
    d = _dat()
    assert proc.process([d], vbp=(2, 10), hfilt=(0, 40), ahfilt=10, migrate=True)
    assert calls == [['vbp', 'hfilt', 'ahfilt', 'migrate']]
    del calls[:]
    assert proc.process([d], migrate=True, crop=(0.1, 'top', 'twtt'), nmo=10., ahfilt=10, vbp=(2, 10), rev=True,
                        restack=(3,), hcrop=(5, 'left', 'tnum'))
    assert calls == [['hcrop', 'restack', 'rev', 'vbp', 'ahfilt', 'nmo', 'crop', 'migrate']]
    del calls[:]
    assert proc.process([d], vbp=(2, 10), nmo=(10., 1.69e8), denoise=(3, 3), crop=(0.1, 'top', 'twtt'), migrate=True)
    assert calls == [['vbp', 'nmo', 'denoise', 'crop', 'migrate']]


def test_chain_result_dtypes_follow_reference():
    d = _dat()
    f32, f64 = np.dtype(np.float32), np.dtype(np.float64)
    assert proc._host_dtype_after([('vbp', (2, 10)), ('crop', (1., 'top', 'snum'))], f32, d) == f32
    assert proc._host_dtype_after([('hcrop', (5., 'left', 'tnum')), ('migrate', None)], f32, d) == f32
    assert proc._host_dtype_after([('restack', 3)], f32, d) == f64
    assert proc._host_dtype_after([('nmo', (10., 1.69e8)), ('migrate', None)], f32, d) == f64
    assert proc._host_dtype_after([('crop', (0., 'top', 'pretrig'))], f32, d) == f32     # scalar trigger: a view
    d.trig = np.zeros(d.tnum)
    assert proc._host_dtype_after([('crop', (0., 'top', 'pretrig'))], f32, d) == f64     # per-trace: NaN-padded f64


def test_process_sharded_split():
    dats = [_dat() for _ in range(5)]
    done, mine = proc.process_sharded(dats, rank=1, world=2)
    assert done is False and mine == [1, 3]


@pytest.mark.gpu
def test_process_matches_stepwise_and_oracle():
    from oracle import filtering as of, migration as om
    shapes = [(256, 512), (200, 300), (256, 512), (130, 64), (256, 512)]
    dats = [synthetic_dat(S, T, seed=11 + i) for i, (S, T) in enumerate(shapes)]
    dats[3].data = dats[3].data.astype(np.float64)
    ref = [synthetic_dat(S, T, seed=11 + i) for i, (S, T) in enumerate(shapes)]
    ref[3].data = ref[3].data.astype(np.float64)
    assert impdar_b200.process.process(dats, vbp=(2, 10), hfilt=(0, 64), ahfilt=30, migrate=True, n_streams=3)
    for d, r in zip(dats, ref):
        T = r.tnum
        r.vertical_band_pass(2, 10)
        r.hfilt(ftype='hfilt', bounds=(0, 64))
        r.hfilt(ftype='adaptive', window_size=30)
        r.migrate(mtype='stolt')
        assert isinstance(d.data, np.ndarray) and d.data.dtype == r.data.dtype and d.data.shape == r.data.shape
        assert d.flags.mig == 'stolt' and d.flags.bpass[0] == 1 and d.flags.hfilt[1] == 4
        if r.data.dtype == np.float32:
            assert np.array_equal(d.data, r.data)       # same kernels on the same inputs: bit-identical
        else:
            # stepwise, a float64 profile is rounded to fp32 only at the Stolt upload; identical here
            assert np.allclose(d.data, r.data, rtol=0, atol=1e-6 * np.abs(r.data).max())
    # against the oracle (float64 restatement of the reference) for one profile
    S, T = shapes[0]
    x64 = synthetic_dat(S, T, seed=11).data.astype(np.float64)
    tt = dats[0].travel_time
    y = of.vertical_band_pass(x64, 1e-8, 2, 10)
    y = of.horizontalfilt(y, tt, 0, 64)
    y = of.adaptivehfilt(y, tt, 30)
    _, want = om.stolt(y, 1e-8, dats[0].trace_int, dats[0].dist, 1.68e8, 10, 10)
    rel = np.linalg.norm(dats[0].data - want) / np.linalg.norm(want)
    print("process chain vs oracle rel-L2 %.3e" % rel)
    assert rel < 1e-5


@pytest.mark.gpu
def test_process_full_chain_matches_stepwise():
    """hcrop -> restack -> reverse -> vbp -> hfilt -> nmo -> crop -> migrate in one device-resident chain equals the
    same methods called one by one on host arrays (each of which is bit-exact / 1e-5 against the reference)."""
    kw = dict(hcrop=(9, 'left', 'tnum'), restack=3, rev=True, vbp=(2, 10), hfilt=(0, 64), nmo=(30., 1.69e8),
              crop=(0.3, 'top', 'twtt'), migrate=True)
    for dtype in (np.float32, np.float64):
        d = synthetic_dat(300, 520, seed=3, dtype=dtype)
        r = synthetic_dat(300, 520, seed=3, dtype=dtype)
        for o in (d, r):
            o.trig = np.zeros(o.tnum)
        assert impdar_b200.process.process([d], **kw)
        r.hcrop(*kw['hcrop']); r.restack(3); r.reverse(); r.vertical_band_pass(2, 10)
        r.hfilt(ftype='hfilt', bounds=(0, 64)); r.nmo(*kw['nmo']); r.crop(*kw['crop']); r.migrate(mtype='stolt')
        assert isinstance(d.data, np.ndarray) and d.data.dtype == r.data.dtype == np.float64
        assert d.data.shape == r.data.shape and d.snum == r.snum and d.tnum == r.tnum
        assert np.array_equal(d.travel_time, r.travel_time) and np.array_equal(d.dist, r.dist)
        assert np.allclose(d.data, r.data, rtol=0, atol=2e-6 * np.abs(r.data).max())
        assert d.flags.mig == 'stolt' and d.flags.restack and d.flags.reverse and d.flags.crop[0] == 1


@pytest.mark.gpu
def test_process_device_resident_input_stays_on_device():
    import torch
    d = synthetic_dat(128, 256, seed=5)
    r = synthetic_dat(128, 256, seed=5)
    d.data = torch.from_numpy(d.data).cuda()
    assert impdar_b200.process.process([d], vbp=(2, 10), migrate=True)
    assert isinstance(d.data, torch.Tensor) and d.data.is_cuda
    r.vertical_band_pass(2, 10)
    r.migrate(mtype='stolt')
    assert np.array_equal(d.data.cpu().numpy(), r.data)


@pytest.mark.gpu
def test_process_failing_profile_leaves_the_others_processed():
    """One profile in a longer list than n_streams makes a step raise (denoise on constant data: zero local variance).
    Like the reference's serial loop, every earlier profile is completely processed and holds a host array; nothing is
    left with data = None, and the failing profile keeps a host array (the state before the failing step)."""
    dats = [synthetic_dat(64, 96, seed=20 + i) for i in range(6)]
    ref = [synthetic_dat(64, 96, seed=20 + i) for i in range(6)]
    bad = 4
    dats[bad].data[:] = 1.0
    dats[bad].data[:, 40:] = np.arange(64, dtype=np.float32)[:, None]    # zero local variance on the left, not overall
    with pytest.raises(ValueError):
        impdar_b200.process.process(dats, rev=True, denoise=(1, 3), n_streams=3)
    for i, (d, r) in enumerate(zip(dats, ref)):
        assert isinstance(d.data, np.ndarray), "profile %d was left without a host array" % i
        if i < bad:
            r.reverse()
            r.denoise(1, 3)
            # the Wiener noise estimate is reduced with float64 atomics: equal to rounding, not bit for bit
            assert d.data.dtype == r.data.dtype and np.allclose(d.data, r.data, rtol=1e-10, atol=1e-12)
        elif i > bad:
            assert np.array_equal(d.data, r.data)        # never started
    assert dats[bad].data.shape == (64, 96) and np.all(np.isfinite(dats[bad].data))
