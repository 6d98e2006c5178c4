"""The oracle (oracle/*.py) must reproduce the golden vectors, which are outputs of the unmodified
reference (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from conftest import golden_names, load_golden, rel_l2
from oracle import filtering as of
from oracle import migration as om

TOL = 1e-12


@pytest.mark.parametrize("name", golden_names(contains="_kirch_"))
def test_kirchhoff(name):
    g = load_golden(name)
    if g["data"].size > 40000:
        pytest.skip("oracle Kirchhoff on the tutorial radargram is covered by the gpu suite's golden check")
    out = om.kirchhoff(g["data"], g["travel_time"], g["dist"], float(g["vel"]), bool(g["nearfield"]))
    assert rel_l2(out, g["out"]) < TOL


@pytest.mark.parametrize("name", golden_names(contains="_stolt"))
def test_stolt(name):
    g = load_golden(name)
    _, out = om.stolt(g["data"].astype(np.float64), float(g["dt"]), g["trace_int"], g["dist"], float(g["vel"]),
                      float(g["htaper"]), float(g["vtaper"]))
    tol = 2e-6 if g["out"].dtype == np.float32 else TOL   # tutorial fixture: float32 in the reference itself
    assert out.shape == g["out"].shape
    assert rel_l2(out, g["out"]) < tol


@pytest.mark.parametrize("name", golden_names(contains="_phsh_"))
def test_phase_shift(name):
    g = load_golden(name)
    if g["data"].size > 40000:
        pytest.skip("too slow for the CPU suite")
    _, out = om.phase_shift(g["data"].astype(np.float64), float(g["dt"]), g["travel_time"], g["trace_int"],
                            g["dist"], g["vel"] if g["vel"].ndim else float(g["vel"]),
                            float(g["htaper"]), float(g["vtaper"]))
    assert rel_l2(out, g["out"]) < TOL


@pytest.mark.parametrize("name", golden_names(contains="_tk"))
def test_time_wavenumber(name):
    g = load_golden(name)
    out = om.time_wavenumber(g["data"], float(g["htaper"]), float(g["vtaper"]))
    assert np.array_equal(out, g["out"])


@pytest.mark.parametrize("name", golden_names(contains="_hfilt"))
def test_hfilt(name):
    g = load_golden(name)
    out = of.horizontalfilt(g["data"], g["travel_time"], int(g["ntr1"]), int(g["ntr2"]))
    assert np.array_equal(out, g["out"])
    if "target" in g:   # the reference's exact known-answer fixture (test_RadarDataFiltering.py:53-57)
        assert np.all(out == g["target"])


@pytest.mark.parametrize("name", golden_names(contains="_ahfilt_"))
def test_ahfilt(name):
    g = load_golden(name)
    out = of.adaptivehfilt(g["data"], g["travel_time"], int(g["window_size"]))
    assert rel_l2(out, g["out"]) < TOL
    out2 = of.adaptivehfilt_loops(g["data"], g["travel_time"], int(g["window_size"]))
    assert np.array_equal(out2, g["out"])


@pytest.mark.parametrize("name", golden_names(contains="_vbp_"))
def test_vbp(name):
    g = load_golden(name)
    out = of.vertical_band_pass(g["data"], float(g["dt"]), float(g["low"]), float(g["high"]),
                                order=int(g["order"]), filttype=str(g["filttype"]))
    assert np.array_equal(out, g["out"])
    if str(g["filttype"]) != "fir":
        b, a = of.bandpass_coefficients(float(g["low"]), float(g["high"]), float(g["dt"]),
                                        order=int(g["order"]), filttype=str(g["filttype"]))
        assert rel_l2(of.filtfilt_explicit(b, a, g["data"]), g["out"]) < 1e-6


def test_loop_flavours_match():
    rng = np.random.default_rng(5)
    x = rng.standard_normal((24, 20))
    tt = np.arange(24) * 0.01
    dist = np.arange(20) * 0.005
    a = om.kirchhoff(x, tt, dist, 1.69e8, True)
    b = om.kirchhoff_loops(x, tt, dist, 1.69e8, True)
    assert rel_l2(a, b) < 1e-14
    _, s1 = om.stolt(x, 1e-8, np.ones(20) * 5., dist, 1.68e8, 3, 4)
    _, s2 = om.stolt_loops(x, 1e-8, np.ones(20) * 5., dist, 1.68e8, 3, 4)
    assert rel_l2(s1, s2) < 1e-12
    assert om.kirchhoff_pair_count(tt, dist, 1.69e8) > 0


# ------------------------------------------------------------------ sibling filters
@pytest.mark.parametrize("name", golden_names(contains="_winavg_"))
def test_winavg(name):
    g = load_golden(name)
    out = of.winavg_hfilt(g["data"], g["travel_time"], int(g["avg_win"]), str(g["taper"]), int(g["filtdepth"]))
    assert np.array_equal(np.isnan(out), np.isnan(g["out"]))
    assert np.array_equal(np.nan_to_num(out), np.nan_to_num(g["out"]))


@pytest.mark.parametrize("name", ["r96x160_highpass", "r96x160_lowpass", "r96x160_horizontal_band_pass"])
def test_horizontal_iir(name):
    g = load_golden(name)
    if name.endswith("highpass"):
        out = of.highpass(g["data"], float(g["args"][0]), float(g["tracespace"]), float(g["dt"]))
    elif name.endswith("lowpass"):
        out = of.lowpass(g["data"], float(g["args"][0]), float(g["tracespace"]), float(g["dt"]))
    else:
        out = of.horizontal_band_pass(g["data"], float(g["args"][0]), float(g["args"][1]), float(g["tracespace"]))
    assert out.dtype == np.float64 and np.array_equal(out, g["out"])


@pytest.mark.parametrize("name", golden_names(contains="_rangegain_") + golden_names(contains="_agc"))
def test_gains(name):
    g = load_golden(name)
    if "_agc" in name:
        out = of.agc(g["data"], int(g["window"]), int(g["scaling_factor"]))
    else:
        trig = g["trig"] if g["trig"].ndim else int(g["trig"])
        out = of.rangegain(g["data"], g["travel_time"], trig, float(g["slope"]))
    assert np.array_equal(out, g["out"])


@pytest.mark.parametrize("name", golden_names(contains="_denoise_"))
def test_denoise_wiener(name):
    """The reference's scipy.signal.wiener runs an FFT correlation (complex64 for float32 input), the oracle sums the
    window directly in float64: agreement to rounding noise of the reference (1e-15 float64, 2e-7 float32 input)."""
    g = load_golden(name)
    noise = None if np.isnan(g["noise"]) else float(g["noise"])
    out = of.wiener(g["data"], int(g["vert_win"]), int(g["hor_win"]), noise)
    assert out.dtype == np.float64 and out.shape == g["out"].shape
    rel = np.linalg.norm(out - g["out"]) / np.linalg.norm(g["out"])
    assert rel < (1e-6 if g["data"].dtype == np.float32 else 1e-13)


def test_denoise_zero_variance_raises_like_reference():
    with pytest.raises(ValueError):
        of.wiener(np.pad(np.ones((4, 40)), ((0, 4), (0, 0))) + np.arange(8)[:, None] * np.r_[np.zeros(20), np.ones(20)], 1, 3)


@pytest.mark.parametrize("name", golden_names(contains="_median_"))
def test_denoise_median(name):
    g = load_golden(name)
    out = of.median_filter(g["data"], int(g["vert_win"]), int(g["hor_win"]))
    assert out.dtype == g["out"].dtype == g["data"].dtype and np.array_equal(out, g["out"])


def test_sparse_and_trace_oracles_equal_the_full_ones():
    """The large-shape variants (selected output samples / traces) restate the same arithmetic as the full-image
    functions that the golden vectors pin."""
    from oracle import migration as om
    rng = np.random.default_rng(0)
    S, T = 150, 260
    x = rng.standard_normal((S, T)).astype(np.float32)
    tt, dk = 0.05 + np.arange(S) * 0.01, np.arange(T) * 0.002
    for nf in (False, True):
        full = om.kirchhoff(x.astype(np.float64), tt, dk, 1.69e8, nf)
        rows = np.array([0, 3, 77, S - 1])
        traces = [0, 1, 130, T - 1]
        part = om.kirchhoff_sparse(x, tt, dk, 1.69e8, nf, traces, rows)
        assert np.max(np.abs(part - full[rows][:, traces])) <= 1e-13 * np.max(np.abs(full))
    xn = x.copy()
    xn[40:60, 100:120] = np.nan
    full = om.kirchhoff(xn.astype(np.float64), tt, dk, 1.69e8, True)
    part = om.kirchhoff_sparse(xn, tt, dk, 1.69e8, True, [90, 110, 140])
    assert np.array_equal(np.isnan(part), np.isnan(full[:, [90, 110, 140]]))
    assert np.nanmax(np.abs(part - full[:, [90, 110, 140]])) <= 1e-13 * np.nanmax(np.abs(full))
    for (S, T) in ((64, 96), (65, 50)):
        x = rng.standard_normal((S, T)).astype(np.float32)
        ti, dk = np.ones(T) * 5.0, np.arange(T) * 0.005
        _, full = om.stolt(x.astype(np.float64), 1e-8, ti, dk, 1.68e8, 10, 7)
        tr = [0, 1, T // 2, T - 1]
        part = om.stolt_traces(x, 1e-8, ti, dk, 1.68e8, 10, 7, traces=tr, row_chunk=13)
        assert part.shape == (full.shape[0], 4)
        assert np.max(np.abs(part - full[:, tr])) <= 1e-13 * np.max(np.abs(full))
