"""ORACLE (test infrastructure only) for the index / resampling operations either side of the hot path
(SURVEY.md 8f rank 3): float64 / float32 numpy restatements of the radargram passes of the reference's
RadarData/_RadarDataProcessing.py.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg import
this module; the product never does.

Parity is PINNED: tests/golden/proc_*.npz hold the UNMODIFIED reference's object state before and after every
operation (tests/golden/make_golden_processing.py), and tests/test_processing.py checks these restatements
against them bit for bit.
"""
import numpy as np


def crop_block(data, r0, r1, c0, c1, flip_lr=False):
    """crop :299 / hcrop :400 / reverse :29 - a block of the radargram, optionally np.fliplr'ed."""
    out = np.asarray(data)[r0:r1, c0:c1]
    return np.ascontiguousarray(out[:, ::-1] if flip_lr else out)


def shift_traces(data, shift, snum_out):
    """crop(dimension='pretrig') with a trigger vector (:314-330, shift = trig) and elev_correct (:617-626,
    shift = -top_ind): out[i, t] = data[i + shift[t], t] inside the radargram, NaN outside; float64 result."""
    data = np.asarray(data)
    S, T = data.shape
    out = np.full((snum_out, T), np.nan)
    rows = np.arange(snum_out)[:, None] + np.asarray(shift, dtype=np.int64)[None, :]
    ok = (rows >= 0) & (rows < S)
    cols = np.broadcast_to(np.arange(T)[None, :], rows.shape)
    out[ok] = data[rows[ok], cols[ok]]
    return out


def restack_mean(data, traces):
    """restack :451-455: np.mean over consecutive groups of `traces` columns (numpy's pairwise summation along the
    contiguous axis, in the input precision for floating input), stored into a float64 array."""
    data = np.asarray(data)
    S, T = data.shape
    To = T // traces
    grouped = np.ascontiguousarray(data[:, :To * traces]).reshape(S, To, traces)
    return np.mean(grouped, axis=2).astype(np.float64)


def nodes_scipy_linear(x, x_new):
    """(lo, hi, w_hi, w_lo) of scipy.interpolate.interp1d._call_linear."""
    x = np.asarray(x, dtype=np.float64)
    x_new = np.asarray(x_new, dtype=np.float64)
    hi = np.searchsorted(x, x_new).clip(1, len(x) - 1).astype(int)
    lo = hi - 1
    return lo, hi, (x_new - x[lo]) / (x[hi] - x[lo]), (x[hi] - x_new) / (x[hi] - x[lo])


def interp_rows_scipy(data, x, x_new):
    """interp1d(x, data.T)(x_new).T for 2-D data (constant_sample_depth_spacing :60) and the per-trace interp1d of
    a float32 trace (nmo :166-169): the two-weight form, float64 weights."""
    lo, hi, w_hi, w_lo = nodes_scipy_linear(x, x_new)
    data = np.asarray(data)
    return w_hi[:, None] * data[hi] + w_lo[:, None] * data[lo]


def interp_cols_scipy(data, x, x_new, keep=None):
    """interp1d(x, data[:, keep])(x_new) (constant_space :554)."""
    data = np.asarray(data)
    if keep is not None:
        data = data[:, keep]
    lo, hi, w_hi, w_lo = nodes_scipy_linear(x, x_new)
    return w_hi[None, :] * data[:, hi] + w_lo[None, :] * data[:, lo]


def interp_rows_numpy(data, xp, x_new):
    """np.interp(x_new, xp, trace) for every float64 trace (what interp1d dispatches to for a 1-D float64 y,
    nmo :166-169), vectorised over traces with numpy.interp's own operation order."""
    data = np.asarray(data, dtype=np.float64)
    xp = np.asarray(xp, dtype=np.float64)
    x_new = np.asarray(x_new, dtype=np.float64)
    n = len(xp)
    j = np.searchsorted(xp, x_new, side='right') - 1
    last = j >= n - 1
    jj = np.clip(j, 0, n - 2)
    with np.errstate(invalid='ignore', divide='ignore'):
        slope = (data[jj + 1] - data[jj]) / (xp[jj + 1] - xp[jj])[:, None]
        out = slope * (x_new - xp[jj])[:, None] + data[jj]
        bad = np.isnan(out)
        if bad.any():
            alt = slope * (x_new - xp[jj + 1])[:, None] + data[jj + 1]
            out[bad] = alt[bad]
            still = np.isnan(out) & (data[jj] == data[jj + 1])
            out[still] = data[jj][still]
    exact = last | (x_new == xp[jj])
    src = np.where(last, n - 1, jj)
    out[exact] = data[src[exact]]
    return out


def nmo_times(travel_time, ant_sep, u_rms):
    """nmo :151-160 for a constant velocity: vertical two-way time [us] of every sample."""
    tsep = 1e6 * (ant_sep / u_rms)
    return np.sqrt((np.asarray(travel_time, dtype=np.float64) + tsep) ** 2. - tsep ** 2.)


def nmo_data(data, travel_time, dt, ant_sep, uice):
    """The radargram pass of nmo (:133-170) for a constant ice velocity -> (new data float64, new travel_time)."""
    travel_time = np.asarray(travel_time, dtype=np.float64)
    nmotime = np.array([np.sqrt((t + 1e6 * (ant_sep / uice)) ** 2. - (1e6 * (ant_sep / uice)) ** 2.)
                        for t in travel_time])
    new_tt = np.arange(min(travel_time), max(nmotime), dt * 1e6)
    data = np.asarray(data)
    if data.dtype == np.float32:
        return interp_rows_scipy(data, nmotime, new_tt), new_tt
    return interp_rows_numpy(data.astype(np.float64), nmotime, new_tt), new_tt
