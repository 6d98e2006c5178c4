"""CPU oracle for the filtering hot path: float64 numpy/scipy restatement of
RadarData/_RadarDataFiltering.py (vertical_band_pass, horizontalfilt, adaptivehfilt).

TEST INFRASTRUCTURE ONLY (see oracle/migration.py header).  Parity is PINNED against the running
reference (tests/test_oracle_vs_reference.py) and against tests/golden/*.npz generated from it,
including the reference's exact known-answer fixture ``hfilt_target_output``
(NoInitRadarData.py:74-76, test_RadarDataFiltering.py:53-57).
"""
import numpy as np


def exp_taper(travel_time_us):
    """exp(-0.05 tt)/exp(-0.05 tt[0]), _RadarDataFiltering.py:59 and :129-130."""
    tt = np.asarray(travel_time_us, dtype=np.float64).flatten()
    return np.exp(-tt * 0.05) / np.exp(-tt[0] * 0.05)


def hfilt_bounds(ntr1, ntr2, tnum):
    """_RadarDataFiltering.py:122-123."""
    htr1 = int(max(0, min(ntr1, tnum - 1)))
    htrn = int(max(htr1 + 1, min(ntr2, tnum)))
    return htr1, htrn


def horizontalfilt(data, travel_time_us, ntr1, ntr2):
    """_RadarDataFiltering.py:93-135."""
    data = np.asarray(data)
    htr1, htrn = hfilt_bounds(ntr1, ntr2, data.shape[1])
    avg_trace = np.mean(data[:, htr1:htrn], axis=-1)
    avg_trace = avg_trace * exp_taper(travel_time_us)
    return data - np.atleast_2d(avg_trace).transpose().astype(data.dtype)


def adaptive_window(i, tnum, window_size):
    """Column window [lo, hi) of trace i, _RadarDataFiltering.py:67-72, with Python's slice rules
    (negative start wraps once, bounds clip)."""
    h = window_size // 2
    if i <= h:
        sl = slice(0, h + i)
    elif i >= tnum - h:
        sl = slice(int(tnum) - window_size, int(tnum))
    else:
        sl = slice(i - h + 1, i + h)
    lo, hi, _ = sl.indices(tnum)
    return lo, max(lo, hi)


def adaptivehfilt_loops(data, travel_time_us, window_size):
    """Same loop as the reference (_RadarDataFiltering.py:65-85): per trace re-mean + scipy filtfilt."""
    from scipy.signal import filtfilt
    data = np.asarray(data)
    S, T = data.shape
    scale = exp_taper(travel_time_us)
    out = np.zeros_like(data, dtype=data.dtype)
    for i in range(int(T)):
        if i <= window_size // 2:
            pk = data[:, 0:window_size // 2 + i].copy()
        elif i >= T - window_size // 2:
            pk = data[:, int(T) - window_size:int(T)].copy()
        else:
            pk = data[:, i - window_size // 2 + 1:i + window_size // 2].copy()
        low = filtfilt([.25, .25, .25, .25], 1, np.mean(pk, axis=-1)).flatten() * scale.flatten()
        out[:, i] = data[:, i].copy() - low
    return out.astype(data.dtype)


def adaptivehfilt(data, travel_time_us, window_size):
    """Vectorised restatement of adaptivehfilt (_RadarDataFiltering.py:19-90).

    mean over the per-trace window via float64 prefix sums; ``filtfilt([.25]*4, 1, m)`` is the 7-tap
    triangular kernel [1,2,3,4,3,2,1]/16 applied to the odd-extended mean trace (padlen 12 > 3 taps
    of memory, so the lfilter_zi transients never reach the kept samples)."""
    data = np.asarray(data)
    S, T = data.shape
    if S <= 12:
        raise ValueError("The length of the input vector x must be greater than padlen, which is 12.")
    d64 = data.astype(np.float64)
    P = np.concatenate([np.zeros((S, 1)), np.cumsum(d64, axis=1)], axis=1)
    lo = np.empty(T, dtype=np.int64)
    hi = np.empty(T, dtype=np.int64)
    for i in range(T):
        lo[i], hi[i] = adaptive_window(i, T, window_size)
    with np.errstate(invalid='ignore', divide='ignore'):
        m = (P[:, hi] - P[:, lo]) / (hi - lo)[None, :]
    ext = np.concatenate([2 * m[0:1] - m[3:0:-1], m, 2 * m[-1:] - m[-2:-5:-1]], axis=0)
    c = np.array([1, 2, 3, 4, 3, 2, 1]) / 16.
    low = sum(c[j] * ext[j:j + S] for j in range(7))
    out = d64 - low * exp_taper(travel_time_us)[:, None]
    return out.astype(data.dtype)


def bandpass_coefficients(low_mhz, high_mhz, dt, order=5, filttype='butter', cheb_rp=5):
    """Corner frequencies and (b, a) of vertical_band_pass, _RadarDataFiltering.py:510-535."""
    from scipy.signal import butter, cheby1, bessel
    nyq = 0.5 * (1.0 / dt)
    corner = np.zeros((2,))
    corner[0] = low_mhz * 1.0e6 / nyq
    corner[1] = high_mhz * 1.0e6 / nyq
    ft = filttype.lower()
    if ft in ['butter', 'butterworth']:
        return butter(order, corner, 'bandpass')
    if ft in ['cheb', 'chebyshev']:
        return cheby1(order, cheb_rp, corner, 'bandpass')
    if ft == 'bessel':
        return bessel(order, corner, 'bandpass')
    raise ValueError('Filter type {:s} is not recognized'.format(filttype))


def vertical_band_pass(data, dt, low_mhz, high_mhz, order=5, filttype='butter', cheb_rp=5,
                       fir_window='hamming'):
    """_RadarDataFiltering.py:469-549 (data part; flags are host bookkeeping)."""
    from scipy.signal import filtfilt, firwin, lfilter
    data = np.asarray(data)
    if filttype.lower() == 'fir':
        nyq = 0.5 * (1.0 / dt)
        corner = np.array([low_mhz * 1.0e6 / nyq, high_mhz * 1.0e6 / nyq])
        taps = firwin(order + 1, corner, pass_zero=False)
        out = np.array(data, copy=True)
        out[:-order, :] = lfilter(taps, 1.0, data, axis=0).astype(data.dtype)[order:, :]
        return out
    b, a = bandpass_coefficients(low_mhz, high_mhz, dt, order, filttype, cheb_rp)
    return filtfilt(b, a, data, axis=0).astype(data.dtype)


def filtfilt_explicit(b, a, x, padlen=None):
    """scipy.signal.filtfilt(b, a, x, axis=0) (padtype='odd', method='pad') written out as the plain
    transposed direct-form-II recurrence the CUDA kernel implements; used to pin the recurrence
    itself (state layout, zi scaling, odd extension) against scipy."""
    from scipy.signal import lfilter_zi
    b = np.atleast_1d(np.asarray(b, dtype=np.float64))
    a = np.atleast_1d(np.asarray(a, dtype=np.float64))
    n = max(len(a), len(b))
    if padlen is None:
        padlen = 3 * n
    bb = np.zeros(n)
    aa = np.zeros(n)
    bb[:len(b)] = b / a[0]
    aa[:len(a)] = a / a[0]
    zi = lfilter_zi(bb, aa) if n > 1 else np.zeros(0)
    x = np.asarray(x, dtype=np.float64)
    S = x.shape[0]
    if S <= padlen:
        raise ValueError("The length of the input vector x must be greater than padlen, which is %d." % padlen)
    ext = np.concatenate([2 * x[0:1] - x[padlen:0:-1], x, 2 * x[-1:] - x[-2:-padlen - 2:-1]], axis=0)

    def run(sig):
        z = zi[:, None] * sig[0][None, :] if sig.ndim == 2 else zi * sig[0]
        y = np.empty_like(sig)
        for k in range(sig.shape[0]):
            xk = sig[k]
            yk = bb[0] * xk + (z[0] if n > 1 else 0.0)
            for i in range(n - 2):
                z[i] = bb[i + 1] * xk - aa[i + 1] * yk + z[i + 1]
            if n > 1:
                z[n - 2] = bb[n - 1] * xk - aa[n - 1] * yk
            y[k] = yk
        return y

    y = run(ext)
    y = run(y[::-1])[::-1]
    return y[padlen:padlen + S]


# --------------------------------------------------------------------------------------------
# Sibling filters: winavg_hfilt, highpass / lowpass / horizontal_band_pass, rangegain, agc
# --------------------------------------------------------------------------------------------


def winavg_taper(travel_time_us, snum, taper='full', filtdepth=100):
    """_RadarDataFiltering.py:399-417 (the hard-coded 'tukey' taper is marked unused there and is not restated)."""
    exptaper = exp_taper(travel_time_us)
    if taper == 'full':
        pass
    elif taper == 'pexp':
        exptaper[:filtdepth] = exptaper[:filtdepth] - exptaper[filtdepth]
        exptaper[filtdepth:snum] = 0
        exptaper = exptaper / np.max(exptaper)
    else:
        raise ValueError('Unrecognized taper. Options are full, pexp, or tukey')
    return exptaper


def winavg_hfilt(data, travel_time_us, avg_win, taper='full', filtdepth=100):
    """_RadarDataFiltering.py:353-440, the reference's own loop (it is O(T * avg_win * S), fine at test sizes)."""
    data = np.asarray(data)
    S, T = data.shape
    if avg_win > T:
        avg_win = T
    if avg_win % 2 == 0:
        avg_win = avg_win + 1
    exptaper = winavg_taper(travel_time_us, S, taper, filtdepth)
    out = np.zeros_like(data, dtype=data.dtype)
    for i in range(int(T)):
        range_start = max(i - ((avg_win - 1) // 2), 0)
        range_end = min(i + ((avg_win - 1) // 2), T)
        with np.errstate(invalid='ignore', divide='ignore'):
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                avg_trace = np.mean(data[:, range_start:range_end], axis=-1) * exptaper
        out[:, i] = data[:, i] - avg_trace
    return out


def horizontal_corner(wavelength, tracespace, dt, tnum):
    """Corner frequency of highpass / lowpass, _RadarDataFiltering.py:176-199 / :246-269."""
    wavelength = int(wavelength)
    nsamp = int(wavelength / tracespace)
    if nsamp < 1:
        raise ValueError('wavelength is too small, causing no samples per wavelength')
    if nsamp > tnum:
        raise ValueError('wavelength is too large, bigger than the whole radargram')
    return (100. / float(nsamp)) * 1.0e6 / ((1. / dt) / 2.0)


def highpass(data, wavelength, tracespace, dt):
    """_RadarDataFiltering.py:138-209: butter(5, 'high'), filtfilt along the last (trace) axis, float64 result."""
    from scipy.signal import butter, filtfilt
    b, a = butter(5, horizontal_corner(wavelength, tracespace, dt, data.shape[1]), 'high')
    return filtfilt(b, a, data)


def lowpass(data, wavelength, tracespace, dt):
    """_RadarDataFiltering.py:212-279: butter(3, 'low')."""
    from scipy.signal import butter, filtfilt
    b, a = butter(3, horizontal_corner(wavelength, tracespace, dt, data.shape[1]), 'low')
    return filtfilt(b, a, data)


def horizontal_band_pass(data, low, high, tracespace):
    """_RadarDataFiltering.py:282-350 (note: the Nyquist here is fsamp / 2 = 50, not 1 / (2 dt), :325)."""
    from scipy.signal import butter, filtfilt
    if low >= high:
        raise ValueError('Low must be less than high')
    if low <= 0.0:
        raise ValueError('Low must be larger than 0 but is {:f}'.format(low))
    nsamp_high = int(low / tracespace)
    nsamp_low = int(high / tracespace)
    if nsamp_high < 1:
        raise ValueError('Minimum wavelength is too small, causing no samples per wavelength')
    if nsamp_low > data.shape[1]:
        raise ValueError('Maximum wavelength is too long, causing more samples per wavelength than tnum, use lowpass instead?')
    corner = np.array([(100. / float(nsamp_low)) / 50., (100. / float(nsamp_high)) / 50.])
    b, a = butter(5, corner, 'bandpass')
    return filtfilt(b, a, data, axis=1)


def rangegain(data, travel_time_us, trig, slope):
    """_RadarDataProcessing.py:456-471."""
    data = np.array(data, copy=True)
    tt = np.asarray(travel_time_us)
    if isinstance(trig, (float, int, np.int64)):
        gain = tt[int(trig) + 1:] * slope
        data[int(trig + 1):, :] *= np.atleast_2d(gain).transpose()
    else:
        for i, tr in enumerate(trig):
            gain = tt[int(tr) + 1:] * slope
            data[int(tr) + 1:, i] *= gain
    return data


def agc(data, window=50, scaling_factor=50):
    """_RadarDataProcessing.py:474-488."""
    data = np.array(data, copy=True)
    S = data.shape[0]
    maxamp = np.zeros((S,))
    for i in range(S):
        maxamp[i] = np.max(np.abs(data[max(0, i - window // 2):min(i + window // 2, S), :]))
    maxamp[maxamp == 0] = 1.0e-6
    data *= (scaling_factor / np.atleast_2d(maxamp).transpose()).astype(data.dtype)
    return data


def wiener(data, vert_win=1, hor_win=10, noise=None):
    """denoise(ftype='wiener') (_RadarDataFiltering.py:573-582) = scipy.signal.wiener(data, (vert_win, hor_win), noise):
    zero padded box statistics over rows [s - V//2, s + (V-1)//2] and columns [t - H//2, t + (H-1)//2]
    (scipy.signal.correlate(..., 'same') with a ones window), local variance E[x^2] - mean^2 with the squares taken
    in the input precision, noise = mean local variance unless given; float64 result.  scipy picks an FFT correlation
    for these sizes, so the reference carries ~1e-16 relative rounding noise this direct summation does not."""
    x = np.asarray(data)
    if not np.issubdtype(x.dtype, np.floating):
        x = x.astype(np.float64)
    S, T = x.shape
    V, H = int(vert_win), int(hor_win)
    n = float(V * H)

    def boxsum(img):
        pad = np.zeros((S + V - 1, T + H - 1))
        pad[V // 2:V // 2 + S, H // 2:H // 2 + T] = img
        out = np.zeros((S, T))
        for dv in range(V):
            for dh in range(H):
                out += pad[dv:dv + S, dh:dh + T]
        return out

    mean = boxsum(x) / n
    var = boxsum(x ** 2) / n - mean ** 2
    if noise is None:
        noise = np.mean(var.reshape(-1))
        if np.any(var == 0) and noise != 0:
            raise ValueError('Could not compute variance, specify noise for denoise')
    with np.errstate(divide='ignore', invalid='ignore'):
        res = (x - mean)
        res = res * (1 - noise / var)
        res = res + mean
    return np.where(var < noise, mean, res)


def median_filter(data, vert_win=1, hor_win=10):
    """denoise(ftype='median') (_RadarDataFiltering.py:583-584) = scipy.ndimage.median_filter(data, size=(V, H)): rank
    (V*H)//2 of the window rows [s - V//2, s - V//2 + V), columns likewise, edges reflected with the edge sample
    repeated (numpy's 'symmetric' padding); the dtype is kept."""
    from numpy.lib.stride_tricks import sliding_window_view
    x = np.asarray(data)
    S, T = x.shape
    V, H = int(vert_win), int(hor_win)
    padded = np.pad(x, ((V // 2, V - 1 - V // 2), (H // 2, H - 1 - H // 2)), mode='symmetric')
    windows = sliding_window_view(padded, (V, H)).reshape(S, T, V * H)
    return np.sort(windows, axis=2)[:, :, (V * H) // 2]
