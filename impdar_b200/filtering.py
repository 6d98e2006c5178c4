"""Host side of the filtering hot path: drop-in mirrors of the RadarData filter methods
(RadarData/_RadarDataFiltering.py :19 adaptivehfilt, :93 horizontalfilt, :443 hfilt,
:469 vertical_band_pass, :590 migrate).  Each function takes the RadarData-like object as ``self`` so it
can be bound onto the reference's class (impdar_b200.install()) or onto impdar_b200.RadarData.

Filter design (scipy butter/cheby1/bessel/firwin, lfilter_zi) is O(order) host work; the data passes run
in libimpdar_b200.so.  float64 radargrams are filtered by the f64 kernels (so the reference's exact
known-answer test for horizontalfilt holds bit for bit), float32 ones by the f32 kernels; integer input
follows the reference's cast-back rules.  There is no CPU fallback.
"""
import numpy as np

from . import _lib, device
from . import migrationlib


def _exp_taper(self):
    """_RadarDataFiltering.py:59 / :129-130."""
    tt = np.asarray(self.travel_time, dtype=np.float64)
    return np.exp(-tt.flatten() * 0.05) / np.exp(-tt[0] * 0.05)


def _stage(data):
    """-> (cuda tensor in the compute dtype, suffix 'f32'|'f64', original numpy dtype or None, was_device)."""
    import torch
    if device.is_device_array(data):
        if data.dtype == torch.float64:
            return data.contiguous(), 'f64', None, True
        return device.to_device(data, torch.float32), 'f32', None, True
    a = np.asarray(data)
    if a.dtype == np.float32:
        return device.to_device(a, torch.float32), 'f32', a.dtype, False
    return device.to_device(a, torch.float64), 'f64', a.dtype, False


def _unstage(self, out, np_dtype, was_device):
    """Leave the result where the input lived; host results get the reference's dtype (``.astype(dtype)``,
    _RadarDataFiltering.py:85, :131, :529 - truncation for integer radargrams)."""
    import torch
    if was_device:
        self.data = out
        return
    host = device.to_host(out, np.float64 if out.dtype == torch.float64 else np.float32)
    if np_dtype is not None and host.dtype != np_dtype:
        host = host.astype(np_dtype)
    self.data = host


def horizontalfilt_device(x, suffix, taper, htr1, htrn, trunc_avg=False):
    import torch
    lib = _lib.load()
    S, T = x.shape[-2], x.shape[-1]
    B = 1 if x.dim() == 2 else x.shape[0]
    out = torch.empty_like(x)
    tp = device.to_device(taper, torch.float64)
    fn = getattr(lib, 'impdar_hfilt_' + suffix)
    _lib.check(fn(device.ptr(x), device.ptr(out), S, T, B, int(htr1), int(htrn), device.ptr(tp),
                  int(bool(trunc_avg)), device.current_stream_ptr()))
    return out


def horizontalfilt(self, ntr1, ntr2, *args, **kwargs):
    """Remove the (tapered) average trace; mirrors _RadarDataFiltering.py:93-135."""
    htr1 = int(max(0, min(ntr1, self.tnum - 1)))
    htrn = int(max(htr1 + 1, min(ntr2, self.tnum)))
    print('Subtracting mean trace found between {:d} and {:d}'.format(htr1, htrn))
    x, suffix, np_dtype, was_device = _stage(self.data)
    is_int = np_dtype is not None and np.issubdtype(np_dtype, np.integer)
    out = horizontalfilt_device(x, suffix, _exp_taper(self), htr1, htrn, trunc_avg=is_int)
    _unstage(self, out, np_dtype, was_device)
    print('Horizontal filter complete.')
    self.flags.hfilt = np.ones((2,))


def adaptivehfilt_device(x, suffix, taper, window_size):
    import torch
    lib = _lib.load()
    S, T = x.shape[-2], x.shape[-1]
    B = 1 if x.dim() == 2 else x.shape[0]
    out = torch.empty_like(x)
    tp = device.to_device(taper, torch.float64)
    nbytes = lib.impdar_ahfilt_workspace_bytes(S, T, B)
    ws = device.workspace(nbytes) if nbytes else None
    fn = getattr(lib, 'impdar_ahfilt_' + suffix)
    _lib.check(fn(device.ptr(x), device.ptr(out), S, T, B, int(window_size), device.ptr(tp), device.ptr(ws),
                  0 if ws is None else ws.numel(), device.current_stream_ptr()))
    return out


def adaptivehfilt(self, window_size, *args, **kwargs):
    """Moving-window average-trace removal; mirrors _RadarDataFiltering.py:19-90."""
    print('Adaptive filtering')
    x, suffix, np_dtype, was_device = _stage(self.data)
    out = adaptivehfilt_device(x, suffix, _exp_taper(self), int(window_size))
    _unstage(self, out, np_dtype, was_device)
    print('Adaptive filtering complete')
    self.flags.hfilt[0] = 1
    self.flags.hfilt[1] = 4


def hfilt(self, ftype='hfilt', bounds=None, window_size=None):
    """Dispatch wrapper; mirrors _RadarDataFiltering.py:443-466."""
    if ftype == 'hfilt':
        self.horizontalfilt(bounds[0], bounds[1])
    elif ftype == 'adaptive':
        self.adaptivehfilt(window_size=window_size)
    else:
        raise ValueError('Unrecognized filter type')


def iir_prepare(b, a):
    """Normalise (b, a) like scipy.signal.lfilter does and compute lfilter_zi; returns (b, a, zi, padlen)
    with padlen = 3 * max(len(a), len(b)), filtfilt's default."""
    from scipy.signal import lfilter_zi
    b = np.atleast_1d(np.asarray(b, dtype=np.float64))
    a = np.atleast_1d(np.asarray(a, dtype=np.float64))
    n = max(len(a), len(b))
    padlen = 3 * n
    bb = np.zeros(n)
    aa = np.zeros(n)
    bb[:len(b)] = b / a[0]
    aa[:len(a)] = a / a[0]
    zi = lfilter_zi(bb, aa)
    return bb, aa, np.ascontiguousarray(zi, dtype=np.float64), padlen


def filtfilt_device(x, suffix, b, a):
    """scipy.signal.filtfilt(b, a, x, axis=-2) on a (.., snum, tnum) CUDA tensor."""
    import torch
    lib = _lib.load()
    S, T = x.shape[-2], x.shape[-1]
    B = 1 if x.dim() == 2 else x.shape[0]
    bb, aa, zi, padlen = iir_prepare(b, a)
    if S <= padlen:
        raise ValueError("The length of the input vector x must be greater than padlen, which is %d." % padlen)
    out = torch.empty_like(x)
    nbytes = lib.impdar_filtfilt_workspace_bytes(S, T, B, padlen, x.element_size())
    ws = device.workspace(nbytes)
    fn = getattr(lib, 'impdar_filtfilt_' + suffix)
    _lib.check(fn(device.ptr(x), device.ptr(out), S, T, B, device.ptr(bb), device.ptr(aa), len(bb),
                  device.ptr(zi), padlen, device.ptr(ws), ws.numel(), device.current_stream_ptr()))
    return out


def fir_device(x, suffix, taps):
    import torch
    lib = _lib.load()
    S, T = x.shape[-2], x.shape[-1]
    B = 1 if x.dim() == 2 else x.shape[0]
    taps = np.ascontiguousarray(taps, dtype=np.float64)
    out = torch.empty_like(x)
    fn = getattr(lib, 'impdar_fir_' + suffix)
    _lib.check(fn(device.ptr(x), device.ptr(out), S, T, B, device.ptr(taps), len(taps),
                  device.current_stream_ptr()))
    return out


def vertical_band_pass(self, low, high, order=5, filttype='butter', cheb_rp=5, fir_window='hamming',
                       *args, **kwargs):
    """Forward-backward band-pass along time; mirrors _RadarDataFiltering.py:469-549."""
    from scipy.signal import butter, cheby1, bessel, firwin
    sample_freq = 1.0 / self.dt
    nyquist_freq = 0.5 * sample_freq
    corner_freq = np.zeros((2,))
    corner_freq[0] = low * 1.0e6 / nyquist_freq
    corner_freq[1] = high * 1.0e6 / nyquist_freq
    print('Bandpassing from {:4.1f} to {:4.1f} MHz...'.format(low, high))

    ft = filttype.lower()
    if ft in ['butter', 'butterworth']:
        b, a = butter(order, corner_freq, 'bandpass')
    elif ft in ['cheb', 'chebyshev']:
        b, a = cheby1(order, cheb_rp, corner_freq, 'bandpass')
    elif ft == 'bessel':
        b, a = bessel(order, corner_freq, 'bandpass')
    elif ft == 'fir':
        taps = firwin(order + 1, corner_freq, pass_zero=False)
    else:
        raise ValueError('Filter type {:s} is not recognized'.format(filttype))

    x, suffix, np_dtype, was_device = _stage(self.data)
    if ft == 'fir':
        # self.data[:-order, :] = lfilter(taps, 1.0, self.data, axis=0).astype(dtype)[order:, :]  (:539-540)
        y = fir_device(x, suffix, taps)
        out = x.clone()
        if order > 0:
            out[..., :-order, :] = y[..., order:, :]
        if np_dtype is not None and np.issubdtype(np_dtype, np.integer):
            out[..., :-order, :] = out[..., :-order, :].trunc()
    else:
        out = filtfilt_device(x, suffix, b, a)
    _unstage(self, out, np_dtype, was_device)
    print('Bandpass filter complete.')
    self.flags.bpass[0] = 1
    self.flags.bpass[1] = low
    self.flags.bpass[2] = high


# ------------------------------------------------------------------ sibling filters
try:  # the reference's own exception class when ImpDAR is importable, so `except ImpdarError` keeps working
    from impdar.lib.ImpdarError import ImpdarError
except Exception:  # pragma: no cover - ImpDAR is not installed on the GPU box
    class ImpdarError(Exception):
        """Used for exceptions caused by something radar-y (ImpdarError.py:12)."""


def filtfilt_rows_device(x, suffix, b, a):
    """scipy.signal.filtfilt(b, a, x, axis=-1) (along the trace axis) on a (.., snum, tnum) CUDA tensor."""
    import torch
    lib = _lib.load()
    S, T = x.shape[-2], x.shape[-1]
    B = 1 if x.dim() == 2 else x.shape[0]
    bb, aa, zi, padlen = iir_prepare(b, a)
    if T <= padlen:
        raise ValueError("The length of the input vector x must be greater than padlen, which is %d." % padlen)
    out = torch.empty_like(x)
    nbytes = lib.impdar_filtfilt_rows_workspace_bytes(S, T, B, padlen, x.element_size())
    ws = device.workspace(nbytes)
    fn = getattr(lib, 'impdar_filtfilt_rows_' + suffix)
    _lib.check(fn(device.ptr(x), device.ptr(out), S, T, B, device.ptr(bb), device.ptr(aa), len(bb),
                  device.ptr(zi), padlen, device.ptr(ws), ws.numel(), device.current_stream_ptr()))
    return out


def _horizontal_iir(self, b, a):
    """``self.data = filtfilt(b, a, self.data)`` (_RadarDataFiltering.py:203, :273, :338): the result is float64
    whatever the input dtype; a device-resident radargram stays on the device in its compute dtype."""
    x, suffix, np_dtype, was_device = _stage(self.data)
    out = filtfilt_rows_device(x, suffix, b, a)
    _unstage(self, out, None if was_device else np.float64, was_device)
    self.flags.hfilt = np.ones((2,))
    self.flags.hfilt[1] = 3


def _check_constant_spacing(self):
    if self.flags.interp is None or not self.flags.interp[0]:
        raise ImpdarError('This method can only be used on constantly spaced data')
    if self.flags.elev:
        raise ImpdarError('This will not work with elevation corrected data')


def _horizontal_corner(self, wavelength, label):
    """Shared arithmetic of highpass / lowpass (_RadarDataFiltering.py:170-199, :240-269)."""
    tracespace = self.flags.interp[1]
    wavelength = int(wavelength)
    fsamp = 100.
    nsamp = int(wavelength / tracespace)
    if nsamp < 1:
        raise ValueError('wavelength is too small, causing no samples per wavelength')
    if nsamp > self.tnum:
        raise ValueError('wavelength is too large, bigger than the whole radargram')
    print('Sample resolution = {:d}'.format(nsamp))
    high_corner_freq = fsamp / float(nsamp)
    print('{:s} cutoff at {:4.2f} MHz...'.format(label, high_corner_freq))
    nyquist_freq = (1. / self.dt) / 2.0
    return high_corner_freq * 1.0e6 / nyquist_freq


def highpass(self, wavelength):
    """High pass in the horizontal for a given wavelength; mirrors _RadarDataFiltering.py:138-209."""
    from scipy.signal import butter
    _check_constant_spacing(self)
    corner_freq = _horizontal_corner(self, wavelength, 'High')
    b, a = butter(5, corner_freq, 'high')
    _horizontal_iir(self, b, a)
    print('Highpass filter complete.')


def lowpass(self, wavelength):
    """Low pass in the horizontal for a given wavelength; mirrors _RadarDataFiltering.py:212-279."""
    from scipy.signal import butter
    _check_constant_spacing(self)
    corner_freq = _horizontal_corner(self, wavelength, 'Low')
    b, a = butter(3, corner_freq, 'low')
    _horizontal_iir(self, b, a)
    print('Lowpass filter complete.')


def horizontal_band_pass(self, low, high):
    """Bandpass in the horizontal for a pair of wavelengths; mirrors _RadarDataFiltering.py:282-350."""
    from scipy.signal import butter
    _check_constant_spacing(self)
    if low >= high:
        raise ValueError('Low must be less than high')
    if low <= 0.0:
        raise ValueError('Low must be larger than 0 but is {:f}'.format(low))
    tracespace = self.flags.interp[1]
    fsamp = 100.
    nsamp_high = int(low / tracespace)
    nsamp_low = int(high / tracespace)
    if nsamp_high < 1:
        raise ValueError('Minimum wavelength is too small, causing no samples per wavelength')
    if nsamp_low > self.tnum:
        raise ValueError('Maximum wavelength is too long, causing more samples per wavelength than tnum, use lowpass instead?')
    print('Sample resolution high = {:d}'.format(nsamp_high))
    print('Sample resolution low = {:d}'.format(nsamp_low))
    nyquist_freq = fsamp / 2.0
    corner_freq = np.zeros((2,))
    corner_freq[0] = (fsamp / float(nsamp_low)) / nyquist_freq
    corner_freq[1] = (fsamp / float(nsamp_high)) / nyquist_freq
    b, a = butter(5, corner_freq, 'bandpass')
    _horizontal_iir(self, b, a)
    print('Highpass filter complete.')


def winavg_device(x, suffix, taper, half):
    import torch
    lib = _lib.load()
    S, T = x.shape[-2], x.shape[-1]
    B = 1 if x.dim() == 2 else x.shape[0]
    out = torch.empty_like(x)
    tp = device.to_device(taper, torch.float64)
    nbytes = lib.impdar_winavg_workspace_bytes(S, T, B)
    ws = device.workspace(nbytes) if nbytes else None
    fn = getattr(lib, 'impdar_winavg_' + suffix)
    _lib.check(fn(device.ptr(x), device.ptr(out), S, T, B, int(half), device.ptr(tp), device.ptr(ws),
                  0 if ws is None else ws.numel(), device.current_stream_ptr()))
    return out


def winavg_hfilt(self, avg_win, taper='full', filtdepth=100):
    """Moving-window average-trace removal; mirrors _RadarDataFiltering.py:353-440."""
    if avg_win > self.tnum:
        print('Cannot average over more than the whole data matrix. Reducing avg_win to tnum')
        avg_win = self.tnum
    if avg_win % 2 == 0:
        avg_win = avg_win + 1
        print('The averaging window must be an odd number of traces.')
        print('The averaging window has been changed to {:d}'.format(avg_win))
    exptaper = _exp_taper(self)
    if taper == 'full':
        pass
    elif taper == 'pexp':
        exptaper[:filtdepth] = exptaper[:filtdepth] - exptaper[filtdepth]
        exptaper[filtdepth:self.snum] = 0
        exptaper = exptaper / np.max(exptaper)
    elif taper == 'tukey':
        raise NotImplementedError("the hard-coded StoDeep 'tukey' taper is marked unused in the reference "
                                  "(_RadarDataFiltering.py:410-415, pragma: no cover) and is not on the B200 path")
    else:
        raise ValueError('Unrecognized taper. Options are full, pexp, or tukey')
    x, suffix, np_dtype, was_device = _stage(self.data)
    out = winavg_device(x, suffix, exptaper, (int(avg_win) - 1) // 2)
    _unstage(self, out, np_dtype, was_device)
    self.flags.hfilt = np.zeros((2,))
    self.flags.hfilt[1] = 2
    print('Horizontal filter complete.')


def rowgain_device(x, suffix, gain, trig=None, in_double=True):
    """y[s, t] = x[s, t] * gain[s] for s > trig[t] (all rows when trig is None), in place."""
    import torch
    lib = _lib.load()
    S, T = x.shape[-2], x.shape[-1]
    B = 1 if x.dim() == 2 else x.shape[0]
    g = device.to_device(np.ascontiguousarray(gain, dtype=np.float64), torch.float64)
    tr = None if trig is None else torch.from_numpy(np.ascontiguousarray(trig, dtype=np.int32)).cuda()
    fn = getattr(lib, 'impdar_rowgain_' + suffix)
    _lib.check(fn(device.ptr(x), device.ptr(x), S, T, B, device.ptr(g), device.ptr(tr), int(bool(in_double)),
                  device.current_stream_ptr()))
    return x


def _reject_integer_gain(np_dtype):
    if np_dtype is not None and not np.issubdtype(np_dtype, np.floating):
        # the reference multiplies in place by a float array, which numpy refuses for integer radargrams
        raise TypeError("Cannot cast ufunc 'multiply' output from dtype('float64') to dtype('%s') "
                        "with casting rule 'same_kind'" % np_dtype)


def rangegain(self, slope):
    """Linear range gain below the trigger sample; mirrors _RadarDataProcessing.py:456-471."""
    x, suffix, np_dtype, was_device = _stage(self.data)
    _reject_integer_gain(np_dtype)
    tt = np.asarray(self.travel_time, dtype=np.float64).flatten()
    gain = np.ones(self.snum)
    if isinstance(self.trig, (float, int, np.int64)):
        t0 = int(self.trig) + 1
        # gain = travel_time[int(trig) + 1:] * slope, applied to data[int(trig + 1):, :]
        start = int(self.trig + 1)
        g = tt[t0:] * slope
        if len(g) != max(self.snum - start, 0) and len(g) != 1:
            raise ValueError('operands could not be broadcast together with shapes (%d,%d) (%d,1)'
                             % (max(self.snum - start, 0), self.tnum, len(g)))
        gain[start:] = g
        out = rowgain_device(x, suffix, gain, None, True)
    else:
        trig = np.asarray(self.trig).astype(int).flatten()
        # per-trace trigger: data[int(trig) + 1:, i] *= travel_time[int(trig) + 1:] * slope.  The slice start follows
        # Python's rules: trig + 1 < 0 counts from the end (negative triggers occur after crop(..., zero_trig=False)).
        # The kernel applies the gain to rows s > trig'[i]; trig' = start - 1 with the start resolved here.
        start = trig + 1
        start = np.where(start < 0, np.maximum(self.snum + start, 0), start)
        out = rowgain_device(x, suffix, tt * slope, (start - 1).astype(np.int32), True)
    _unstage(self, out, np_dtype, was_device)
    self.flags.rgain = True


def rowabsmax_device(x, suffix):
    import torch
    lib = _lib.load()
    S, T = x.shape[-2], x.shape[-1]
    B = 1 if x.dim() == 2 else x.shape[0]
    out = torch.empty((B, S) if x.dim() == 3 else (S,), dtype=torch.float64, device=x.device)
    fn = getattr(lib, 'impdar_rowabsmax_' + suffix)
    _lib.check(fn(device.ptr(x), device.ptr(out), S, T, B, device.current_stream_ptr()))
    return out


def agc(self, window=50, scaling_factor=50):
    """Automatic gain control; mirrors _RadarDataProcessing.py:474-488.  The per-row max |data| is reduced on the
    device; the (snum,) window maximum and the scale vector are O(snum * window) host work."""
    x, suffix, np_dtype, was_device = _stage(self.data)
    rowmax = device.to_host(rowabsmax_device(x, suffix), np.float64)
    maxamp = np.zeros((self.snum,))
    for i in range(self.snum):
        maxamp[i] = np.max(rowmax[max(0, i - window // 2):min(i + window // 2, self.snum)])
    maxamp[maxamp == 0] = 1.0e-6
    dtype = np_dtype if np_dtype is not None else (np.float32 if suffix == 'f32' else np.float64)
    scale = (scaling_factor / maxamp).astype(dtype).astype(np.float64)
    out = rowgain_device(x, suffix, scale, None, False)
    _unstage(self, out, np_dtype, was_device)
    self.flags.agc = True


def wiener_device(x, suffix, vert_win, hor_win, noise=None):
    """scipy.signal.wiener on a (snum, tnum) CUDA tensor -> (float64 result, scratch tensor [sum of lVar, zero flag])."""
    import torch
    lib = _lib.load()
    S, T = int(x.shape[0]), int(x.shape[1])
    out = torch.empty((S, T), dtype=torch.float64, device=x.device)
    scratch = torch.zeros(2, dtype=torch.float64, device=x.device)
    fn = getattr(lib, 'impdar_wiener_' + suffix)
    _lib.check(fn(device.ptr(x), device.ptr(out), S, T, int(vert_win), int(hor_win), int(noise is None),
                  0.0 if noise is None else float(noise), device.ptr(scratch), device.current_stream_ptr()))
    return out, scratch


def denoise(self, vert_win=1, hor_win=10, noise=None, ftype='wiener'):
    """Denoising filter; mirrors _RadarDataFiltering.py:552-587.  The Wiener filter runs on the device in float64
    (result float64 like scipy's); with noise=None a local variance of exactly zero raises the reference's ValueError."""
    if ftype == 'wiener':
        import torch
        x, suffix, np_dtype, was_device = _stage(self.data)
        out, scratch = wiener_device(x, suffix, vert_win, hor_win, noise)
        if noise is None:
            flag = int(scratch.view(torch.int32)[2].item())           # the int at byte 8
            total = float(scratch[0].item())
            if (flag & 1) and total != 0.0:                           # noise / 0 with noise != 0: numpy's 'divide' error
                raise ValueError('Could not compute variance, specify noise for denoise')
        if was_device:
            # scipy's wiener returns float64; inside process()'s chain (_b200_reference_dtypes) the float64 result is
            # kept like restack / nmo do, a bare float32 device tensor keeps its dtype
            keep64 = x.dtype == torch.float64 or getattr(self, '_b200_reference_dtypes', False)
            self.data = out if keep64 else out.float()
        else:
            self.data = device.to_host(out, np.float64)
    elif ftype == 'median':
        # scipy.ndimage.median_filter(data, size=(vert_win, hor_win)): pure selection, the dtype is kept
        import torch
        x, suffix, np_dtype, was_device = _stage(self.data)
        if x.dim() != 2:
            raise ValueError('denoise expects a (snum, tnum) radargram')
        out = torch.empty_like(x)
        fn = getattr(_lib.load(), 'impdar_median_' + suffix)
        _lib.check(fn(device.ptr(x), device.ptr(out), int(x.shape[0]), int(x.shape[1]), int(vert_win), int(hor_win),
                      device.current_stream_ptr()))
        _unstage(self, out, np_dtype, was_device)
    else:
        raise ValueError('Only the wiener filter has been implemented for denoising.')


def migrate(self, mtype='stolt', vtaper=10, htaper=10, tmig=0, vel_fn=None, vel=1.68e8, nxpad=10,
            nearfield=False, verbose=0):
    """Dispatch on mtype exactly like _RadarDataFiltering.py:590-637; the callables are looked up on
    impdar_b200.migrationlib at call time, so they stay patchable like the reference's seam."""
    if mtype == 'kirch':
        migrationlib.migrationKirchhoff(self, vel=vel, nearfield=nearfield)
    elif mtype == 'stolt':
        migrationlib.migrationStolt(self, vel=vel, htaper=htaper, vtaper=vtaper)
    elif mtype == 'phsh':
        migrationlib.migrationPhaseShift(self, vel=vel, vel_fn=vel_fn, htaper=htaper, vtaper=vtaper)
    elif mtype == 'tk':
        migrationlib.migrationTimeWavenumber(self, vel=vel, vel_fn=vel_fn, htaper=htaper, vtaper=vtaper)
    elif mtype[:2] == 'su':
        raise NotImplementedError('SeisUnix migrations shell out to external binaries (mig_su.py); '
                                  'they are outside the B200 hot path - use the reference for mtype=%s' % mtype)
    else:
        raise ValueError('Unrecognized migration routine')
    self.flags.mig = mtype
