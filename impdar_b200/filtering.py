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
