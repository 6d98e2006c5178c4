"""Host side of the index / resampling operations either side of the hot path: drop-in
mirrors of RadarData/_RadarDataProcessing.py :20 reverse, :50 constant_sample_depth_spacing, :66 nmo,
:191 traveltime_to_depth, :238 crop, :352 hcrop, :424 restack, :499 constant_space, :587 elev_correct.

Every function takes the RadarData-like object as ``self`` (bound onto ImpDAR's class by
``impdar_b200.install()`` or onto ``impdar_b200.RadarData``).  The O(snum) / O(tnum) bookkeeping (travel_time,
nmo_depth, dist, GPS vectors, flags, picks) is float64 numpy on the host, as in the reference; every pass over
the (snum, tnum) radargram runs in libimpdar_b200.so (csrc/indexops.cu) and is bit-exact against the reference
(numpy 2.3 / scipy 1.18 arithmetic order).  With ``dat.data`` a CUDA tensor the radargram never leaves the GPU,
so a chain  vbp -> hfilt -> nmo -> crop -> migrate  needs one upload and one download.  There is no CPU
fallback for the radargram passes.
"""
import numpy as np

from . import _lib, device

try:  # raise ImpDAR's own exception class when ImpDAR is installed (isinstance checks in user code keep working)
    from impdar.lib.ImpdarError import ImpdarError
except Exception:  # noqa: BLE001 - ImpDAR absent (GPU box) or not importable (matplotlib / h5py missing)
    class ImpdarError(Exception):
        """Used for exceptions caused by something radar-y (lib/ImpdarError.py)."""


def firn_permittivity(rhof, rhoi=917., epsi_real=3.12, epsi_imag=-9.5):
    """DECOMP mixing model, Wilhelms (2005); lib/permittivity_models.py:46-71."""
    cube_root = (epsi_real - 1j * epsi_imag) ** (1 / 3.)
    return (1. + (rhof / rhoi) * (cube_root - 1)) ** 3.


_NODE_DTYPE = np.dtype([('lo', '<i4'), ('hi', '<i4'), ('a', '<f8'), ('b', '<f8'), ('den', '<f8'),
                        ('exact', '<i4'), ('pad', '<i4')])


# ----------------------------------------------------------------------------------------------- staging
def _stage_any(data):
    """-> (contiguous CUDA tensor, numpy dtype of a host input or None for the device lane, raw).  raw: the tensor
    is a uint8 byte carrier of shape (snum, tnum, itemsize) - any numpy dtype travels, torch need not know it."""
    import torch
    if device.is_device_array(data):
        return data.contiguous(), None, False
    device.require_cuda()
    a = np.ascontiguousarray(np.asarray(data))
    if a.dtype.itemsize not in (1, 2, 4, 8, 16):
        raise TypeError('radargram dtype %s is not supported' % a.dtype)
    raw = torch.from_numpy(a.view(np.uint8).reshape(a.shape + (a.dtype.itemsize,)))
    return raw.cuda(non_blocking=True), a.dtype, True


def _stage_float(data):
    """-> (CUDA tensor f32|f64, 'f32'|'f64', host numpy dtype or None).  Host float32 stays float32 (the reference
    keeps it through np.mean / interp1d); everything else that is not floating is cast to float64 first, which is
    what np.mean(dtype=None) and scipy's interp1d do with integer input."""
    import torch
    if device.is_device_array(data):
        if data.dtype == torch.float64:
            return data.contiguous(), 'f64', None
        return device.to_device(data, torch.float32), 'f32', None
    a = np.asarray(data)
    if np.iscomplexobj(a):
        raise NotImplementedError('complex radargrams are outside the B200 hot path for this operation')
    if a.dtype == np.float32:
        return device.to_device(a, torch.float32), 'f32', a.dtype
    return device.to_device(a, torch.float64), 'f64', a.dtype


def _out_kind(self, suffix, host_dtype):
    """Kernel suffix and torch dtype of the result: the reference's float64 on the host lane, the tensor's own
    dtype on the device lane (unless the object asks for the reference's dtypes: impdar_b200.process sets
    ``_b200_reference_dtypes`` while it runs a chain whose result goes back to the host)."""
    import torch
    if host_dtype is None and not getattr(self, '_b200_reference_dtypes', False):
        return suffix, (torch.float32 if suffix == 'f32' else torch.float64)
    return ('f32_f64' if suffix == 'f32' else 'f64'), torch.float64


def _finish(self, out, host_dtype, np_dtype=np.float64):
    if host_dtype is None:
        self.data = out
    else:
        self.data = device.to_host(out, np_dtype)


# ------------------------------------------------------------------------------------------- device passes
def crop_device(x, r0, r1, c0, c1, flip_lr=False, elem_bytes=None):
    """x[..., r0:r1, c0:c1] (np.fliplr'ed if flip_lr) as a new contiguous tensor; any element size."""
    import torch
    lib = _lib.load()
    if elem_bytes is None:
        elem_bytes = x.element_size()
        shape = x.shape
    else:                       # raw byte carrier (..., S, T, elem_bytes)
        shape = x.shape[:-1]
    S, T = int(shape[-2]), int(shape[-1])
    B = 1 if len(shape) == 2 else int(shape[0])
    out_shape = tuple(shape[:-2]) + (r1 - r0, c1 - c0)
    if x.shape != shape:
        out_shape = out_shape + (elem_bytes,)
    out = torch.empty(out_shape, dtype=x.dtype, device=x.device)
    _lib.check(lib.impdar_crop_bytes(device.ptr(x), device.ptr(out), S, T, B, int(r0), int(r1), int(c0), int(c1),
                                     int(bool(flip_lr)), int(elem_bytes), device.current_stream_ptr()))
    return out


def _crop_any(self, r0, r1, c0, c1, flip_lr=False):
    """Block copy of dat.data for any dtype, result left where the input lived (host results keep the dtype)."""
    x, host_dtype, raw = _stage_any(self.data)
    out = crop_device(x, r0, r1, c0, c1, flip_lr, elem_bytes=host_dtype.itemsize if raw else None)
    if host_dtype is None:
        self.data = out
        return
    import torch
    host = torch.empty(out.shape, dtype=out.dtype, pin_memory=True)
    host.copy_(out, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    arr = host.numpy()
    if raw:
        arr = arr.view(host_dtype).reshape(arr.shape[:-1])
    self.data = arr


def shift_traces_device(x, suffix_out, out_dtype, shift, snum_out):
    import torch
    lib = _lib.load()
    S, T = int(x.shape[0]), int(x.shape[1])
    out = torch.empty((int(snum_out), T), dtype=out_dtype, device=x.device)
    sh = torch.from_numpy(np.ascontiguousarray(shift, dtype=np.int32)).to(x.device, non_blocking=True)
    fn = getattr(lib, 'impdar_shift_traces_' + suffix_out)
    _lib.check(fn(device.ptr(x), device.ptr(out), S, T, int(snum_out), device.ptr(sh), device.current_stream_ptr()))
    return out


def restack_device(x, suffix_out, out_dtype, traces):
    import torch
    lib = _lib.load()
    S, T = int(x.shape[0]), int(x.shape[1])
    out = torch.empty((S, T // int(traces)), dtype=out_dtype, device=x.device)
    fn = getattr(lib, 'impdar_restack_' + suffix_out)
    _lib.check(fn(device.ptr(x), device.ptr(out), S, T, int(traces), device.current_stream_ptr()))
    return out


def _nodes_to_device(nodes, dev):
    import torch
    assert nodes.dtype == _NODE_DTYPE and nodes.dtype.itemsize == _lib.load().impdar_interp_node_bytes()
    return torch.from_numpy(np.ascontiguousarray(nodes).view(np.uint8)).to(dev, non_blocking=True)


def interp_rows_device(x, suffix_out, out_dtype, nodes, mode):
    import torch
    lib = _lib.load()
    S, T = int(x.shape[0]), int(x.shape[1])
    out = torch.empty((len(nodes), T), dtype=out_dtype, device=x.device)
    nd = _nodes_to_device(nodes, x.device)
    fn = getattr(lib, 'impdar_interp_rows_' + suffix_out)
    _lib.check(fn(device.ptr(x), device.ptr(out), S, T, len(nodes), device.ptr(nd), int(mode),
                  device.current_stream_ptr()))
    return out


def interp_cols_device(x, suffix_out, out_dtype, nodes, mode):
    import torch
    lib = _lib.load()
    S, T = int(x.shape[0]), int(x.shape[1])
    out = torch.empty((S, len(nodes)), dtype=out_dtype, device=x.device)
    nd = _nodes_to_device(nodes, x.device)
    fn = getattr(lib, 'impdar_interp_cols_' + suffix_out)
    _lib.check(fn(device.ptr(x), device.ptr(out), S, T, len(nodes), device.ptr(nd), int(mode),
                  device.current_stream_ptr()))
    return out


# --------------------------------------------------------------------- interpolation tables (host, float64)
def _check_interp_bounds(x, x_new):
    """scipy.interpolate.interp1d(bounds_error=True)._check_bounds: ValueError outside [x[0], x[-1]]."""
    if len(x_new) and np.any(x_new < x[0]):
        raise ValueError("A value ({}) in x_new is below the interpolation range's minimum value ({})."
                         .format(x_new[np.argmax(x_new < x[0])], x[0]))
    if len(x_new) and np.any(x_new > x[-1]):
        raise ValueError("A value ({}) in x_new is above the interpolation range's maximum value ({})."
                         .format(x_new[np.argmax(x_new > x[-1])], x[-1]))


def _interp1d_prepare(x):
    """interp1d's constructor checks for the abscissa (assume_sorted=False: it sorts; the reference's x are
    ascending - nmotime, nmo_depth, cumulative dist - and unsorted input would permute the rows as well, so it
    is rejected here instead)."""
    x = np.asarray(x, dtype=np.float64)
    if x.ndim != 1:
        raise ValueError("the x array must have exactly one dimension.")
    if len(x) < 2:
        raise ValueError("x and y arrays must have at least 2 entries")
    if np.any(np.diff(x) < 0):
        raise NotImplementedError('interpolation abscissa must be ascending')
    return x


def linear_nodes_scipy(x, x_new):
    """Node table of scipy's interp1d._call_linear (mode 0): searchsorted(side='left') clipped to [1, n-1],
    w_hi = (x_new - x_lo) / (x_hi - x_lo), w_lo = (x_hi - x_new) / (x_hi - x_lo)."""
    x = _interp1d_prepare(x)
    x_new = np.asarray(x_new, dtype=np.float64)
    _check_interp_bounds(x, x_new)
    hi = np.searchsorted(x, x_new).clip(1, len(x) - 1).astype(int)
    lo = hi - 1
    x_lo, x_hi = x[lo], x[hi]
    nodes = np.zeros(len(x_new), dtype=_NODE_DTYPE)
    nodes['lo'], nodes['hi'] = lo, hi
    with np.errstate(divide='ignore', invalid='ignore'):
        nodes['a'] = (x_new - x_lo) / (x_hi - x_lo)
        nodes['b'] = (x_hi - x_new) / (x_hi - x_lo)
    return nodes


def linear_nodes_numpy(xp, x_new):
    """Node table of numpy.interp (mode 1; numpy/_core/src/multiarray/compiled_base.c arr_interp): j = the last
    node with xp[j] <= x; x == xp[j] (or the last node) copies y[j]; otherwise slope form from the left node."""
    xp = _interp1d_prepare(xp)
    x_new = np.asarray(x_new, dtype=np.float64)
    _check_interp_bounds(xp, x_new)
    n = len(xp)
    j = np.searchsorted(xp, x_new, side='right') - 1
    last = j >= n - 1
    j = np.clip(j, 0, n - 2)
    nodes = np.zeros(len(x_new), dtype=_NODE_DTYPE)
    nodes['lo'] = np.where(last, n - 1, j)
    nodes['hi'] = np.where(last, n - 1, j + 1)
    nodes['a'] = x_new - xp[j]
    nodes['b'] = x_new - xp[j + 1]
    nodes['den'] = xp[j + 1] - xp[j]
    nodes['exact'] = (last | (x_new == xp[j])).astype(np.int32)
    return nodes


def _interp1d_vector(x, y, x_new):
    """interp1d(x, y)(x_new) for the 1-D per-trace vectors; host float64, through scipy itself."""
    from scipy.interpolate import interp1d
    return interp1d(x, y)(x_new)


# ------------------------------------------------------------------------------------------ the methods
_PER_TRACE = ('lat', 'long', 'pressure', 'trace_int', 'trig', 'elev', 'x_coord', 'y_coord', 'decday')


def _picks(self):
    return getattr(self, 'picks', None)


def _set_flag_pair(flags, name, value):
    """flags.<name> = [1, value]; flag vectors that came back malformed from old .mat files are rebuilt
    (_RadarDataProcessing.py:182-187, :578-584)."""
    try:
        vec = getattr(flags, name)
        vec[0] = 1
        vec[1] = value
    except (IndexError, TypeError):
        setattr(flags, name, np.array([1., value]))


def _note_vertical_crop(self, first, last):
    """flags.crop = [1, first row kept, one past the last row kept] relative to the original file (:338-349)."""
    try:
        base = self.flags.crop[1]
        self.flags.crop[0] = 1
        self.flags.crop[2] = base + last
    except (IndexError, TypeError):
        self.flags.crop = np.array([1., 0., float(last)])
    self.flags.crop[1] = self.flags.crop[1] + first
    print('Vertical samples reduced to subset [{:d}:{:d}] of original'.format(int(self.flags.crop[1]),
                                                                             int(self.flags.crop[2])))


def reverse(self):
    """Flip the profile left-right; mirrors _RadarDataProcessing.py:20-47."""
    nrow, ncol = (int(n) for n in self.data.shape[:2])
    _crop_any(self, 0, nrow, 0, ncol, flip_lr=True)
    for name in ('x_coord', 'y_coord', 'decday', 'lat', 'long', 'elev'):
        vec = getattr(self, name, None)
        if vec is not None:       # impdar_b200.RadarData leaves absent GPS vectors as None
            setattr(self, name, np.flip(vec, 0))
    if _picks(self) is not None:
        self.picks.reverse()
    was_reversed = bool(self.flags.reverse)
    print('Back to original direction' if was_reversed else 'Profile direction reversed')
    self.flags.reverse = not was_reversed


def constant_sample_depth_spacing(self):
    """Resample rows to constant depth spacing; mirrors _RadarDataProcessing.py:50-63."""
    old = self.nmo_depth
    if old is None:
        raise AttributeError('Call nmo first...')
    if np.allclose(np.diff(old), np.ones((self.snum - 1,)) * (old[1] - old[0])):
        print('No constant sampling when you already have constant sampling...')
        return 1
    new = np.linspace(np.min(old[0], 0), old[-1], len(old))
    x, suffix, host_dtype = _stage_float(self.data)
    kernel_suffix, out_dtype = _out_kind(self, suffix, host_dtype)
    _finish(self, interp_rows_device(x, kernel_suffix, out_dtype, linear_nodes_scipy(old, new), 0), host_dtype)
    self.travel_time = _interp1d_vector(old, self.travel_time, new)
    self.nmo_depth = new


def traveltime_to_depth(self, profile_depth, profile_rho, c=3.0e8, permittivity_model=firn_permittivity):
    """Depth of every sample for a density profile; mirrors _RadarDataProcessing.py:191-235 (O(snum) host)."""
    speed = c / np.sqrt(np.real(permittivity_model(profile_rho)))
    depth = self.travel_time / 2. * c / np.sqrt(np.real(permittivity_model(917.))) * 1.0e-6
    z = 0.
    step_us = self.dt * 1.0e6
    for k, t in enumerate(self.travel_time):
        if t < 0.:
            continue                                    # pre-trigger samples keep the solid-ice estimate
        if t < step_us:
            z += t / 2. * speed[0] * 1.0e-6
        else:
            z += self.dt / 2. * speed[np.nanargmin(abs(profile_depth - z))]
        depth[k] = z
    return depth


def _nmo_times(self, ant_sep, uice, u_interp=None, d_interp=None):
    """Vertical two-way time of every sample (_RadarDataProcessing.py:133-160), float64 host.  With a firn profile the
    RMS velocity above the reflector is iterated to a fixed point exactly like the reference (at least five rounds,
    np.mean of the squared speeds above the current depth guess)."""
    out = np.zeros((len(self.travel_time)))
    tol = 0.1 * self.dt / 2. * uice
    for k, t in enumerate(self.travel_time):
        u_rms = uice
        if u_interp is not None:
            guess = t / 2. * uice * 1.0e-6
            previous = guess.copy()
            rounds = 0
            while abs(guess - previous) > tol or rounds < 5:
                previous = guess.copy()
                u_rms = np.sqrt(np.mean(u_interp[d_interp <= guess] ** 2.))
                guess = t / 2. * u_rms * 1.0e-6
                rounds += 1
        t_sep = 1e6 * (ant_sep / u_rms)                 # direct arrival across the antenna separation [us]
        out[k] = np.sqrt((t + t_sep) ** 2. - t_sep ** 2.)
    return out


def nmo(self, ant_sep, uice=1.69e8, uair=3.0e8, const_firn_offset=None, rho_profile=None,
        permittivity_model=firn_permittivity, const_sample=False):
    """Normal move-out correction; mirrors _RadarDataProcessing.py:66-188.  The per-trace
    ``interp1d(nmotime, trace)(travel_time)`` loop is one row-interpolation pass on the device."""
    if np.any(self.trig > 0):
        raise ImpdarError('Crop out the pretrigger before doing the nmo correction.')

    firn = None
    if rho_profile is not None:
        try:
            table = np.genfromtxt(rho_profile, delimiter=',')
            firn = (table[:, 0], table[:, 1])
        except IndexError:
            raise IndexError('Cannot load the depth-density profile')
        speed = uair / np.sqrt(np.real(permittivity_model(firn[1])))
        d_interp = np.linspace(np.min(firn[0], 0), max(firn[0]), 10 * self.snum)
        u_interp = _interp1d_vector(firn[0], speed, d_interp)
        print('Iterating velocity profile in firn...')
        nmotime = _nmo_times(self, ant_sep, uice, u_interp, d_interp)
    else:
        nmotime = _nmo_times(self, ant_sep, uice)
    new_tt = np.arange(min(self.travel_time), max(nmotime), self.dt * 1e6)

    x, suffix, host_dtype = _stage_float(self.data)
    # scipy hands a 1-D float64 trace to numpy.interp and everything else to its own two-weight formula
    if suffix == 'f64':
        nodes, mode = linear_nodes_numpy(nmotime, new_tt), 1
    else:
        nodes, mode = linear_nodes_scipy(nmotime, new_tt), 0
    kernel_suffix, out_dtype = _out_kind(self, suffix, host_dtype)
    resampled = interp_rows_device(x, kernel_suffix, out_dtype, nodes, mode)
    self.travel_time = new_tt
    self.snum = len(new_tt)
    _finish(self, resampled, host_dtype)

    if firn is None:
        self.nmo_depth = self.travel_time / 2. * uice * 1.0e-6
    else:
        self.nmo_depth = traveltime_to_depth(self, firn[0], firn[1], c=uair, permittivity_model=permittivity_model)
    if const_sample:
        constant_sample_depth_spacing(self)
    if const_firn_offset is not None:
        self.nmo_depth = self.nmo_depth + const_firn_offset
    print('Normal Moveout filter complete.')
    _set_flag_pair(self.flags, 'nmo', ant_sep)


def _crop_index(self, lim, dimension, uice):
    """First sample at or beyond `lim` in the chosen unit (_RadarDataProcessing.py:271-288)."""
    if dimension == 'twtt':
        return np.min(np.argwhere(self.travel_time >= lim))
    if dimension == 'depth':
        depth = getattr(self, 'nmo_depth', None)
        if depth is None:
            depth = self.travel_time / 2. * uice * 1.0e-6
        return np.min(np.argwhere(depth >= lim))
    if dimension == 'pretrig':
        return self.trig.astype(int) if isinstance(self.trig, np.ndarray) else int(self.trig)
    return int(lim)


def crop(self, lim, top_or_bottom='top', dimension='snum', uice=1.69e8, rezero=True, zero_trig=True):
    """Crop in the vertical; mirrors _RadarDataProcessing.py:238-349."""
    if top_or_bottom not in ['top', 'bottom']:
        raise ValueError('top_or_bottom must be "top" or "bottom" not {:s}'.format(top_or_bottom))
    if dimension not in ['snum', 'twtt', 'depth', 'pretrig']:
        raise ValueError('Dimension must be in [\'snum\', \'twtt\', \'depth\']')
    if top_or_bottom == 'bottom' and dimension == 'pretrig':
        raise ValueError('Only use pretrig to crop from the top')

    ind = _crop_index(self, lim, dimension, uice)
    nrow, ncol = (int(n) for n in self.data.shape[:2])
    per_trace = isinstance(ind, np.ndarray) and dimension == 'pretrig'
    if not per_trace:
        # one limit for the whole profile: a row block
        if top_or_bottom == 'top':
            first, last = ind, nrow
            self.trig = np.zeros_like(self.trig) if zero_trig else self.trig - ind
        else:
            first, last = 0, ind
        keep = slice(first, last)
        r0, r1, _ = keep.indices(nrow)                  # Python slice semantics (negative / oversize limits)
        _crop_any(self, r0, max(r0, r1), 0, ncol)
        if getattr(self, 'nmo_depth', None) is not None:
            self.nmo_depth = self.nmo_depth[keep]
    else:
        # pretrigger given per trace: every trace moves up by its own trigger sample, NaN below
        ind = np.asarray(ind)
        if ind.shape != (ncol,):
            raise ValueError('trig must have one entry per trace')
        first, last = np.nanmin(ind), nrow
        if first < 0:
            raise ValueError('could not broadcast input array: negative trigger samples cannot be cropped')
        keep = slice(first, last)
        self.trig = self.trig - ind
        x, suffix, host_dtype = _stage_float(self.data)
        kernel_suffix, out_dtype = _out_kind(self, suffix, host_dtype)
        _finish(self, shift_traces_device(x, kernel_suffix, out_dtype, ind, nrow - int(first)), host_dtype)
    tt = self.travel_time[keep]
    self.travel_time = tt - tt[0] if rezero else tt
    self.snum = self.data.shape[0]

    if top_or_bottom == 'top' and _picks(self) is not None:
        self.picks.crop(ind)
    _note_vertical_crop(self, first, last)


def _hcrop_index(self, lim, dimension):
    """Zero-based trace index of the crop limit with the reference's validity rules (:377-392)."""
    if dimension == 'dist':
        if lim > np.max(self.dist):
            raise ValueError('lim is larger than largest distance')
        if lim <= 0:
            raise ValueError('Distance should be strictly positive')
        return np.min(np.argwhere(self.dist >= lim))
    if int(lim) in (0, 1):
        raise ValueError('lim should be at least two to preserve some data')
    if lim > self.tnum:
        raise ValueError('lim should be less than tnum+1 {:d} in order to do anything'.format(self.tnum + 1))
    if lim == -1 or lim < -int(self.tnum):
        raise ValueError('If negative, lim should be in [-self.tnum; -1)')
    return int(lim) - 1                                 # trace numbers are 1-indexed


def hcrop(self, lim, left_or_right='left', dimension='tnum'):
    """Crop in the horizontal; mirrors _RadarDataProcessing.py:352-421."""
    if left_or_right not in ['left', 'right']:
        raise ValueError('left_or_right must be left or right, not {:s}'.format(left_or_right))
    if dimension not in ['tnum', 'dist']:
        raise ValueError('Dimension must be in ["tnum", "dist"]')
    ind = _hcrop_index(self, lim, dimension)
    nrow, ncol = (int(n) for n in self.data.shape[:2])
    first, last = (ind, ncol) if left_or_right == 'left' else (0, ind)
    keep = slice(first, last)
    c0, c1, _ = keep.indices(ncol)
    _crop_any(self, 0, nrow, c0, max(c0, c1))
    for name in _PER_TRACE:                             # scalars (a float trig, a scalar trace_int) stay as they are
        vec = getattr(self, name, None)
        if isinstance(vec, np.ndarray):
            setattr(self, name, vec[keep])
    if _picks(self) is not None:
        self.picks.hcrop([first, last])
    if self.dist is not None:
        self.dist = self.dist[keep] - self.dist[first]
    if getattr(self, 'trace_num', None) is not None:
        self.trace_num = self.trace_num[keep] - first + 1
    self.tnum = self.data.shape[1]


_RESTACK_VARS = ('dist', 'pressure', 'lat', 'long', 'x_coord', 'y_coord', 'elev', 'decday', 'trig')


def restack(self, traces):
    """Average groups of `traces` adjacent traces; mirrors _RadarDataProcessing.py:424-477."""
    traces = int(traces)
    if traces % 2 == 0:
        print('Only will stack odd numbers of traces. Using {:d}'.format(int(traces + 1)))
        traces = traces + 1
    groups = int(np.floor(self.tnum / traces))
    x, suffix, host_dtype = _stage_float(self.data)
    kernel_suffix, out_dtype = _out_kind(self, suffix, host_dtype)
    stacked = restack_device(x, kernel_suffix, out_dtype, traces)

    width = int(x.shape[1])
    averaged = {}
    for name in _RESTACK_VARS:
        vec = getattr(self, name, None)
        averaged[name] = None if vec is None else np.array(
            [np.mean(vec[g * traces:min((g + 1) * traces, width)]) for g in range(groups)], dtype=np.float64).reshape(groups)
    self.tnum = groups
    _finish(self, stacked, host_dtype)
    self.trace_num = np.arange(groups).astype(int) + 1
    self.trace_int = np.zeros((groups, ))                # the reference leaves the restacked spacing at zero (:452, :468)
    if _picks(self) is not None:
        self.picks.restack(traces)
    for name, vec in averaged.items():
        setattr(self, name, vec)
    self.flags.restack = True


def constant_space(self, spacing, min_movement=1.0e-2, show_nomove=False):
    """Resample to constant trace spacing; mirrors _RadarDataProcessing.py:499-584: stationary traces are dropped
    (column compaction) and the rest interpolated linearly in distance - one column-gather pass on the device."""
    moving = np.hstack((np.array([True]), np.diff(self.dist * 1000.) >= min_movement))
    for k in np.flatnonzero(~moving):                   # close the gaps the dropped traces leave in the distance axis
        self.dist[k:] = self.dist[k:] - (self.dist[k] - self.dist[k - 1])
    old = self.dist[moving]
    new = np.arange(np.min(old), np.max(old), step=spacing / 1000.0)

    nodes = linear_nodes_scipy(old, new)
    kept = np.flatnonzero(moving)
    nodes['lo'], nodes['hi'] = kept[nodes['lo']], kept[nodes['hi']]      # compaction folded into the gather

    if (not device.is_device_array(self.data)) and np.iscomplexobj(np.asarray(self.data)):
        # real weights times complex samples act on the two components separately: run the float64 kernel on the
        # interleaved (snum, 2 tnum) view with every node duplicated for the real and the imaginary column
        import torch
        z = np.ascontiguousarray(np.asarray(self.data), dtype=np.complex128)
        both = np.repeat(nodes, 2)
        part = np.tile([0, 1], len(nodes))
        both['lo'], both['hi'] = 2 * both['lo'] + part, 2 * both['hi'] + part
        x = device.to_device(z.view(np.float64), torch.float64)
        self.data = device.to_host(interp_cols_device(x, 'f64', torch.float64, both, 0), np.float64).view(np.complex128)
    else:
        x, suffix, host_dtype = _stage_float(self.data)
        kernel_suffix, out_dtype = _out_kind(self, suffix, host_dtype)
        _finish(self, interp_cols_device(x, kernel_suffix, out_dtype, nodes, 0), host_dtype)

    def onto_new(values):
        return _interp1d_vector(old, values, new)

    for name in ('lat', 'long', 'x_coord', 'y_coord', 'decday', 'pressure', 'trig', 'elev'):
        vec = getattr(self, name)
        if name == 'elev' and vec is None:              # elev is the one optional vector (:562-566)
            continue
        setattr(self, name, onto_new(vec[moving]))
    picks = _picks(self)
    if picks is not None:
        for name, rounded in (('samp1', True), ('samp2', True), ('samp3', True), ('power', False), ('time', False)):
            grid = getattr(picks, name)
            if grid is not None:
                grid = onto_new(grid[:, moving])
                setattr(picks, name, np.round(grid) if rounded else grid)

    self.tnum = self.data.shape[1]
    self.trace_num = np.arange(self.tnum).astype(int) + 1
    self.dist = new
    self.trace_int = np.hstack((np.array(np.nanmean(np.diff(new))), np.diff(new))) * 1000.
    _set_flag_pair(self.flags, 'interp', spacing)


def elev_correct(self, v_avg=1.69e8):
    """Shift every trace down by its surface elevation difference; mirrors _RadarDataProcessing.py:587-637."""
    if getattr(self, 'nmo_depth', None) is None:
        raise ValueError('Run nmo before elev_correct so that we have depth scale')
    top, bottom = np.max(self.elev), np.min(self.elev)
    below_top = top - self.elev                          # how far every trace's surface sits below the highest one
    dz = self.dt * (v_avg / 2.)
    extra_rows = int(np.floor(np.max(below_top) / dz))
    shift = (below_top / dz).astype(int)

    x, suffix, host_dtype = _stage_float(self.data)
    kernel_suffix, out_dtype = _out_kind(self, suffix, host_dtype)
    nrow = int(x.shape[0])
    if np.any(shift > extra_rows) or np.any(shift < 0):
        raise ValueError('could not broadcast input array: a trace would be shifted outside the padded radargram')
    _finish(self, shift_traces_device(x, kernel_suffix, out_dtype, -shift, nrow + extra_rows), host_dtype)

    if _picks(self) is not None:
        self.picks.crop(-shift - 1)
    self.elevation = np.hstack((np.arange(top, bottom, -dz), bottom - self.nmo_depth))
    self.flags.elev = 1


def clean_GPS(self):
    """Fill GPS gaps by linear interpolation / extrapolation over the trace number; mirrors
    _RadarDataProcessing.py:640-654 (O(tnum) host vectors only - the radargram is untouched)."""
    from scipy.interpolate import interp1d
    for name in ('x_coord', 'y_coord', 'decday', 'lat', 'long', 'elev'):
        vec = getattr(self, name, None)
        if vec is None:
            continue
        ok = np.isfinite(vec)
        setattr(self, name, interp1d(self.trace_num[ok], vec[ok], fill_value="extrapolate",
                                     assume_sorted=True)(self.trace_num))
