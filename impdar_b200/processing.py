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
def reverse(self):
    """Flip the profile left-right; mirrors _RadarDataProcessing.py:20-47."""
    S, T = int(self.data.shape[0]), int(self.data.shape[1])
    _crop_any(self, 0, S, 0, T, flip_lr=True)
    for name in ('x_coord', 'y_coord', 'decday', 'lat', 'long', 'elev'):
        if getattr(self, name, None) is not None:       # impdar_b200.RadarData leaves absent GPS vectors as None
            setattr(self, name, np.flip(getattr(self, name), 0))
    if getattr(self, 'picks', None) is not None:
        self.picks.reverse()
    if self.flags.reverse:
        print('Back to original direction')
        self.flags.reverse = False
    else:
        print('Profile direction reversed')
        self.flags.reverse = True


def constant_sample_depth_spacing(self):
    """Resample rows to constant depth spacing; mirrors _RadarDataProcessing.py:50-63."""
    if self.nmo_depth is None:
        raise AttributeError('Call nmo first...')
    if np.allclose(np.diff(self.nmo_depth), np.ones((self.snum - 1,)) * (self.nmo_depth[1] - self.nmo_depth[0])):
        print('No constant sampling when you already have constant sampling...')
        return 1
    depths = np.linspace(np.min(self.nmo_depth[0], 0), self.nmo_depth[-1], len(self.nmo_depth))
    x, suffix, host_dtype = _stage_float(self.data)
    nodes = linear_nodes_scipy(self.nmo_depth, depths)
    so, od = _out_kind(self, suffix, host_dtype)
    _finish(self, interp_rows_device(x, so, od, nodes, 0), host_dtype)
    self.travel_time = _interp1d_vector(self.nmo_depth, self.travel_time, depths)
    self.nmo_depth = depths


def traveltime_to_depth(self, profile_depth, profile_rho, c=3.0e8, permittivity_model=firn_permittivity):
    """Depth of every sample for a density profile; mirrors _RadarDataProcessing.py:191-235 (O(snum) host)."""
    profile_u = c / np.sqrt(np.real(permittivity_model(profile_rho)))
    depth = self.travel_time / 2. * c / np.sqrt(np.real(permittivity_model(917.))) * 1.0e-6
    z = 0.
    first = self.dt * 1.0e6
    for i, t in enumerate(self.travel_time):
        if t < 0.:
            continue
        if t < first:
            z += t / 2. * profile_u[0] * 1.0e-6
        else:
            z += self.dt / 2. * profile_u[np.nanargmin(abs(profile_depth - z))]
        depth[i] = z
    return depth


def _nmo_times(self, ant_sep, uice, u_interp=None, d_interp=None):
    """Vertical two-way time of every sample (_RadarDataProcessing.py:133-160), float64 host."""
    tt = self.travel_time
    nmotime = np.zeros((len(tt)))
    for i, t in enumerate(tt):
        if u_interp is None:
            u_rms = uice
        else:
            d = t / 2. * uice * 1.0e-6
            d_last = d.copy()
            j, tol = 0, 0.1 * self.dt / 2. * uice
            while abs(d - d_last) > tol or j < 5:
                d_last = d.copy()
                u_rms = np.sqrt(np.mean(u_interp[d_interp <= d] ** 2.))
                d = t / 2. * u_rms * 1.0e-6
                j += 1
        tsep_ice = 1e6 * (ant_sep / u_rms)
        nmotime[i] = np.sqrt((t + tsep_ice) ** 2. - tsep_ice ** 2.)
    return nmotime


def nmo(self, ant_sep, uice=1.69e8, uair=3.0e8, const_firn_offset=None, rho_profile=None,
        permittivity_model=firn_permittivity, const_sample=False):
    """Normal move-out correction; mirrors _RadarDataProcessing.py:66-188.  The per-trace
    ``interp1d(nmotime, trace)(travel_time)`` loop is one row-interpolation pass on the device."""
    if np.any(self.trig > 0):
        raise ImpdarError('Crop out the pretrigger before doing the nmo correction.')

    u_interp = d_interp = None
    if rho_profile is not None:
        try:
            rho_profile_data = np.genfromtxt(rho_profile, delimiter=',')
            profile_depth = rho_profile_data[:, 0]
            profile_rho = rho_profile_data[:, 1]
        except IndexError:
            raise IndexError('Cannot load the depth-density profile')
        profile_u = uair / np.sqrt(np.real(permittivity_model(profile_rho)))
        d_interp = np.linspace(np.min(profile_depth, 0), max(profile_depth), 10 * self.snum)
        u_interp = _interp1d_vector(profile_depth, profile_u, d_interp)
        print('Iterating velocity profile in firn...')

    nmotime = _nmo_times(self, ant_sep, uice, u_interp, d_interp)
    new_tt = np.arange(min(self.travel_time), max(nmotime), self.dt * 1e6)

    x, suffix, host_dtype = _stage_float(self.data)
    # scipy hands a 1-D float64 trace to numpy.interp and everything else to its own two-weight formula
    if suffix == 'f64':
        nodes, mode = linear_nodes_numpy(nmotime, new_tt), 1
    else:
        nodes, mode = linear_nodes_scipy(nmotime, new_tt), 0
    so, od = _out_kind(self, suffix, host_dtype)
    out = interp_rows_device(x, so, od, nodes, mode)
    self.travel_time = new_tt
    self.snum = len(new_tt)
    _finish(self, out, host_dtype)

    if rho_profile is None:
        self.nmo_depth = self.travel_time / 2. * uice * 1.0e-6
    else:
        self.nmo_depth = traveltime_to_depth(self, profile_depth, profile_rho, c=uair,
                                             permittivity_model=permittivity_model)
    if const_sample:
        constant_sample_depth_spacing(self)
    if const_firn_offset is not None:
        self.nmo_depth = self.nmo_depth + const_firn_offset
    print('Normal Moveout filter complete.')
    try:
        self.flags.nmo[0] = 1
        self.flags.nmo[1] = ant_sep
    except (IndexError, TypeError):
        self.flags.nmo = np.ones((2, ))
        self.flags.nmo[1] = ant_sep


def crop(self, lim, top_or_bottom='top', dimension='snum', uice=1.69e8, rezero=True, zero_trig=True):
    """Crop in the vertical; mirrors _RadarDataProcessing.py:238-349."""
    if top_or_bottom not in ['top', 'bottom']:
        raise ValueError('top_or_bottom must be "top" or "bottom" not {:s}'.format(top_or_bottom))
    if dimension not in ['snum', 'twtt', 'depth', 'pretrig']:
        raise ValueError('Dimension must be in [\'snum\', \'twtt\', \'depth\']')
    if top_or_bottom == 'bottom' and dimension == 'pretrig':
        raise ValueError('Only use pretrig to crop from the top')

    if dimension == 'twtt':
        ind = np.min(np.argwhere(self.travel_time >= lim))
    elif dimension == 'depth':
        nmo_depth = getattr(self, 'nmo_depth', None)
        depth = nmo_depth if nmo_depth is not None else self.travel_time / 2. * uice * 1.0e-6
        ind = np.min(np.argwhere(depth >= lim))
    elif dimension == 'pretrig':
        ind = self.trig.astype(int) if isinstance(self.trig, np.ndarray) else int(self.trig)
    else:
        ind = int(lim)

    S, T = int(self.data.shape[0]), int(self.data.shape[1])
    if not isinstance(ind, np.ndarray) or (dimension != 'pretrig'):
        if top_or_bottom == 'top':
            lims = [ind, S]
            self.trig = self.trig - ind
            if zero_trig:
                self.trig = np.zeros_like(self.trig)
        else:
            lims = [0, ind]
        r0, r1, _ = slice(lims[0], lims[1]).indices(S)       # Python slice semantics (negative / oversize limits)
        _crop_any(self, r0, max(r0, r1), 0, T)
        self.travel_time = self.travel_time[lims[0]:lims[1]]
        if rezero:
            self.travel_time = self.travel_time - self.travel_time[0]
        if getattr(self, 'nmo_depth', None) is not None:
            self.nmo_depth = self.nmo_depth[lims[0]:lims[1]]
        self.snum = self.data.shape[0]
    else:
        # pretrigger given per trace: shift every trace up by its own trigger sample, NaN below
        ind = np.asarray(ind)
        if ind.shape != (T,):
            raise ValueError('trig must have one entry per trace')
        mintrig = np.nanmin(ind)
        if mintrig < 0:
            raise ValueError('could not broadcast input array: negative trigger samples cannot be cropped')
        lims = [mintrig, S]
        self.trig = self.trig - ind
        x, suffix, host_dtype = _stage_float(self.data)
        so, od = _out_kind(self, suffix, host_dtype)
        _finish(self, shift_traces_device(x, so, od, ind, S - int(mintrig)), host_dtype)
        self.travel_time = self.travel_time[lims[0]:lims[1]]
        if rezero:
            self.travel_time = self.travel_time - self.travel_time[0]
        self.snum = self.data.shape[0]

    if top_or_bottom == 'top':
        if getattr(self, 'picks', None) is not None:
            self.picks.crop(ind)

    try:
        self.flags.crop[0] = 1
        self.flags.crop[2] = self.flags.crop[1] + lims[1]
    except (IndexError, TypeError):
        self.flags.crop = np.zeros((3,))
        self.flags.crop[0] = 1
        self.flags.crop[2] = self.flags.crop[1] + lims[1]
    self.flags.crop[1] = self.flags.crop[1] + lims[0]
    print('Vertical samples reduced to subset [{:d}:{:d}] of original'.format(
        int(self.flags.crop[1]), int(self.flags.crop[2])))


def hcrop(self, lim, left_or_right='left', dimension='tnum'):
    """Crop in the horizontal; mirrors _RadarDataProcessing.py:352-421."""
    if left_or_right not in ['left', 'right']:
        raise ValueError('left_or_right must be left or right, not {:s}'.format(left_or_right))
    if dimension not in ['tnum', 'dist']:
        raise ValueError('Dimension must be in ["tnum", "dist"]')

    if dimension == 'dist':
        if lim > np.max(self.dist):
            raise ValueError('lim is larger than largest distance')
        if lim <= 0:
            raise ValueError('Distance should be strictly positive')
        ind = np.min(np.argwhere(self.dist >= lim))
    else:
        if int(lim) in (0, 1):
            raise ValueError('lim should be at least two to preserve some data')
        if lim > self.tnum:
            raise ValueError('lim should be less than tnum+1 {:d} in order to do anything'.format(self.tnum + 1))
        if lim == -1 or lim < -int(self.tnum):
            raise ValueError('If negative, lim should be in [-self.tnum; -1)')
        ind = int(lim) - 1

    S, T = int(self.data.shape[0]), int(self.data.shape[1])
    lims = [ind, T] if left_or_right == 'left' else [0, ind]
    c0, c1, _ = slice(lims[0], lims[1]).indices(T)
    _crop_any(self, 0, S, c0, max(c0, c1))
    for var in ['lat', 'long', 'pressure', 'trace_int', 'trig', 'elev', 'x_coord', 'y_coord', 'decday']:
        val = getattr(self, var, None)
        if val is not None and isinstance(val, np.ndarray):
            setattr(self, var, val[lims[0]:lims[1]])
    if getattr(self, 'picks', None) is not None:
        self.picks.hcrop(lims)
    if self.dist is not None:
        self.dist = self.dist[lims[0]:lims[1]] - self.dist[lims[0]]
    if getattr(self, 'trace_num', None) is not None:
        self.trace_num = self.trace_num[lims[0]:lims[1]] - lims[0] + 1
    self.tnum = self.data.shape[1]


_RESTACK_VARS = ('dist', 'pressure', 'lat', 'long', 'x_coord', 'y_coord', 'elev', 'decday', 'trig')


def restack(self, traces):
    """Average groups of `traces` adjacent traces; mirrors _RadarDataProcessing.py:424-477."""
    traces = int(traces)
    if traces % 2 == 0:
        print('Only will stack odd numbers of traces. Using {:d}'.format(int(traces + 1)))
        traces = traces + 1
    tnum = int(np.floor(self.tnum / traces))
    x, suffix, host_dtype = _stage_float(self.data)
    so, od = _out_kind(self, suffix, host_dtype)
    out = restack_device(x, so, od, traces)

    new_vectors = {}
    for key in _RESTACK_VARS:
        val = getattr(self, key, None)
        if val is None:
            new_vectors[key] = None
            continue
        grouped = np.zeros((tnum, ))
        for j in range(tnum):
            grouped[j] = np.mean(val[j * traces:min((j + 1) * traces, int(x.shape[1]))])
        new_vectors[key] = grouped
    self.tnum = tnum
    _finish(self, out, host_dtype)
    self.trace_num = np.arange(self.tnum).astype(int) + 1
    self.trace_int = np.zeros((tnum, ))
    if getattr(self, 'picks', None) is not None:
        self.picks.restack(traces)
    for key, val in new_vectors.items():
        setattr(self, key, val)
    self.flags.restack = True


def constant_space(self, spacing, min_movement=1.0e-2, show_nomove=False):
    """Resample to constant trace spacing; mirrors _RadarDataProcessing.py:499-584: stationary traces are dropped
    (column compaction) and the rest interpolated linearly in distance - one column-gather pass on the device."""
    good_vals = np.hstack((np.array([True]), np.diff(self.dist * 1000.) >= min_movement))
    for i in range(len(self.dist)):
        if not good_vals[i]:
            self.dist[i:] = self.dist[i:] - (self.dist[i] - self.dist[i - 1])
    temp_dist = self.dist[good_vals]
    new_dists = np.arange(np.min(temp_dist), np.max(temp_dist), step=spacing / 1000.0)

    nodes = linear_nodes_scipy(temp_dist, new_dists)
    kept = np.flatnonzero(good_vals)
    nodes['lo'], nodes['hi'] = kept[nodes['lo']], kept[nodes['hi']]      # compaction folded into the gather

    is_complex = (not device.is_device_array(self.data)) and np.iscomplexobj(np.asarray(self.data))
    if is_complex:
        # real weights times complex samples act on the two components separately: run the float64 kernel on the
        # interleaved (snum, 2 tnum) view with every node duplicated for the real and the imaginary column
        z = np.ascontiguousarray(np.asarray(self.data), dtype=np.complex128)
        both = np.repeat(nodes, 2)
        both['lo'] = 2 * both['lo'] + np.tile([0, 1], len(nodes))
        both['hi'] = 2 * both['hi'] + np.tile([0, 1], len(nodes))
        import torch
        x = device.to_device(z.view(np.float64), torch.float64)
        out = interp_cols_device(x, 'f64', torch.float64, both, 0)
        self.data = device.to_host(out, np.float64).view(np.complex128)
    else:
        x, suffix, host_dtype = _stage_float(self.data)
        so, od = _out_kind(self, suffix, host_dtype)
        _finish(self, interp_cols_device(x, so, od, nodes, 0), host_dtype)

    for attr in ['lat', 'long', 'x_coord', 'y_coord', 'decday', 'pressure', 'trig']:
        setattr(self, attr, _interp1d_vector(temp_dist, getattr(self, attr)[good_vals], new_dists))
    for attr in ['elev']:
        if getattr(self, attr) is not None:
            setattr(self, attr, _interp1d_vector(temp_dist, getattr(self, attr)[good_vals], new_dists))

    picks = getattr(self, 'picks', None)
    if picks is not None:
        for attr in ['samp1', 'samp2', 'samp3']:
            if getattr(picks, attr) is not None:
                setattr(picks, attr, np.round(_interp1d_vector(temp_dist, getattr(picks, attr)[:, good_vals],
                                                               new_dists)))
        for attr in ['power', 'time']:
            if getattr(picks, attr) is not None:
                setattr(picks, attr, _interp1d_vector(temp_dist, getattr(picks, attr)[:, good_vals], new_dists))

    self.tnum = self.data.shape[1]
    self.trace_num = np.arange(self.tnum).astype(int) + 1
    self.dist = new_dists
    self.trace_int = np.hstack((np.array(np.nanmean(np.diff(self.dist))), np.diff(self.dist))) * 1000.
    try:
        self.flags.interp[0] = 1
        self.flags.interp[1] = spacing
    except (IndexError, TypeError):
        self.flags.interp = np.ones((2,))
        self.flags.interp[1] = spacing


def elev_correct(self, v_avg=1.69e8):
    """Shift every trace down by its surface elevation difference; mirrors _RadarDataProcessing.py:587-637."""
    if getattr(self, 'nmo_depth', None) is None:
        raise ValueError('Run nmo before elev_correct so that we have depth scale')
    elev_diffs = np.max(self.elev) - self.elev
    max_diff = np.max(elev_diffs)
    dz_avg = self.dt * (v_avg / 2.)
    max_samp = int(np.floor(max_diff / dz_avg))
    top_inds = (elev_diffs / dz_avg).astype(int)

    x, suffix, host_dtype = _stage_float(self.data)
    so, od = _out_kind(self, suffix, host_dtype)
    S = int(x.shape[0])
    if np.any(top_inds + S > S + max_samp) or np.any(top_inds < 0):
        raise ValueError('could not broadcast input array: a trace would be shifted outside the padded radargram')
    _finish(self, shift_traces_device(x, so, od, -top_inds, S + max_samp), host_dtype)

    if getattr(self, 'picks', None) is not None:
        self.picks.crop(-top_inds - 1)
    self.elevation = np.hstack((np.arange(np.max(self.elev), np.min(self.elev), -dz_avg),
                                np.min(self.elev) - self.nmo_depth))
    self.flags.elev = 1
