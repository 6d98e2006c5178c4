"""StoDeep / ImpDAR ``.mat`` files for ``impdar_b200.RadarData``: the on-disk format either side
of the hot path.  ``load_mat`` mirrors ``RadarData.__init__(fn_mat)`` and ``check_attrs``
(RadarData/__init__.py:207-321), ``save`` mirrors ``RadarData.save`` (RadarData/_RadarDataSaving.py:32-78) including
its dtype-preservation rules; ``RadarFlags.to_matlab / from_matlab`` follow RadarFlags.py:63-104.

The container itself is scipy.io's MATLAB v5 reader / writer, exactly what the reference uses.  What changes for the
device path: the radargram of a loaded file lands in page-locked memory (when a CUDA device is present), so the upload
that follows is one DMA at PCIe speed instead of a staged copy.  Picks are interpretation state outside the hot path:
a ``picks`` struct found in a file is kept raw as ``dat.picks_struct`` and ``save`` writes it back verbatim as long as
the radargram still has the axes it was loaded with (load -> filter / migrate -> save keeps the picks); once a step has
changed the trace or sample axis the raw struct is stale, and ``save`` warns (RuntimeWarning) instead of silently
dropping or corrupting it.  An object with a real ``picks.to_struct()`` - ImpDAR's own RadarData - is written like the
reference does.
"""
import numpy as np

from .processing import ImpdarError
from .radardata import RadarData, RadarFlags

#: names a radargram may be stored under, in priority order (RadarData/__init__.py:23)
STODEEP_ATTRS = ['data', 'migdata', 'interp_data', 'nmo_data', 'filtdata', 'hfilt_data']
#: attributes every file must carry / may carry (RadarData/__init__.py:38-61)
ATTRS_GUARANTEED = ['chan', 'data', 'decday', 'dt', 'pressure', 'snum', 'tnum', 'trace_int', 'trace_num',
                    'travel_time', 'trig', 'trig_level']
ATTRS_OPTIONAL = ['nmo_depth', 'lat', 'long', 'elev', 'dist', 'x_coord', 'y_coord', 'fn', 't_srs']
_PER_TRACE = ['lat', 'long', 'pressure', 'trig', 'elev', 'dist', 'x_coord', 'y_coord', 'decday']


def _unbox(value):
    """MATLAB stores everything 2-D: (1, 1) -> scalar, a row or column -> vector, anything else as is."""
    if value.shape == (1, 1):
        return value[0][0]
    if value.shape[0] == 1 or (len(value.shape) > 1 and value.shape[1] == 1):
        return value.flatten()
    return value


def _pick_radargram(mat):
    """The first of STODEEP_ATTRS present becomes ``data``; lower-priority arrays keep their own names."""
    found = {}
    for name in STODEEP_ATTRS:
        if name in mat:
            arr = mat[name]
            if len(arr.dtype) > 0:
                print('Warning: Multiple arrays stored in {:s}, taking the first.'.format(name))
                arr = arr[0][0][0]
            found[name] = arr
    for rank, name in enumerate(STODEEP_ATTRS):
        if name in found:
            if rank > 0:
                print('First priority data {:s} not in structure, using {:s}'.format(STODEEP_ATTRS[0], name))
                print('(caused a rename of {:s}'.format(name))
                found['data'] = found.pop(name)
            return found
    raise KeyError('Data do not appear to be in StoDeep format')


def _pin(array):
    """A copy of `array` in page-locked memory when a CUDA device is there (float / integer dtypes torch can carry)."""
    try:
        import torch
        if not torch.cuda.is_available():
            return array
        src = torch.from_numpy(np.ascontiguousarray(array))
        host = torch.empty(src.shape, dtype=src.dtype, pin_memory=True)
        host.copy_(src)
        return host.numpy()
    except (TypeError, RuntimeError):
        return array


def check_attrs(dat):
    """Mirror of RadarData.check_attrs (RadarData/__init__.py:267-321): required attributes present, shapes consistent,
    MATLAB's scalar zeros turned back into None, a scalar trigger broadcast to a vector."""
    for attr in ATTRS_GUARANTEED + ['fn']:
        if not hasattr(dat, attr):
            raise ImpdarError('{:s} is missing. It appears that this is an ill-defined RadarData object'.format(attr))
        if getattr(dat, attr) is None:
            raise ImpdarError('{:s} is None. It appears that this is an ill-defined RadarData object'.format(attr))
    for attr in ATTRS_OPTIONAL:
        if not hasattr(dat, attr):
            raise ImpdarError('{:s} is missing. It appears that this is an ill-defined RadarData object'.format(attr))
    if (dat.data.shape != (dat.snum, dat.tnum)) and (dat.elev is None):
        raise ImpdarError('The data shape does not match the snum and tnum values!!!')
    if getattr(dat, 'nmo_depth', None) is not None:
        if (dat.nmo_depth.shape[0] != dat.snum) and (dat.elev is None):
            raise ImpdarError('The nmo_depth shape does not match the tnum value!!!')
    for attr in _PER_TRACE:
        val = getattr(dat, attr, None)
        if val is None:
            continue
        if (not hasattr(val, 'shape')) or len(val.shape) < 1:
            if val == 0:
                setattr(dat, attr, None)            # None written through MATLAB comes back as a scalar zero
            elif attr == 'trig':
                dat.trig = np.ones((dat.tnum,), dtype=int) * int(dat.trig)
            else:
                raise ImpdarError('{:s} needs to be a vector'.format(attr))
        elif val.shape[0] != dat.tnum:
            raise ImpdarError('{:s} needs length tnum {:d}'.format(attr, dat.tnum))
    if getattr(dat, 'data_dtype', None) is None:
        dat.data_dtype = dat.data.dtype


def _struct_to_plain(v):
    """A MATLAB struct as scipy.io.loadmat returns it ((1, 1) structured array) -> nested dicts of its fields, which
    scipy.io.savemat writes back as the same struct (writing the structured array itself would box every field once
    more).  load -> save -> load is then the identity on the struct, however often it is repeated."""
    if isinstance(v, np.ndarray) and v.dtype.names:
        return {n: _struct_to_plain(v[n][0, 0]) for n in v.dtype.names}
    return v


def load_mat(fn_mat, pinned=True):
    """Read a StoDeep / ImpDAR .mat file into an impdar_b200.RadarData; mirrors RadarData/__init__.py:207-244
    (KeyError for files that are not in the format, ImpdarError for inconsistent ones)."""
    from scipy.io import loadmat
    mat = loadmat(fn_mat)
    dat = RadarData(None)
    for attr in ATTRS_GUARANTEED:
        if attr == 'data':
            for name, arr in _pick_radargram(mat).items():
                setattr(dat, name, arr)
        elif attr not in mat:
            raise KeyError('.mat file does not appear to be in the StoDeep/ImpDAR format')
        else:
            setattr(dat, attr, _unbox(mat[attr]))
    for attr in ATTRS_OPTIONAL:
        setattr(dat, attr, _unbox(mat[attr]) if attr in mat else None)
    dat.data_dtype = dat.data.dtype
    if pinned:
        dat.data = _pin(dat.data)
    dat.fn = fn_mat
    dat.flags = RadarFlags()
    dat.flags.from_matlab(mat['flags'])
    dat.picks = None
    dat.picks_struct = mat['picks'] if 'picks' in mat else None      # raw MATLAB struct; save() writes it back verbatim
    dat._picks_axes = (int(dat.data.shape[0]), int(dat.data.shape[1]))  # ... while the radargram keeps these axes
    check_attrs(dat)
    return dat


def save(self, fn):
    """Write the radargram and its metadata as a .mat file; mirrors RadarData/_RadarDataSaving.py:32-78: None in a
    guaranteed attribute is written as 0, the radargram goes back to the dtype it was loaded with (data_dtype) unless
    NaNs have appeared in integer data (then the smallest float that keeps them)."""
    from scipy.io import savemat
    from . import device
    mat = {}
    for attr in ATTRS_GUARANTEED:
        val = getattr(self, attr, None)
        mat[attr] = val if val is not None else 0
    for attr in ATTRS_OPTIONAL + STODEEP_ATTRS:
        val = getattr(self, attr, None)
        if val is not None:
            mat[attr] = val
    if device.is_device_array(mat['data']):                           # device-resident lane: download for the file
        mat['data'] = mat['data'].cpu().numpy()
    picks = getattr(self, 'picks', None)
    picks_struct = getattr(self, 'picks_struct', None)
    if picks is not None:
        mat['picks'] = picks.to_struct()
    elif picks_struct is not None:
        # load_mat keeps the file's picks as the raw MATLAB struct (no Picks object on this side of the seam): it goes
        # back into the file verbatim, so load_mat -> process -> save does not lose picks, as long as the per-trace
        # arrays still match the radargram; after a step that changed the trace or sample axis they are stale, and
        # dropping them silently would lose data without a trace - warn loudly instead.
        ok = True
        has_picks = True
        try:
            samp = picks_struct['samp1'][0, 0]
            # an unpicked profile carries the placeholder the reference writes for None (a scalar 0): nothing to keep
            has_picks = samp.size > 1 or (samp.size == 1 and samp.ravel()[0] != 0)
            tnum = int(np.shape(mat['data'])[1])
            ok = samp.ndim < 2 or samp.shape[1] == tnum
            loaded = getattr(self, '_picks_axes', None)
            if loaded is not None and loaded != (int(np.shape(mat['data'])[0]), tnum):
                ok = False
        except Exception:
            ok = False
        if not has_picks:
            pass
        elif ok:
            mat['picks'] = _struct_to_plain(picks_struct)
        else:
            import warnings
            warnings.warn('%s: the picks loaded with this file no longer match the radargram (the trace or sample axis '
                          'changed) and impdar_b200 carries no Picks object to update them; they are NOT written' % fn,
                          RuntimeWarning, stacklevel=2)
    flags = self.flags if self.flags is not None else RadarFlags()
    mat['flags'] = flags.to_matlab()

    want = getattr(self, 'data_dtype', None)
    if want is not None and want != mat['data'].dtype:
        has_nan = np.issubdtype(mat['data'].dtype, np.floating) and bool(np.any(np.isnan(mat['data'])))
        if want in [int, np.int8, np.int16] and has_nan:
            print('Warning: new file is float16 rather than ', want, ' since we now have NaNs')
            mat['data'] = mat['data'].astype(np.float16)
        elif want in [np.int32] and has_nan:
            print('Warning: new file is float32 rather than ', want, ' since we now have NaNs')
            mat['data'] = mat['data'].astype(np.float32)
        elif want in [np.int64] and has_nan:
            print('Warning: new file is float64 rather than ', want, ' since we now have NaNs')
            mat['data'] = mat['data'].astype(np.float64)
        else:
            mat['data'] = mat['data'].astype(want)
    savemat(fn, mat)
