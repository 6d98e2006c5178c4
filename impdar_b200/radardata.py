"""A minimal host-side radargram object with the attributes the hot path touches
(RadarData/__init__.py:132-204: data, snum, tnum, dt, travel_time [us], dist [km], trace_int [m], flags)
and the hot-path methods bound to the B200 backend.  It exists so the backend can be used and tested
where ImpDAR itself is not installed (the GPU box); with ImpDAR present, ``impdar_b200.install()`` binds
the same functions onto ImpDAR's own RadarData instead.
"""
import numpy as np

from . import filtering, processing


class RadarFlags(object):
    """The processing-history flags the hot path sets (RadarFlags.py:44-61)."""

    def __init__(self):
        self.batch = False
        self.bpass = np.zeros((3,))
        self.hfilt = np.zeros((2,))
        self.rgain = False
        self.agc = False
        self.restack = False
        self.reverse = False
        self.crop = np.zeros((3,))
        self.nmo = np.zeros((2,))
        self.interp = np.zeros((2,))
        self.mig = 'none'
        self.elev = 0
        self.elevation = 0
        self.attrs = ['batch', 'bpass', 'hfilt', 'rgain', 'agc', 'restack', 'reverse', 'crop', 'nmo', 'interp', 'mig',
                      'elev']
        self.attr_dims = [None, 3, 2, None, None, None, None, 3, 2, 2, None, None]
        self.bool_attrs = ['agc', 'batch', 'restack', 'reverse', 'rgain']

    def to_matlab(self):
        """dict for scipy.io.savemat; booleans as 0 / 1 (RadarFlags.py:63-75)."""
        out = {name: getattr(self, name) for name in self.attrs}
        for name in self.bool_attrs:
            out[name] = 1 if out[name] else 0
        return out

    def from_matlab(self, matlab_struct):
        """Fill from the `flags` struct of scipy.io.loadmat (RadarFlags.py:77-104): vector flags that MATLAB stored as a
        lazily appended scalar are re-allocated at their proper length."""
        for name, dim in zip(self.attrs, self.attr_dims):
            val = matlab_struct[name][0][0][0]
            if dim is not None and val.shape[0] == 1:
                val = np.zeros((dim, ))
            setattr(self, name, val)
        for name in self.bool_attrs:
            setattr(self, name, True if matlab_struct[name][0][0][0] == 1 else 0)


class RadarData(object):
    """Radargram + geometry.  ``data`` is (snum, tnum), C-order, traces contiguous."""

    def __init__(self, data=None, dt=None, travel_time=None, dist=None, trace_int=None):
        self.data = data
        self.snum = None if data is None else int(data.shape[0])
        self.tnum = None if data is None else int(data.shape[1])
        self.dt = dt
        self.travel_time = travel_time
        self.dist = dist
        self.trace_int = trace_int
        self.flags = RadarFlags()
        self.fn = None
        self.trig = 0
        # trace-wise vectors and depth scale the index / resampling methods keep in step with the radargram
        # (RadarData/__init__.py:132-204); None = absent
        self.trace_num = None if data is None else np.arange(self.tnum) + 1
        self.lat = self.long = self.x_coord = self.y_coord = self.elev = self.decday = self.pressure = None
        self.nmo_depth = None
        self.elevation = None
        self.picks = None
        # file-format attributes (RadarData/__init__.py:38-61); filled by impdar_b200.load_mat
        self.chan = None
        self.trig_level = None
        self.t_srs = None
        self.data_dtype = None if data is None else np.asarray(data).dtype if not hasattr(data, 'is_cuda') else None

    def save(self, fn):
        """Write a StoDeep / ImpDAR .mat file (RadarData/_RadarDataSaving.py:32-78)."""
        from . import matio
        return matio.save(self, fn)

    adaptivehfilt = filtering.adaptivehfilt
    horizontalfilt = filtering.horizontalfilt
    hfilt = filtering.hfilt
    vertical_band_pass = filtering.vertical_band_pass
    migrate = filtering.migrate
    highpass = filtering.highpass
    lowpass = filtering.lowpass
    horizontal_band_pass = filtering.horizontal_band_pass
    winavg_hfilt = filtering.winavg_hfilt
    rangegain = filtering.rangegain
    agc = filtering.agc
    denoise = filtering.denoise
    reverse = processing.reverse
    crop = processing.crop
    hcrop = processing.hcrop
    restack = processing.restack
    nmo = processing.nmo
    constant_sample_depth_spacing = processing.constant_sample_depth_spacing
    traveltime_to_depth = processing.traveltime_to_depth
    constant_space = processing.constant_space
    elev_correct = processing.elev_correct
    clean_GPS = processing.clean_GPS
