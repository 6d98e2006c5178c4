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
