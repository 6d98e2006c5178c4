#!/usr/bin/env python
"""Generate tests/golden/mat_*.mat with the UNMODIFIED reference (build container only):
  mat_ref_saved.mat    small_data.mat loaded, cropped, nmo'ed and flagged by the reference, saved by RadarData.save
  mat_ref_resaved.mat  that file loaded and saved once more by the reference (what a load -> save round trip must give)
  mat_ref_f32.mat      a float32 radargram whose data_dtype is float32 (dtype preservation)
  mat_ref_picks.mat    small_data_picks.mat (a picked profile) loaded and saved by the reference: Picks.to_struct()
  mat_ref_picks_resaved.mat  that file through the reference's load -> save once more
      python tests/golden/make_golden_mat.py"""
import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle._refimport import import_reference  # noqa: E402

_, RadarData, _ = import_reference()
SRC = os.path.join(os.path.dirname(os.environ.get("IMPDAR_REFERENCE_SRC", "/root/reference/src")), "test", "input_data",
                   "small_data.mat")


def quiet(f, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return f(*a, **k)


d = RadarData(SRC)
d.trig = d.trig * 0.
quiet(d.crop, 3, 'top', 'snum')
quiet(d.nmo, 0.)
d.flags.bpass = np.array([1., 2., 10.])
d.flags.mig = 'stolt'
d.save(os.path.join(HERE, 'mat_ref_saved.mat'))
RadarData(os.path.join(HERE, 'mat_ref_saved.mat')).save(os.path.join(HERE, 'mat_ref_resaved.mat'))
d = RadarData(SRC)
d.data = d.data.astype(np.float32)
d.data_dtype = np.dtype(np.float32)
d.elev = None
d.save(os.path.join(HERE, 'mat_ref_f32.mat'))
SRC_P = os.path.join(os.path.dirname(SRC), "small_data_picks.mat")
RadarData(SRC_P).save(os.path.join(HERE, 'mat_ref_picks.mat'))
RadarData(os.path.join(HERE, 'mat_ref_picks.mat')).save(os.path.join(HERE, 'mat_ref_picks_resaved.mat'))
print('wrote', [f for f in os.listdir(HERE) if f.startswith('mat_')])
