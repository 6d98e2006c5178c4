"""Diagnostic (development): where does the constant-velocity phase shift lose parity at snum = 4096?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch
from impdar_b200 import migrationlib as ml, _lib
from oracle import migration as om
from util import synthetic_dat

lib = _lib.load()
VEL = 1.69e8
for (S, T, legacy) in [(4096, 256, 0), (4096, 256, 1), (2048, 256, 0), (4096, 64, 0), (4096, 255, 0), (1024, 256, 0), (4096, 32, 0)]:
    d = synthetic_dat(S, T, seed=31)
    x64 = d.data.astype(np.float64)
    _, want = om.phase_shift(x64, d.dt, d.travel_time, d.trace_int, d.dist, VEL, 10, 10)
    lib.impdar_phsh_set_legacy(legacy)
    xd = torch.from_numpy(d.data).cuda()
    got = ml.phase_shift_device(xd, d.dt, 5.0, d.travel_time, VEL, 10, 10).double().cpu().numpy()
    lib.impdar_phsh_set_legacy(0)
    err = got - want
    rel = np.linalg.norm(err) / np.linalg.norm(want)
    ek = np.abs(np.fft.rfft(err, axis=1)).mean(axis=0)       # error spectrum over kx
    et = np.abs(np.fft.rfft(err, axis=0)).mean(axis=1)       # error spectrum over tau
    print("S=%d T=%d legacy=%d rel-L2 %.3e | top kx bins %s | top tau-freq bins %s | err row std %.2e, row-mean-err std %.2e"
          % (S, T, legacy, rel, np.argsort(ek)[-3:][::-1].tolist(), np.argsort(et)[-3:][::-1].tolist(),
             err.std(), err.mean(axis=0).std()), flush=True)
