"""Development check of the tensor-core constant-velocity phase shift (impdar_phsh_set_legacy(2)) against the oracle and
against the SIMT pair kernel, then its timing at BASELINE config 3."""
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
shapes = [(256, 64), (300, 50), (1024, 256), (4096, 256), (2048, 130)]
if len(sys.argv) > 1 and sys.argv[1] == "big":
    shapes = []
for (S, T) in shapes:
    d = synthetic_dat(S, T, seed=31)
    x64 = d.data.astype(np.float64)
    _, want = om.phase_shift(x64, d.dt, d.travel_time, d.trace_int, d.dist, VEL, 10, 10)
    xd = torch.from_numpy(d.data).cuda()
    res = {}
    for mode in (0, 2):
        lib.impdar_phsh_set_legacy(mode)
        got = ml.phase_shift_device(xd, d.dt, 5.0, d.travel_time, VEL, 10, 10).double().cpu().numpy()
        lib.impdar_phsh_set_legacy(0)
        res[mode] = np.linalg.norm(got - want) / np.linalg.norm(want)
    print("S=%d T=%d rel-L2 vs oracle: pair kernel %.3e  tensor-core %.3e" % (S, T, res[0], res[2]), flush=True)

def ev(fn, n=3):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts)

for (S, T) in [(4096, 16384)]:
    x = torch.randn(S, T, device='cuda')
    tt = np.arange(S) * 0.01
    for mode in (0, 2):
        lib.impdar_phsh_set_legacy(mode)
        ms = ev(lambda: ml.phase_shift_device(x, 1e-8, 5.0, tt, VEL, 10, 10))
        lib.impdar_phsh_set_legacy(0)
        print("phsh const %dx%d mode %d: %.2f ms" % (S, T, mode, ms), flush=True)
