"""One small call of every kernel family of the hot path, for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool racecheck python scripts/sanitize_r02.py
Shapes are small (the tools slow kernels down 10-100x) but chosen so that every code path of the large runs is taken:
several row blocks / stages / producer rounds of the Kirchhoff tile kernel, the five-pass Stolt transform, the (+w, -w)
pair kernels and the tensor-core kernel of the phase shift, both adaptive-filter strip kernels, filtfilt, hfilt."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from impdar_b200 import migrationlib as ml, filtering as fl, _lib

lib = _lib.load()
torch.manual_seed(0)
VEL = 1.69e8


def geom(S, T):
    return np.arange(S) * 0.01, np.arange(T) * 0.005


S, T = 160, 700
x = torch.randn(S, T, device="cuda")
tt, dk = geom(S, T)
out = ml.kirchhoff_device(x, tt, dk, VEL, False)
print("kirchhoff far:", ml.kirchhoff_last_kernel(), float(out.abs().sum()))
ml.kirchhoff_device(x, tt, dk, VEL, True)
print("kirchhoff near:", ml.kirchhoff_last_kernel())
ml.set_kirchhoff_mode(ml.KIRCHHOFF_GENERAL)
ml.kirchhoff_device(x, tt, dk, VEL, False, 100, 164)
ml.set_kirchhoff_mode(ml.KIRCHHOFF_AUTO)
c0, c1 = ml.kirchhoff_input_window(S, tt, dk, VEL, 300, 420)
ml.kirchhoff_window_device(x[:, c0:c1].contiguous(), c0, T, tt, dk, VEL, False, 300, 420)
h = ml.kirchhoff_host(x.cpu().numpy(), tt, dk, VEL, False, nchunks=3)
print("kirchhoff window / host pipeline ok", float(np.abs(h).sum()))

xs = torch.randn(256, 512, device="cuda")
ml.stolt_device(xs, 1e-8, 5.0, 1.68e8, 10, 10)
print("stolt:", ml.stolt_last_pipeline())
ml.stolt_device(torch.randn(96, 80, device="cuda"), 1e-8, 5.0, 1.68e8, 10, 10)

xp = torch.randn(256, 96, device="cuda")
tp_, _ = geom(256, 96)
for mode in (0, 2):
    lib.impdar_phsh_set_legacy(mode)
    ml.phase_shift_device(xp, 1e-8, 5.0, tp_, VEL, 10, 10)
lib.impdar_phsh_set_legacy(0)
ml.phase_shift_device(xp, 1e-8, 5.0, tp_, np.linspace(1.69e8, 2.2e8, 256), 10, 10)
print("phase shift ok")

xf = torch.randn(2, 200, 4096, device="cuda")
taper = np.exp(-np.arange(200) * 0.01 * 0.05)
fl.horizontalfilt_device(xf, 'f32', taper, 0, 4096)
for mode in (0, 2):
    lib.impdar_ahfilt_force_rowwise(mode)
    fl.adaptivehfilt_device(xf, 'f32', taper, 300)
lib.impdar_ahfilt_force_rowwise(0)
from scipy.signal import butter
b, a = butter(5, [0.04, 0.2], 'bandpass')
fl.filtfilt_device(xf, 'f32', b, a)
torch.cuda.synchronize()
print("filters ok")
