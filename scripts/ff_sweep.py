"""Scratch: filtfilt timing sweep over CTA sizes (IMPDAR_FF_BLOCK) for the loaded library build."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scipy.signal import butter
from impdar_b200 import filtering as fl

def ev(fn, n=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(n):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts)

S, T, B = 2048, 8192, 8
x = torch.randn(B, S, T, device='cuda')
b, a = butter(5, [2e6 / 50e6, 10e6 / 50e6], 'bandpass')
b2 = b.copy(); b2[1] = 1e-30    # defeats the zero-b specialisation: same arithmetic cost as a general numerator
for blk in (32, 64, 128, 256):
    os.environ['IMPDAR_FF_BLOCK'] = str(blk)
    ms = ev(lambda: fl.filtfilt_device(x, 'f32', b, a))
    ms2 = ev(lambda: fl.filtfilt_device(x, 'f32', b2, a))
    print(f'{os.environ.get("IMPDAR_B200_LIB", "default")[-14:]} block {blk}: bandpass(BZ) {ms:.3f} ms  general {ms2:.3f} ms  ({B*S*T*8/ms*1e3/1e9:.0f} GB/s on 8 B/sample)', flush=True)
