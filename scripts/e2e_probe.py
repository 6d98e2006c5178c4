"""Scratch: where does host-to-host time go (PCIe bandwidth, pinned allocation, chain overlap)?"""
import os, sys, time, contextlib, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import impdar_b200
from impdar_b200 import synthetic, device

def wall(fn, n=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(n):
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    return min(ts) * 1e3

def quiet(fn):
    def g():
        with contextlib.redirect_stdout(io.StringIO()):
            return fn()
    return g

S, T = 2048, 4096
h32 = torch.randn(S, T).pin_memory(); d32 = torch.empty(S, T, device='cuda')
h64 = torch.empty(S, T, dtype=torch.float64).pin_memory(); d64 = torch.randn(S, T, device='cuda', dtype=torch.float64)
print('H2D 32 MiB pinned: %.3f ms' % wall(lambda: d32.copy_(h32, non_blocking=True)))
print('D2H 64 MiB pinned: %.3f ms' % wall(lambda: h64.copy_(d64, non_blocking=True)))
print('D2H 64 MiB pinned, 4 column blocks (2-D copies): %.3f ms' % wall(lambda: [h64[:, i * 1024:(i + 1) * 1024].copy_(d64[:, i * 1024:(i + 1) * 1024], non_blocking=True) for i in range(4)]))
print('pinned alloc 64 MiB (cached): %.3f ms' % wall(lambda: torch.empty(S, T, dtype=torch.float64, pin_memory=True)))
tt, dist, ti = synthetic.geometry(S, T)
def kir():
    d = impdar_b200.RadarData(h32.numpy(), dt=1e-8, travel_time=tt, dist=dist, trace_int=ti)
    d.migrate(mtype='kirch', vel=1.69e8)
    return d
print('kirchhoff e2e (RadarData.migrate on pinned host data): %.3f ms' % wall(quiet(kir)))
from impdar_b200 import migrationlib as ml
for nc in (1, 2, 4, 8, 16, 32):
    print('kirchhoff_host nchunks=%d: %.3f ms' % (nc, wall(lambda: ml.kirchhoff_host(h32.numpy(), tt, dist, 1.69e8, False, nchunks=nc))))
x = h32.cuda()
print('kirchhoff device only: %.3f ms' % wall(lambda: ml.kirchhoff_device(x, tt, dist, 1.69e8, False)))
print('to_device: %.3f ms' % wall(lambda: device.to_device(h32.numpy())))
o = ml.kirchhoff_device(x, tt, dist, 1.69e8, False)
print('to_host f64: %.3f ms' % wall(lambda: device.to_host(o, np.float64)))
# pipeline
S, T, P = 2048, 8192, 8
hp = torch.randn(P, S, T).pin_memory()
tt, dist, ti = synthetic.geometry(S, T)
for ns in (3,):
    def pipe():
        h = hp.numpy()
        dats = [impdar_b200.RadarData(h[p], dt=1e-8, travel_time=tt, dist=dist, trace_int=ti) for p in range(P)]
        impdar_b200.process.process(dats, vbp=(2, 10), hfilt=(0, T), migrate=True, n_streams=ns)
        return dats
    print('pipeline process() 8 profiles, n_streams=%d: %.3f ms' % (ns, wall(quiet(pipe), n=4)))
dp = torch.empty(P, S, T, device='cuda'); hq = torch.empty(P, S, T).pin_memory()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def both():
    with torch.cuda.stream(s1): dp.copy_(hp, non_blocking=True)
    with torch.cuda.stream(s2): hq.copy_(dp, non_blocking=True)
print('H2D 256 MiB || D2H 256 MiB concurrently: %.3f ms' % wall(both))
print('H2D 256 MiB alone: %.3f ms' % wall(lambda: dp.copy_(hp, non_blocking=True)))
