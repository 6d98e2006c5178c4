"""Scratch timing of every device entry point (not the bench; for development on gpurun)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import impdar_b200
from impdar_b200 import migrationlib as ml, filtering as fl

def ev(fn, n=3, warm=1):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts=[]
    for _ in range(n):
        a=torch.cuda.Event(enable_timing=True); b=torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts)

def geom(S,T,dt=1e-8,dx=5.0):
    return np.arange(S)*dt*1e6, np.arange(T)*dx/1e3

which = sys.argv[1:] or ['kirch','stolt','filters','phsh']
torch.manual_seed(0)
if 'kirch' in which:
    for (S,T) in [(2048,4096)]:
        x=torch.randn(S,T,device='cuda'); tt,dist=geom(S,T)
        ml.enable_kirchhoff_stats(True)
        ml.kirchhoff_device(x,tt,dist,1.69e8,False); torch.cuda.synchronize()
        pairs,exact=ml.kirchhoff_stats()
        ml.enable_kirchhoff_stats(False)
        ms=ev(lambda: ml.kirchhoff_device(x,tt,dist,1.69e8,False))
        print(f'kirch {S}x{T}: {ms:.2f} ms  {S*T/ms*1e3:.3e} samples/s  pairs {pairs:.3e} exact {exact:.3e} ({exact/max(pairs,1):.2%}) {pairs/ms*1e3:.3e} pairs/s',flush=True)
        ms=ev(lambda: ml.kirchhoff_device(x,tt,dist,1.69e8,True))
        print(f'kirch near {S}x{T}: {ms:.2f} ms',flush=True)
if 'stolt' in which:
    for (S,T) in [(2048,8192),(8192,65536)]:
        x=torch.randn(S,T,device='cuda')
        ms=ev(lambda: ml.stolt_device(x,1e-8,5.0,1.68e8,10,10))
        print(f'stolt {S}x{T}: {ms:.3f} ms  {S*T/ms*1e3:.3e} samples/s  {S*T*40/ms*1e3/1e9:.1f} GB/s(40B model)',flush=True)
        del x
if 'filters' in which:
    S,T,B=2048,8192,8
    x=torch.randn(B,S,T,device='cuda'); tp=np.exp(-np.arange(S)*0.01*0.05)
    ms=ev(lambda: fl.horizontalfilt_device(x,'f32',tp,0,T)); print(f'hfilt {B}x{S}x{T}: {ms:.3f} ms {B*S*T*8/ms*1e3/1e9:.1f} GB/s',flush=True)
    ms=ev(lambda: fl.adaptivehfilt_device(x,'f32',tp,1000)); print(f'ahfilt w1000: {ms:.3f} ms {B*S*T*8/ms*1e3/1e9:.1f} GB/s',flush=True)
    from scipy.signal import butter
    b,a=butter(5,[2e6/50e6,10e6/50e6],'bandpass')
    ms=ev(lambda: fl.filtfilt_device(x,'f32',b,a)); print(f'filtfilt B={B}: {ms:.3f} ms {B*S*T*8/ms*1e3/1e9:.1f} GB/s',flush=True)
    x1=x[0].contiguous()
    ms=ev(lambda: fl.filtfilt_device(x1,'f32',b,a)); print(f'filtfilt B=1: {ms:.3f} ms {S*T*8/ms*1e3/1e9:.1f} GB/s',flush=True)
    del x
if 'ahfilt' in which:
    from impdar_b200 import _lib
    lib=_lib.load()
    for (S,T,B) in [(2048,8192,8),(2048,8192,1)]:
        x=torch.randn(B,S,T,device='cuda') if B>1 else torch.randn(S,T,device='cuda')
        tp=np.exp(-np.arange(S)*0.01*0.05)
        for w in (1000,100):
            for mode,name in ((0,'fast'),(2,'strip')):
                lib.impdar_ahfilt_force_rowwise(mode)
                ms=ev(lambda: fl.adaptivehfilt_device(x,'f32',tp,w),n=5,warm=2)
                lib.impdar_ahfilt_force_rowwise(0)
                print(f'ahfilt {B}x{S}x{T} w={w} {name}: {ms:.3f} ms {B*S*T*8/ms*1e3/1e9:.1f} GB/s ({B*S*T*8/ms*1e3/1e9/6552:.1%} of HBM peak)',flush=True)
        del x
if 'phsh' in which:
    for (S,T) in [(1024,2048),(4096,16384)]:
        x=torch.randn(S,T,device='cuda'); tt,dist=geom(S,T)
        ms=ev(lambda: ml.phase_shift_device(x,1e-8,5.0,tt,1.69e8,10,10),n=2)
        print(f'phsh const {S}x{T}: {ms:.2f} ms  {S*T/ms*1e3:.3e} samples/s',flush=True)
        vm=np.linspace(1.69e8,2.2e8,S)
        ms=ev(lambda: ml.phase_shift_device(x,1e-8,5.0,tt,vm,10,10),n=2)
        print(f'phsh layered {S}x{T}: {ms:.2f} ms  {S*T/ms*1e3:.3e} samples/s',flush=True)
if 'kmodes' in which:
    for (S,T) in [(2048,4096),(8192,16384)]:
        x=torch.randn(S,T,device='cuda'); tt,dist=geom(S,T)
        for mode,name in ((1,'general'),(2,'table')):
            ml.set_kirchhoff_mode(mode)
            ml.enable_kirchhoff_stats(True)
            ml.kirchhoff_device(x,tt,dist,1.69e8,False); torch.cuda.synchronize()
            pairs,exact=ml.kirchhoff_stats()
            ml.enable_kirchhoff_stats(False)
            ms=ev(lambda: ml.kirchhoff_device(x,tt,dist,1.69e8,False))
            print(f'kirch[{name}] {S}x{T}: {ms:.2f} ms  {S*T/ms*1e3:.3e} samples/s  pairs {pairs:.3e} exact {exact:.3e} {pairs/ms*1e3:.3e} pairs/s',flush=True)
            ms=ev(lambda: ml.kirchhoff_device(x,tt,dist,1.69e8,True))
            print(f'kirch[{name}] near {S}x{T}: {ms:.2f} ms',flush=True)
        ml.set_kirchhoff_mode(0)

if 'indexops' in which:
    from impdar_b200 import processing as pr
    S,T=4096,16384
    x=torch.randn(S,T,device='cuda'); n=S*T
    ms=ev(lambda: pr.crop_device(x,100,S,50,T)); print(f'crop {S}x{T}: {ms:.3f} ms {(S-100)*(T-50)*8/ms*1e3/1e9:.0f} GB/s',flush=True)
    ms=ev(lambda: pr.crop_device(x,0,S,0,T,True)); print(f'reverse: {ms:.3f} ms {n*8/ms*1e3/1e9:.0f} GB/s',flush=True)
    ms=ev(lambda: pr.restack_device(x,'f32',torch.float32,5)); print(f'restack 5: {ms:.3f} ms {n*4.8/ms*1e3/1e9:.0f} GB/s (4 B read + 0.8 B write per sample)',flush=True)
    tt=np.arange(S)*0.01; nmot=np.sqrt((tt+0.3)**2-0.09); new=np.arange(0,nmot.max(),0.01)
    nodes=pr.linear_nodes_scipy(nmot,new[new>=nmot[0]])
    ms=ev(lambda: pr.interp_rows_device(x,'f32',torch.float32,nodes,0)); print(f'nmo rows ({len(nodes)} out rows): {ms:.3f} ms {len(nodes)*T*8/ms*1e3/1e9:.0f} GB/s (L2 serves the second source row)',flush=True)
    sh=np.random.default_rng(0).integers(0,40,T)
    ms=ev(lambda: pr.shift_traces_device(x,'f32',torch.float32,sh,S)); print(f'shift traces: {ms:.3f} ms {n*8/ms*1e3/1e9:.0f} GB/s',flush=True)
