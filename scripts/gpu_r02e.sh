#!/bin/bash
TAG=${1:-r02e}
O=gpurun_out; mkdir -p $O
echo "== ahfilt + process tests"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_process.py -m gpu -q -x -k "ahfilt or process" 2>&1 | tail -6 | tee $O/tests_$TAG.log
echo "== ahfilt timing"; timeout 600 python scripts/quick_gpu.py ahfilt 2>&1 | tee $O/${TAG}_quick_ahfilt.txt
echo "== tile variants c2"
for lib in "" sleep32 sleep100 np3 flush16 w8n8; do
  if [ -n "$lib" ]; then export IMPDAR_B200_LIB=$PWD/impdar_b200/libimpdar_b200_$lib.so; else unset IMPDAR_B200_LIB; fi
  timeout 600 python bench.py --workload kirchhoff --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('variant [$lib] ms/step %.4f kernel %s %.4f ms pairs/s %.3e parity %.2e' % (d['ms_per_step'], r['kernel'], r['kernel_ms'], r['achieved'], d['parity']['rel_l2']))" | tee -a $O/tile_ab_$TAG.log
done
unset IMPDAR_B200_LIB
echo "== ncu ahfilt"
cat > /tmp/ah_run.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from impdar_b200 import filtering as fl
S,T,B=2048,8192,8
x=torch.randn(B,S,T,device='cuda'); tp=np.exp(-np.arange(S)*0.01*0.05)
for _ in range(3): fl.adaptivehfilt_device(x,'f32',tp,1000)
torch.cuda.synchronize()
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ahfilt_fast_kernel -s 2 -c 1 -f -o $O/full_ahfilt_fast_$TAG python /tmp/ah_run.py > $O/full_ahfilt_fast_$TAG.log 2>&1
ncu -i $O/full_ahfilt_fast_$TAG.ncu-rep --page raw --csv > $O/full_ahfilt_fast_$TAG.csv 2>/dev/null
python scripts/ncu_summary.py $O/full_ahfilt_fast_$TAG.csv $O/traffic_$TAG.json > $O/${TAG}_ncu_full_ahfilt_fast.txt 2>&1
ncu -i $O/full_ahfilt_fast_$TAG.ncu-rep --page source --csv > $O/full_ahfilt_fast_${TAG}_source.csv 2>/dev/null
rm -f $O/full_ahfilt_fast_$TAG.ncu-rep
