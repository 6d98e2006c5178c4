#!/bin/bash
# usage: gpurun --gpus N -- bash scripts/gpu_multi_r02.sh N TAG
N=${1:-2}; TAG=${2:-r02}
O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
echo "== check_sharded N=$N"; timeout 900 $TR scripts/check_sharded.py 8192 32768 2>&1 | grep -v "^W\|^\*\*\|OMP" | tee $O/${TAG}_check_sharded_n$N.txt
echo "== bench north star N=$N"; timeout 1500 $TR bench.py --gpus $N --steps 5 --warmup 3 2>&1 | grep -v "^W\|^\*\*\|OMP" | tail -1 | tee $O/bench_northstar_n${N}_$TAG.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('N=%d value %.3e ms/step %.2f kernel %s %.2f ms/step, share %.3f parity %s e2e %s' % (d['n_gpus'], d['value'], d['ms_per_step'], r['kernel'], r['kernel_ms_per_step'], r['kernel_share_of_step'], d['parity'], d['e2e']))"
