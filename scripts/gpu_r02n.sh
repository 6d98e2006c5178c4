#!/bin/bash
# usage: gpurun --gpus N -- bash scripts/gpu_r02n.sh N TAG : peer-mapped output image - bitwise check, timings, north-star line
N=${1:-2}; TAG=${2:-r02n}
O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 400 $TR scripts/check_sharded.py 4096 16384 2>&1 | grep -v "^W\|^\*\*\|OMP" | tee $O/${TAG}_check_sharded_n$N.txt
timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-records 2>&1 | grep -v "^W\|^\*\*\|OMP" | tail -1 > $O/bench_northstar_n${N}_$TAG.json
python - <<P
import json
d=json.loads(open("$O/bench_northstar_n${N}_$TAG.json").read()); r=d['roofline']
print('N=%d ms/step %.2f kernel %.2f ms/step share %.3f parity %s e2e %.1f ms' % (d['n_gpus'], d['ms_per_step'], r['kernel_ms_per_step'], r['kernel_share_of_step'], d['parity'].get('rel_l2'), d['e2e']['ms_per_step']))
print(d['config']['parallelism'])
P
IMPDAR_PEER_OUTPUT=0 timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-records --no-e2e --no-parity 2>&1 | grep -v "^W\|^\*\*\|OMP" | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('no peer image: N=%d ms/step %.2f kernel %.2f ms/step share %.3f' % (d['n_gpus'], d['ms_per_step'], r['kernel_ms_per_step'], r['kernel_share_of_step']))"
