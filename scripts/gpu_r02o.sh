#!/bin/bash
# usage: gpurun --gpus 8 -- bash scripts/gpu_r02o.sh 8 TAG : peer-mapped output image at 8 GPUs (bitwise check + A/B)
N=${1:-8}; TAG=${2:-r02o}
O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 300 $TR scripts/check_sharded.py 4096 16384 2>&1 | grep -v "^W\|^\*\*\|OMP" | tee $O/${TAG}_check_sharded_n$N.txt
show() { python -c "
import json,sys
d=json.loads(open('$1').read()); r=d['roofline']
print('$2: N=%d ms/step %.2f kernel %.2f ms/step share %.3f by rank %s parity %s e2e %s' % (d['n_gpus'], d['ms_per_step'], r['kernel_ms_per_step'], r['kernel_share_of_step'], r.get('kernel_ms_per_step_by_rank'), (d.get('parity') or {}).get('rel_l2'), (d.get('e2e') or {}).get('ms_per_step')))
print('   ', d['config']['parallelism'][:150], d['clocks'])" | tee -a $O/${TAG}_variants_n$N.txt; }
timeout 300 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-records --no-e2e 2>&1 | grep -v "^W\|^\*\*\|OMP" | tail -1 > $O/bench_northstar_n${N}_$TAG.json
show $O/bench_northstar_n${N}_$TAG.json "default"
IMPDAR_PEER_OUTPUT=0 timeout 300 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-records --no-e2e --no-parity 2>&1 | grep -v "^W\|^\*\*\|OMP" | tail -1 > $O/bench_northstar_n${N}_nopeer_$TAG.json
show $O/bench_northstar_n${N}_nopeer_$TAG.json "no peer image"
for C in 3,3,2,1; do
IMPDAR_C5_CHUNKS=$C timeout 300 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-records --no-e2e --no-parity 2>&1 | grep -v "^W\|^\*\*\|OMP" | tail -1 > $O/tmp.json
show $O/tmp.json "chunks $C"
done
rm -f $O/tmp.json
