#!/bin/bash
# usage: gpurun --gpus N -- bash scripts/gpu_r02p.sh N TAG [check] : peer-mapped windows + image (bitwise check, A/B of the exchange variants)
N=${1:-2}; TAG=${2:-r02p}; CHECK=${3:-check}
O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
[ "$CHECK" = check ] && timeout 300 $TR scripts/check_sharded.py 4096 16384 2>&1 | grep -v "^W\|^\*\*\|OMP" | tee $O/${TAG}_check_sharded_n$N.txt
show() { python -c "
import json,sys
d=json.loads(open('$1').read()); r=d['roofline']
print('$2: N=%d ms/step %.2f kernel %.2f ms/step share %.3f by rank %s parity %s e2e %s' % (d['n_gpus'], d['ms_per_step'], r['kernel_ms_per_step'], r['kernel_share_of_step'], r.get('kernel_ms_per_step_by_rank'), (d.get('parity') or {}).get('rel_l2'), (d.get('e2e') or {}).get('ms_per_step')))
print('   ', d['config']['parallelism'][:190], d['clocks'])" | tee -a $O/${TAG}_variants_n$N.txt; }
timeout 300 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-records --no-e2e 2>&1 | grep -v "^W\|^\*\*\|OMP" | tail -1 > $O/bench_northstar_n${N}_$TAG.json
show $O/bench_northstar_n${N}_$TAG.json "default"
shift 3
for V in "$@"; do
env $V timeout 300 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-records --no-e2e --no-parity 2>&1 | grep -v "^W\|^\*\*\|OMP" | tail -1 > $O/tmp.json
show $O/tmp.json "$V"
done
rm -f $O/tmp.json
