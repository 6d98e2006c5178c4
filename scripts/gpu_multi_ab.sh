#!/bin/bash
# usage: gpurun --gpus N -- bash scripts/gpu_multi_ab.sh N TAG : exchange-pipeline variants of the north-star step
N=${1:-8}; TAG=${2:-r02k}
O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
run() { # label, env...
  local label=$1; shift
  timeout 600 env "$@" $TR bench.py --gpus $N --steps 8 --warmup 3 --no-e2e --no-records --no-parity 2>&1 | grep -v "^W\|^\*\*\|OMP" | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$label: N=%d ms/step %.2f kernel %.2f ms/step share %.3f' % (d['n_gpus'], d['ms_per_step'], r['kernel_ms_per_step'], r['kernel_share_of_step']))" | tee -a $O/${TAG}_exchange_variants_n$N.txt
}
run "4 equal chunks" X=1
run "chunks 1,2,2,1" IMPDAR_C5_CHUNKS=1,2,2,1
run "chunks 1,2,2,2,1" IMPDAR_C5_CHUNKS=1,2,2,2,1
run "chunks 1,2,3,3,2,1" IMPDAR_C5_CHUNKS=1,2,3,3,2,1
