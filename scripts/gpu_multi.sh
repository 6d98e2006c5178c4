#!/bin/bash
# Multi-GPU round: the driver's own launch line at N GPUs for the default workload, the C4 pipeline and the
# C5 Kirchhoff (NCCL broadcast + all-gather).  Usage: bash scripts/gpu_multi.sh TAG N
TAG=${1:-r01f}; N=${2:-2}
O=gpurun_out; mkdir -p $O
run() { # workload extra...
  wl=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --workload $wl "$@" 2>&1 | tail -4 | tee $O/bench_${wl}_n${N}_$TAG.json
}
run kirchhoff --steps 10 --warmup 3 --no-cpu-baseline
run pipeline --steps 5 --warmup 3 --no-cpu-baseline
run kirchhoff_c5 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --impl reference --gpus $N --steps 1 --warmup 1 2>&1 | tail -2 | tee $O/bench_reference_n${N}_$TAG.json
