#!/bin/bash
TAG=${1:-r02c}
O=gpurun_out; mkdir -p $O
echo "== tests"; timeout 1500 python -m pytest tests/test_process.py tests/test_config_shapes.py tests/test_gpu_parity.py -m gpu -q -x -k "process or kirch or sharded" 2>&1 | tail -5 | tee $O/tests_$TAG.log
echo "== tile c2"
timeout 600 python bench.py --workload kirchhoff --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('ms/step %.4f kernel %s %.4f ms pairs/s %.3e parity %.2e' % (d['ms_per_step'], r['kernel'], r['kernel_ms'], r['achieved'], d['parity']['rel_l2']))" | tee -a $O/tile_ab_$TAG.log
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:kirch_tile_kernel -c 1 \
   -f -o $O/full_kirch_tile_$TAG python bench.py --workload kirchhoff --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-parity > $O/full_kirch_tile_$TAG.log 2>&1
ncu -i $O/full_kirch_tile_$TAG.ncu-rep --page raw --csv > $O/full_kirch_tile_$TAG.csv 2>/dev/null
python scripts/ncu_summary.py $O/full_kirch_tile_$TAG.csv $O/traffic_$TAG.json > $O/${TAG}_ncu_full_kirch_tile.txt 2>&1
ncu -i $O/full_kirch_tile_$TAG.ncu-rep --page source --csv > $O/full_kirch_tile_${TAG}_source.csv 2>/dev/null
rm -f $O/full_kirch_tile_$TAG.ncu-rep
