#!/bin/bash
TAG=${1:-r02d}
O=gpurun_out; mkdir -p $O
echo "== failing test"; timeout 600 python -m pytest tests/test_process.py -m gpu -q -x -k failing 2>&1 | tail -40 | tee $O/tests_fail_$TAG.log
echo "== kirch tests"; timeout 1200 python -m pytest tests/test_config_shapes.py tests/test_gpu_parity.py -m gpu -q -x -k "kirch or sharded" 2>&1 | tail -4 | tee $O/tests_$TAG.log
echo "== tile variants c2"
for lib in "" np2 np8 nw14rw2; do
  if [ -n "$lib" ]; then export IMPDAR_B200_LIB=$PWD/impdar_b200/libimpdar_b200_$lib.so; else unset IMPDAR_B200_LIB; fi
  timeout 600 python bench.py --workload kirchhoff --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('variant [$lib] ms/step %.4f kernel %s %.4f ms pairs/s %.3e parity %.2e' % (d['ms_per_step'], r['kernel'], r['kernel_ms'], r['achieved'], d['parity']['rel_l2']))" | tee -a $O/tile_ab_$TAG.log
done
unset IMPDAR_B200_LIB
echo "== c5"
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-records 2>&1 | tail -1 | tee $O/bench_c5_$TAG.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('C5 ms/step %.2f kernel %s %.2f ms pairs/s %.3e parity %.2e' % (d['ms_per_step'], r['kernel'], r['kernel_ms_per_step'], r['achieved'], d['parity']['rel_l2']))" | tee -a $O/tile_ab_$TAG.log
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:kirch_tile_kernel -c 1 \
   -f -o $O/full_kirch_tile_$TAG python bench.py --workload kirchhoff --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-parity > $O/full_kirch_tile_$TAG.log 2>&1
ncu -i $O/full_kirch_tile_$TAG.ncu-rep --page raw --csv > $O/full_kirch_tile_$TAG.csv 2>/dev/null
python scripts/ncu_summary.py $O/full_kirch_tile_$TAG.csv $O/traffic_$TAG.json > $O/${TAG}_ncu_full_kirch_tile.txt 2>&1
ncu -i $O/full_kirch_tile_$TAG.ncu-rep --page source --csv > $O/full_kirch_tile_${TAG}_source.csv 2>/dev/null
rm -f $O/full_kirch_tile_$TAG.ncu-rep
