#!/bin/bash
TAG=${1:-r02l}
O=gpurun_out; mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:phsh_const_tc_kernel -c 1 \
   -f -o $O/full_phsh_tc_$TAG python bench.py --workload phsh --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-parity > $O/full_phsh_tc_$TAG.log 2>&1
ncu -i $O/full_phsh_tc_$TAG.ncu-rep --page raw --csv > $O/full_phsh_tc_$TAG.csv 2>/dev/null
python scripts/ncu_summary.py $O/full_phsh_tc_$TAG.csv $O/traffic_$TAG.json > $O/${TAG}_ncu_full_phsh_const_tc.txt 2>&1
ncu -i $O/full_phsh_tc_$TAG.ncu-rep --page source --csv > $O/full_phsh_tc_${TAG}_source.csv 2>/dev/null
rm -f $O/full_phsh_tc_$TAG.ncu-rep
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "rangegain or gains" 2>&1 | tail -2 | tee $O/tests_$TAG.log
