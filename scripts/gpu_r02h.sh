#!/bin/bash
TAG=${1:-r02h}
O=gpurun_out; mkdir -p $O
echo "== new test"; timeout 600 python -m pytest tests/test_config_shapes.py -m gpu -q -x -k "deterministic" 2>&1 | tail -3 | tee $O/tests_det_$TAG.log
echo "== driver-style bench"; timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 2>&1 | tail -1 | tee $O/bench_northstar_$TAG.json | cut -c1-600
echo "== reference arm"; timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 2>&1 | tail -1 | tee $O/bench_reference_$TAG.json | cut -c1-400
for wl in phsh phsh_layered pipeline stolt_c4; do
  echo "== bench $wl"; timeout 900 python bench.py --workload $wl --steps 10 --warmup 3 2>&1 | tail -1 | tee $O/bench_${wl}_$TAG.json | cut -c1-300
done
echo "== launch list north star"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off -c 400 --csv \
   --log-file $O/launches_northstar_$TAG.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-parity --no-records > $O/launches_northstar_$TAG.log 2>&1
python scripts/launch_summary.py $O/launches_northstar_$TAG.csv > $O/${TAG}_launches_northstar.txt 2>&1
cap() { # name workload regex extra-env
  timeout 900 env $4 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$3 -c 1 \
     -f -o $O/full_$1_$TAG python bench.py --workload $2 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-parity --no-records > $O/full_$1_$TAG.log 2>&1
  ncu -i $O/full_$1_$TAG.ncu-rep --page raw --csv > $O/full_$1_$TAG.csv 2>/dev/null
  python scripts/ncu_summary.py $O/full_$1_$TAG.csv $O/traffic_$TAG.json > $O/${TAG}_ncu_full_$1.txt 2>&1
  rm -f $O/full_$1_$TAG.ncu-rep
}
cap kirch_tile kirchhoff kirch_tile_kernel X=1
cap kirch_tile_c5 kirchhoff_c5 kirch_tile_kernel X=1
cap kirch_general kirchhoff kirch_general_kernel IMPDAR_KIRCH_MODE=1
cap phsh_const_tc phsh phsh_const_tc_kernel X=1
cap stolt_col stolt stolt_col_kernel X=1
ls $O | tail -30
