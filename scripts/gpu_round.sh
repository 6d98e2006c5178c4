#!/bin/bash
# One gpurun call: parity tests, smoke, every bench workload, ncu launch lists, --set full captures.
# Usage (from the repo root on the GPU box):  bash scripts/gpu_round.sh [tag]
TAG=${1:-r01}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 500 > $O/clocks_$TAG.csv &
SMI=$!
echo "== tests"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $O/tests_$TAG.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $O/smoke_$TAG.log
for wl in kirchhoff stolt stolt_c4 pipeline phsh phsh_layered; do
  echo "== bench $wl"
  timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 2>&1 | tail -3 | tee $O/bench_${wl}_$TAG.json
done
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -2 | tee $O/bench_reference_$TAG.json
echo "== quick"; timeout 600 python scripts/quick_gpu.py kirch stolt filters phsh kmodes 2>&1 | tee $O/quick_$TAG.log
for wl in kirchhoff stolt stolt_c4 pipeline phsh phsh_layered; do
  timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
     --profile-from-start off -c 400 --csv --log-file $O/launches_${wl}_$TAG.csv \
     python bench.py --workload $wl --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $O/launches_${wl}_$TAG.log 2>&1
done
# full captures of the dominant kernel per workload
cap() { # name workload regex
  timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$3 -c 1 \
     -f -o $O/full_$1_$TAG python bench.py --workload $2 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $O/full_$1_$TAG.log 2>&1
}
cap kirch_table kirchhoff kirch_table_kernel
cap stolt_remap stolt stolt_remap
cap filtfilt pipeline filtfilt_kernel
cap hfilt pipeline hfilt_kernel
cap phsh_const phsh phsh_const_kernel
kill $SMI
ls -la $O
