#!/bin/bash
# One gpurun call: parity tests, smoke, every bench workload, ncu launch lists, --set full captures.
# Usage (from the repo root on the GPU box):  bash scripts/gpu_round.sh [tag] [stages]
#   stages: any of  tests bench quick launches full   (default: all)
TAG=${1:-r01}
STAGES=${2:-"tests bench quick launches full"}
O=gpurun_out
KEEP_REP=${KEEP_REP:-"stolt_col"}
mkdir -p $O
has() { [[ " $STAGES " == *" $1 "* ]]; }
WLS="kirchhoff stolt stolt_c4 pipeline phsh phsh_layered"
if has tests; then
  echo "== tests"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $O/tests_$TAG.log
  echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $O/smoke_$TAG.log
fi
if has bench; then
  for wl in $WLS; do
    echo "== bench $wl"
    timeout 900 python bench.py --workload $wl --steps 10 --warmup 3 2>&1 | tail -3 | tee $O/bench_${wl}_$TAG.json
  done
  echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -2 | tee $O/bench_reference_$TAG.json
fi
if has quick; then
  echo "== quick"; timeout 600 python scripts/quick_gpu.py kirch stolt filters phsh kmodes indexops 2>&1 | tee $O/quick_$TAG.log
fi
if has launches; then
  for wl in $WLS; do
    timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
       --profile-from-start off -c 400 --csv --log-file $O/launches_${wl}_$TAG.csv \
       python bench.py --workload $wl --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $O/launches_${wl}_$TAG.log 2>&1
    python scripts/launch_summary.py $O/launches_${wl}_$TAG.csv > $O/${TAG}_launches_${wl}.txt 2>&1
  done
fi
# full captures of the dominant kernel(s) per workload; raw pages exported here (the .ncu-rep stays in gpurun_out)
cap() { # name workload regex
  timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$3 -c 1 \
     -f -o $O/full_$1_$TAG python bench.py --workload $2 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $O/full_$1_$TAG.log 2>&1
  ncu -i $O/full_$1_$TAG.ncu-rep --page raw --csv > $O/full_$1_$TAG.csv 2>/dev/null
  python scripts/ncu_summary.py $O/full_$1_$TAG.csv $O/traffic_$TAG.json > $O/${TAG}_ncu_full_$1.txt 2>&1
  ncu -i $O/full_$1_$TAG.ncu-rep --page source --csv > $O/full_$1_${TAG}_source.csv 2>/dev/null
  # gpurun copies back at most 64 MiB: keep the reports listed in KEEP_REP only
  [[ " $KEEP_REP " == *" $1 "* ]] || rm -f $O/full_$1_$TAG.ncu-rep
}
if has full; then
  cap kirch_table kirchhoff kirch_table_kernel
  cap stolt_col stolt stolt_col_kernel
  cap stolt_rowA stolt stolt_rowA_kernel
  cap stolt_rowB stolt 'stolt_rowB_kernel'
  cap filtfilt pipeline filtfilt_kernel
  cap hfilt pipeline hfilt_kernel
  cap phsh_const_pair phsh phsh_const_pair_kernel
  cap phsh_layered_pair phsh_layered phsh_layered_pair_kernel
fi
ls -la $O
