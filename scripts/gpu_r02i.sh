#!/bin/bash
TAG=${1:-r02i}
O=gpurun_out; mkdir -p $O
echo "== phsh tests"; timeout 900 python -m pytest tests/test_config_shapes.py tests/test_gpu_parity.py -m gpu -q -x -k "phase or phsh" 2>&1 | tail -3 | tee $O/tests_$TAG.log
for lib in "" tc16; do
  if [ -n "$lib" ]; then export IMPDAR_B200_LIB=$PWD/impdar_b200/libimpdar_b200_$lib.so; else unset IMPDAR_B200_LIB; fi
  echo "== variant [$lib]"; timeout 600 python scripts/diag_phsh_tc.py 2>&1 | grep -v "Phase-Shift" | tee -a $O/diag_phsh_tc_$TAG.log
done
unset IMPDAR_B200_LIB
timeout 600 python bench.py --workload phsh --steps 10 --warmup 3 2>&1 | tail -1 | tee $O/bench_phsh_$TAG.json | cut -c1-900
