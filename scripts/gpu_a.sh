#!/bin/bash
# r01e step A: parity tests + smoke on HEAD, Stolt column kernel with / without the TMA prefetch
O=gpurun_out; mkdir -p $O
echo "== tests"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $O/tests_r01e.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $O/smoke_r01e.log
echo "== stolt PF=1"; timeout 300 python scripts/quick_gpu.py stolt 2>&1 | tee $O/quick_stolt_pf1.log
echo "== stolt PF=0"; IMPDAR_STOLT_COL_PREFETCH=0 timeout 300 python scripts/quick_gpu.py stolt 2>&1 | tee $O/quick_stolt_pf0.log
echo "== bench stolt"; timeout 600 python bench.py --workload stolt --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee $O/bench_stolt_r01e_pf1.json
