#!/bin/bash
# quick Stolt iteration: stage tests, timings, launch list.  Usage: bash scripts/gpu_stolt.sh TAG [full]
TAG=${1:-x}
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_stolt_stages.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python scripts/quick_gpu.py stolt 2>&1 | tail -3
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none --profile-from-start off -c 100 --csv --log-file $O/launches_stolt_$TAG.csv python bench.py --workload stolt --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none --profile-from-start off -c 100 --csv --log-file $O/launches_stolt_c4_$TAG.csv python bench.py --workload stolt_c4 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
if [ "$2" == "full" ]; then
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:stolt_ -c 5 -f -o $O/full_stolt5_$TAG python bench.py --workload stolt --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $O/full_stolt5_$TAG.log 2>&1
fi
