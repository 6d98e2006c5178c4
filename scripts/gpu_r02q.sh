#!/bin/bash
# One GPU: full -m gpu suite, smoke, the default bench line, launch list of the headline step
TAG=${1:-r02q}
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee $O/${TAG}_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 | tee $O/${TAG}_smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 2>&1 | grep -v "^W\|OMP" | tail -1 > $O/bench_northstar_$TAG.json
python -c "
import json
d=json.loads(open('$O/bench_northstar_$TAG.json').read()); r=d['roofline']
print('N=1 ms/step %.2f kernel %.2f frac %.3f parity %s e2e %.1f ms; cpu %s' % (d['ms_per_step'], r['kernel_ms_per_step'], r['frac'], d['parity']['rel_l2'], d['e2e']['ms_per_step'], d['cpu_baseline']['value']))
for k,v in d['records'].items(): print(k, v['ms_per_step'], v['roofline']['frac'], v['parity']['rel_l2'], v['e2e']['ms_per_step'])
print(d['clocks'])"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
   --profile-from-start off -c 400 --csv --log-file $O/launches_northstar_$TAG.csv \
   python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-records --no-parity > $O/launches_northstar_$TAG.log 2>&1
python scripts/launch_summary.py $O/launches_northstar_$TAG.csv | tee $O/${TAG}_launches_northstar.txt
