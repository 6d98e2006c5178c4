#!/bin/bash
# Round-2 first GPU pass: new parity / boundary tests, the tile kernel A/B, the north-star bench line.
TAG=${1:-r02a}
O=gpurun_out
mkdir -p $O
echo "== config-shape + boundary tests"; timeout 1500 python -m pytest tests/test_config_shapes.py -m gpu -x -q -s 2>&1 | tail -40 | tee $O/tests_cfg_$TAG.log
echo "== all gpu tests"; timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_config_shapes.py 2>&1 | tail -8 | tee $O/tests_$TAG.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $O/smoke_$TAG.log
echo "== kirchhoff c2 A/B"
for mode in 3 2; do
  IMPDAR_KIRCH_MODE=$mode timeout 600 python bench.py --workload kirchhoff --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee $O/bench_kirchhoff_mode${mode}_$TAG.json
done
echo "== north star"; timeout 1200 python bench.py --steps 5 --warmup 3 2>&1 | tail -1 | tee $O/bench_northstar_$TAG.json
echo "== ncu tile"
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:kirch_tile_kernel -c 1 \
   -f -o $O/full_kirch_tile_$TAG python bench.py --workload kirchhoff --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-parity > $O/full_kirch_tile_$TAG.log 2>&1
ncu -i $O/full_kirch_tile_$TAG.ncu-rep --page raw --csv > $O/full_kirch_tile_$TAG.csv 2>/dev/null
python scripts/ncu_summary.py $O/full_kirch_tile_$TAG.csv $O/traffic_$TAG.json > $O/${TAG}_ncu_full_kirch_tile.txt 2>&1
rm -f $O/full_kirch_tile_$TAG.ncu-rep
ls -la $O | tail -20
