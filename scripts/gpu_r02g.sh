#!/bin/bash
TAG=${1:-r02g}
O=gpurun_out; mkdir -p $O
echo "== all gpu tests"; timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -12 | tee $O/tests_$TAG.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $O/smoke_$TAG.log
echo "== ahfilt timing"; timeout 600 python scripts/quick_gpu.py ahfilt 2>&1 | tee $O/${TAG}_quick_ahfilt.txt
echo "== sanitizer memcheck"; timeout 1200 compute-sanitizer --tool memcheck --print-limit 20 python scripts/sanitize_r02.py 2>&1 | tail -25 | tee $O/${TAG}_sanitizer_memcheck.txt
echo "== sanitizer racecheck"; timeout 1500 compute-sanitizer --tool racecheck --print-limit 20 python scripts/sanitize_r02.py 2>&1 | tail -25 | tee $O/${TAG}_sanitizer_racecheck.txt
echo "== sanitizer synccheck"; timeout 900 compute-sanitizer --tool synccheck --print-limit 20 python scripts/sanitize_r02.py 2>&1 | tail -15 | tee $O/${TAG}_sanitizer_synccheck.txt
