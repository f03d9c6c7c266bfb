#!/bin/bash
# One GPU-box visit: parity tests, bench line, per-kernel table, ncu launch list, ncu --set full captures.
# Usage (from the repo root on the box): bash tools/gpu_round.sh <tag>
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $OUT/${TAG}_pytest.txt
echo "== smoke"; timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -5 | tee $OUT/${TAG}_smoke.txt
echo "== bench"; timeout 900 python bench.py 2>$OUT/${TAG}_bench.err | tee $OUT/${TAG}_bench.json
echo "== kernel table"; timeout 900 python tools/kernel_table.py --out $OUT/${TAG}_kernel_table.json > $OUT/${TAG}_kernel_table.log 2>&1; tail -3 $OUT/${TAG}_kernel_table.log
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'copy_kernel|gemm_kernel|match' -c 400 --csv \
   --log-file $OUT/${TAG}_launches.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/${TAG}_launches_bench.log 2>&1
echo "== ncu full: gemm"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 6 -c 2 -f -o $OUT/${TAG}_gemm \
   python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --sizes 16384 > $OUT/${TAG}_ncu_gemm.log 2>&1
echo "== ncu full: copy"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:copy_kernel -s 12 -c 4 -f -o $OUT/${TAG}_copy \
   python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --sizes 16384 > $OUT/${TAG}_ncu_copy.log 2>&1
ls -la $OUT
