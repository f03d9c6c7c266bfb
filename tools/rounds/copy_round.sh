#!/bin/bash
# GPU-box visit after a copy-kernel change: parity tests, per-kernel table, short bench, ncu of the copy paths.
# Usage: bash tools/copy_round.sh <tag>
set -u
TAG=${1:-r01c}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $OUT/${TAG}_pytest.txt
echo "== kernel table"; timeout 900 python tools/kernel_table.py --out $OUT/${TAG}_kernel_table.json > $OUT/${TAG}_kernel_table.log 2>&1; tail -1 $OUT/${TAG}_kernel_table.log | cut -c1-200
echo "== bench"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>$OUT/${TAG}_bench.err | tee $OUT/${TAG}_bench.json | cut -c1-300
echo "== ncu full: copy kernel, P3 (long rows x3 then tiled x3) and P2 (short rows)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:copy_kernel -c 6 -f -o $OUT/${TAG}_copy_p3 \
   python tools/kernel_table.py --names U1_D16384_P3 --dtypes f64 --reps 1 --out $OUT/${TAG}_tmp.json > $OUT/${TAG}_ncu_p3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:copy_kernel -c 3 -f -o $OUT/${TAG}_copy_p2 \
   python tools/kernel_table.py --names U1_D16384_P2 --dtypes f64 --reps 1 --out $OUT/${TAG}_tmp.json > $OUT/${TAG}_ncu_p2.log 2>&1
rm -f $OUT/${TAG}_tmp.json
ls -la $OUT | tail -12
