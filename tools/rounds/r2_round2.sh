#!/bin/bash
# Round 2, GPU call 2: parity after the skinny / run-table / zero-fill / workspace changes, kernel table, ncu of the small-launch
# GEMMs (U1xU1) and of the run-table copy path.
set -u
TAG=${1:-r02b}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu (without the reference suite)"; timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_reference_suite_gpu.py 2>&1 | tail -25 | tee $OUT/${TAG}_pytest.txt
echo "== kernel table"; timeout 600 python tools/kernel_table.py --reps 7 --names U1_D16384_P3 U1_D4096_P3 U1xU1_D4096_P1 U1xU1_D4096_P2 U1xU1_D4096_P3 Z2_D512_P2 U1_D16384_P1 U1_D1024_P1 U1_D4096_T1 --out $OUT/${TAG}_kernel_table.json > $OUT/${TAG}_kernel_table.log 2>&1; tail -1 $OUT/${TAG}_kernel_table.log | cut -c1-200
echo "== bench f64 (no dmrg / baselines)"; timeout 600 python bench.py --no-dmrg --no-cpu-baseline --no-gpu-baseline 2>$OUT/${TAG}_bench.err | tee $OUT/${TAG}_bench_f64.json | cut -c1-300
echo "== ncu"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'gemm_kernel' -s 2 -c 1 -f -o $OUT/${TAG}_gemm_u1u1_p1 \
   python tools/kernel_table.py --names U1xU1_D4096_P1 --dtypes f64 --reps 1 --out $OUT/${TAG}_tmp.json > $OUT/${TAG}_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'gemm_kernel' -s 2 -c 1 -f -o $OUT/${TAG}_gemm_u1u1_p2 \
   python tools/kernel_table.py --names U1xU1_D4096_P2 --dtypes f64 --reps 1 --out $OUT/${TAG}_tmp.json > $OUT/${TAG}_ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'skinny_kernel' -s 2 -c 1 -f -o $OUT/${TAG}_skinny_p3 \
   python tools/kernel_table.py --names U1_D16384_P3 --dtypes f64 --reps 1 --out $OUT/${TAG}_tmp.json > $OUT/${TAG}_ncu3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'copy_kernel|tiled_kernel' -s 2 -c 1 -f -o $OUT/${TAG}_copy_u1u1 \
   python tools/kernel_table.py --names U1xU1_D4096_P1 --dtypes f64 --reps 1 --out $OUT/${TAG}_tmp.json > $OUT/${TAG}_ncu4.log 2>&1
rm -f $OUT/${TAG}_tmp.json
ls -la $OUT | grep ${TAG}
