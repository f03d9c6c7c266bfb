#!/bin/bash
# Round 2, GPU call 3: parity incl. elementwise engine; kernel table; quick bench; DMRG configs 1 and 3 e2e.
set -u
TAG=${1:-r02c}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu (without the reference suite)"; timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_reference_suite_gpu.py 2>&1 | tail -25 | tee $OUT/${TAG}_pytest.txt
echo "== kernel table"; timeout 600 python tools/kernel_table.py --reps 7 --names U1_D16384_P3 U1_D4096_P3 U1xU1_D4096_P1 U1xU1_D4096_P2 U1xU1_D4096_P3 U1_D16384_P1 U1_D1024_P1 U1_D4096_T1 --out $OUT/${TAG}_kernel_table.json > $OUT/${TAG}_kernel_table.log 2>&1; tail -1 $OUT/${TAG}_kernel_table.log | cut -c1-200
echo "== bench f64 (no dmrg / baselines)"; timeout 600 python bench.py --no-dmrg --no-cpu-baseline --no-gpu-baseline 2>$OUT/${TAG}_bench.err | tee $OUT/${TAG}_bench_f64.json | cut -c1-300
E2E=$OUT/${TAG}_e2e.jsonl; rm -f $E2E
echo "== DMRG config 1 (b200)"; timeout 300 python tools/dmrg_bench.py --model heisenberg --N 32 --D 64 --sweeps 3 --backend b200 --out $E2E 2>&1 | tail -1 | cut -c1-300
echo "== DMRG config 1 profile"; timeout 300 python tools/dmrg_bench.py --model heisenberg --N 32 --D 64 --sweeps 3 --backend b200 --profile --out $E2E 2>&1 | tail -1 | cut -c1-1500
echo "== DMRG Hubbard N=20 D=4096 c128 (b200 fused, gemm roofline)"; timeout 900 python tools/dmrg_bench.py --model hubbard --N 20 --D 4096 --D0 4096 --sweeps 1 --backend b200 --fused --gemm-roofline --dtype complex128 --out $E2E 2>&1 | tail -1 | cut -c1-1200
echo "== same with profile"; timeout 900 python tools/dmrg_bench.py --model hubbard --N 20 --D 4096 --D0 4096 --sweeps 1 --backend b200 --fused --profile --dtype complex128 --out $E2E 2>&1 | tail -1 | cut -c1-1800
ls -la $OUT | grep ${TAG}
