#!/bin/bash
set -u
TAG=${1:-r02i}
OUT=gpurun_out
mkdir -p $OUT
echo "== smoke"; timeout 120 python __graft_entry__.py --smoke 2>&1 | tail -3 | tee $OUT/${TAG}_smoke.txt
if ! grep -q "smoke complex128" $OUT/${TAG}_smoke.txt; then echo "SMOKE FAILED - stopping"; exit 1; fi
echo "== pytest gemm-related"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_skinny.py -m gpu -q -x 2>&1 | tail -8 | tee $OUT/${TAG}_pytest.txt
NAMES="U1xU1_D4096_P1 U1xU1_D4096_P2 Z2_D512_P2 U1_D16384_P1 U1_D16384_P2 U1_D1024_P1 U1_D4096_P1 U1_D4096_P2"
echo "== kernel table skip"; timeout 400 python tools/kernel_table.py --reps 7 --names $NAMES --out $OUT/${TAG}_kt_skip.json > $OUT/${TAG}_kt_skip.log 2>&1; tail -1 $OUT/${TAG}_kt_skip.log | cut -c1-100
echo "== kernel table noskip"; YB_GEMM_NOSKIP=1 timeout 400 python tools/kernel_table.py --reps 7 --names $NAMES --out $OUT/${TAG}_kt_noskip.json > $OUT/${TAG}_kt_noskip.log 2>&1; tail -1 $OUT/${TAG}_kt_noskip.log | cut -c1-100
echo "== bench f64"; timeout 600 python bench.py --no-dmrg --no-cpu-baseline --no-gpu-baseline --no-e2e 2>$OUT/${TAG}_bench.err | tee $OUT/${TAG}_bench.json | cut -c1-300
echo "== bench c128"; timeout 600 python bench.py --dtype c128 --steps 5 --no-dmrg --no-cpu-baseline --no-gpu-baseline --no-e2e 2>>$OUT/${TAG}_bench.err | tee $OUT/${TAG}_bench_c128.json | cut -c1-300
echo "== dmrg"; timeout 400 python tools/dmrg_bench.py --model hubbard --N 20 --D 4096 --D0 4096 --sweeps 1 --backend b200 --fused --gemm-roofline --dtype complex128 --out $OUT/${TAG}_e2e.jsonl 2>&1 | tail -1 | cut -c1-200
