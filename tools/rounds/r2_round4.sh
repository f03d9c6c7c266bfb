#!/bin/bash
# Round 2, GPU call: warp-specialised GEMM — smoke first (short timeout: a protocol bug would hang), then parity, A/B kernel table.
set -u
TAG=${1:-r02h}
OUT=gpurun_out
mkdir -p $OUT
echo "== smoke (ws kernel)"; timeout 120 python __graft_entry__.py --smoke 2>&1 | tail -3 | tee $OUT/${TAG}_smoke.txt
if ! grep -q "smoke complex128" $OUT/${TAG}_smoke.txt; then echo "SMOKE FAILED - stopping"; exit 1; fi
echo "== pytest gemm-related"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_skinny.py -m gpu -q -x 2>&1 | tail -8 | tee $OUT/${TAG}_pytest.txt
echo "== kernel table WS"; timeout 400 python tools/kernel_table.py --reps 7 --names U1xU1_D4096_P1 U1xU1_D4096_P2 Z2_D512_P2 U1_D16384_P1 U1_D16384_P2 U1_D1024_P1 U1_D4096_P1 U1_D4096_T1 --out $OUT/${TAG}_kt_ws.json > $OUT/${TAG}_kt_ws.log 2>&1; tail -1 $OUT/${TAG}_kt_ws.log | cut -c1-100
echo "== kernel table classic"; YB_GEMM_CLASSIC=1 timeout 400 python tools/kernel_table.py --reps 7 --names U1xU1_D4096_P1 U1xU1_D4096_P2 Z2_D512_P2 U1_D16384_P1 U1_D16384_P2 U1_D1024_P1 U1_D4096_P1 U1_D4096_T1 --out $OUT/${TAG}_kt_classic.json > $OUT/${TAG}_kt_classic.log 2>&1; tail -1 $OUT/${TAG}_kt_classic.log | cut -c1-100
echo "== bench f64 WS"; timeout 600 python bench.py --no-dmrg --no-cpu-baseline --no-gpu-baseline 2>$OUT/${TAG}_bench.err | tee $OUT/${TAG}_bench_ws.json | cut -c1-300
echo "== bench f64 classic"; YB_GEMM_CLASSIC=1 timeout 600 python bench.py --no-dmrg --no-cpu-baseline --no-gpu-baseline --no-e2e 2>>$OUT/${TAG}_bench.err | tee $OUT/${TAG}_bench_classic.json | cut -c1-300
echo "== ncu WS gemm U1xU1 P1 + D16384 P1"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'gemm_ws_kernel' -s 2 -c 1 -f -o $OUT/${TAG}_ws_u1u1_p1 \
   python tools/kernel_table.py --names U1xU1_D4096_P1 --dtypes f64 --reps 1 --out $OUT/${TAG}_tmp.json > $OUT/${TAG}_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'gemm_ws_kernel' -s 2 -c 1 -f -o $OUT/${TAG}_ws_d16384_p1 \
   python tools/kernel_table.py --names U1_D16384_P1 --dtypes f64 --reps 1 --out $OUT/${TAG}_tmp.json > $OUT/${TAG}_ncu2.log 2>&1
rm -f $OUT/${TAG}_tmp.json
ls $OUT | grep ${TAG}
