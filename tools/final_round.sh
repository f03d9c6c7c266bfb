#!/bin/bash
# Round-end evidence on one GPU box: parity, bench lines, per-kernel table, ncu launch list + full captures, host overhead,
# end-to-end targets (BASELINE.json configs 3 and 4) with our backend and with the stock torch backend on the same GPU.
# Usage: bash tools/final_round.sh <tag>
set -u
TAG=${1:-r01f}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
nproc > $OUT/${TAG}_nproc.txt
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee $OUT/${TAG}_pytest.txt
echo "== smoke"; timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3 | tee $OUT/${TAG}_smoke.txt
echo "== bench f64"; timeout 900 python bench.py 2>$OUT/${TAG}_bench.err | tee $OUT/${TAG}_bench_f64.json | cut -c1-300
echo "== bench c128"; timeout 900 python bench.py --dtype c128 --steps 5 --no-cpu-baseline 2>>$OUT/${TAG}_bench.err | tee $OUT/${TAG}_bench_c128.json | cut -c1-300
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>>$OUT/${TAG}_bench.err | tee $OUT/${TAG}_bench_ref.json | cut -c1-200
echo "== kernel table"; timeout 900 python tools/kernel_table.py --out $OUT/${TAG}_kernel_table.json > $OUT/${TAG}_kernel_table.log 2>&1; tail -1 $OUT/${TAG}_kernel_table.log | cut -c1-100
echo "== host overhead"; timeout 300 python tools/host_overhead.py U1_D64_P1 U1xU1_D4096_P1 U1_D16384_P1 > $OUT/${TAG}_host_overhead.jsonl 2>&1; tail -2 $OUT/${TAG}_host_overhead.jsonl | cut -c1-300
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'copy_kernel|tiled_kernel|gemm_kernel|match' -c 400 --csv \
   --log-file $OUT/${TAG}_launches.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/${TAG}_launches_bench.log 2>&1
echo "== ncu full: gemm (D=16384 P1 and P2)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 6 -c 2 -f -o $OUT/${TAG}_gemm \
   python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --sizes 16384 > $OUT/${TAG}_ncu_gemm.log 2>&1
echo "== ncu full: copy (D=16384 P1 merges: long rows, short rows; T1: tiled)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'copy_kernel|tiled_kernel' -c 6 -f -o $OUT/${TAG}_copy_p1 \
   python tools/kernel_table.py --names U1_D16384_P1 --dtypes f64 --reps 1 --out $OUT/${TAG}_tmp.json > $OUT/${TAG}_ncu_copy_p1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'tiled_kernel' -c 3 -f -o $OUT/${TAG}_copy_t1 \
   python tools/kernel_table.py --names U1_D16384_T1 --dtypes f64 --reps 1 --out $OUT/${TAG}_tmp.json > $OUT/${TAG}_ncu_copy_t1.log 2>&1
rm -f $OUT/${TAG}_tmp.json
E2E=$OUT/${TAG}_e2e.jsonl; rm -f $E2E
echo "== CTMRG D=5 chi=256 (config 4)"
timeout 600 python tools/ctmrg_bench.py --D 5 --chi 256 --sweeps 5 --backend b200 --out $E2E 2>&1 | tail -1 | cut -c1-300
timeout 600 python tools/ctmrg_bench.py --D 5 --chi 256 --sweeps 5 --backend b200 --profile --out $E2E 2>&1 | tail -1 | cut -c1-300
timeout 600 python tools/ctmrg_bench.py --D 5 --chi 256 --sweeps 5 --backend b200 --decomp-workers 1 --out $E2E 2>&1 | tail -1 | cut -c1-300
timeout 600 python tools/ctmrg_bench.py --D 5 --chi 256 --sweeps 5 --backend torch --out $E2E 2>&1 | tail -1 | cut -c1-300
echo "== DMRG Hubbard U1xU1 D=4096 complex128, N=20 (config 3 shape; bonds 6..14 carry the full D): b200 vs stock torch"
timeout 900 python tools/dmrg_bench.py --model hubbard --N 20 --D 4096 --D0 4096 --sweeps 1 --backend b200 --dtype complex128 --out $E2E 2>&1 | tail -1 | cut -c1-300
timeout 900 python tools/dmrg_bench.py --model hubbard --N 20 --D 4096 --D0 4096 --sweeps 1 --backend torch --dtype complex128 --out $E2E 2>&1 | tail -1 | cut -c1-300
echo "== DMRG configs 1 and 2 (b200 / torch / numpy)"
timeout 300 python tools/dmrg_bench.py --model heisenberg --N 32 --D 64 --sweeps 3 --backend b200 --out $E2E 2>&1 | tail -1 | cut -c1-200
timeout 300 python tools/dmrg_bench.py --model heisenberg --N 32 --D 64 --sweeps 3 --backend torch --out $E2E 2>&1 | tail -1 | cut -c1-200
timeout 300 python tools/dmrg_bench.py --model heisenberg --N 32 --D 64 --sweeps 3 --backend np --out $E2E 2>&1 | tail -1 | cut -c1-200
timeout 600 python tools/dmrg_bench.py --model fermions --N 64 --D 512 --sweeps 2 --backend b200 --out $E2E 2>&1 | tail -1 | cut -c1-200
timeout 600 python tools/dmrg_bench.py --model fermions --N 64 --D 512 --sweeps 2 --backend torch --out $E2E 2>&1 | tail -1 | cut -c1-200
echo "== DMRG Hubbard U1xU1 N=64 D=4096 complex128, FULL sweep (config 3)"
timeout 1500 python tools/dmrg_bench.py --model hubbard --N 64 --D 4096 --D0 4096 --sweeps 1 --backend b200 --dtype complex128 --out $E2E 2>&1 | tail -1 | cut -c1-400
ls -la $OUT | tail -30
