#!/bin/bash
# Round-2 end evidence on one GPU box: bench lines (all legs), per-kernel table, ncu launch list, end-to-end targets
# (BASELINE.json configs 1-4) with our backend, the stock torch backend on the same GPU and numpy on the host.
# Usage: bash tools/final_round_r02.sh <tag>
set -u
TAG=${1:-r02f}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
nproc > $OUT/${TAG}_nproc.txt
echo "== smoke"; timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3 | tee $OUT/${TAG}_smoke.txt
echo "== bench f64 (all legs)"; timeout 1500 python bench.py 2>$OUT/${TAG}_bench.err | tee $OUT/${TAG}_bench_f64.json | cut -c1-300
echo "== bench c128"; timeout 900 python bench.py --dtype c128 --steps 5 --no-cpu-baseline --no-gpu-baseline --no-dmrg 2>>$OUT/${TAG}_bench.err | tee $OUT/${TAG}_bench_c128.json | cut -c1-300
echo "== bench reference arm"; timeout 900 python bench.py --impl reference --steps 2 --warmup 1 2>>$OUT/${TAG}_bench.err | tee $OUT/${TAG}_bench_ref.json | cut -c1-200
echo "== kernel table"; timeout 900 python tools/kernel_table.py --out $OUT/${TAG}_kernel_table.json > $OUT/${TAG}_kernel_table.log 2>&1; tail -1 $OUT/${TAG}_kernel_table.log | cut -c1-100
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'copy_kernel|tiled_kernel|gemm_kernel|gemm_ws_kernel|skinny_kernel|panel_kernel|ewise_kernel|svd_jacobi_kernel' -c 500 --csv \
   --log-file $OUT/${TAG}_launches.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-gpu-baseline --no-dmrg > $OUT/${TAG}_launches_bench.log 2>&1
E2E=$OUT/${TAG}_e2e.jsonl; rm -f $E2E
echo "== configs 1 and 2 (b200 plain / chains, stock torch, numpy)"
for cfg in "heisenberg 32 64" "fermions 64 512"; do
  set -- $cfg
  timeout 300 python tools/dmrg_bench.py --model $1 --N $2 --D $3 --sweeps 4 --backend b200 --fused --chains --out $E2E 2>&1 | tail -1 | cut -c1-160
  timeout 300 python tools/dmrg_bench.py --model $1 --N $2 --D $3 --sweeps 4 --backend b200 --out $E2E 2>&1 | tail -1 | cut -c1-160
  timeout 600 python tools/dmrg_bench.py --model $1 --N $2 --D $3 --sweeps 4 --backend torch --out $E2E 2>&1 | tail -1 | cut -c1-160
  timeout 600 python tools/dmrg_bench.py --model $1 --N $2 --D $3 --sweeps 4 --backend np --out $E2E 2>&1 | tail -1 | cut -c1-160
done
echo "== CTMRG D=5 chi=256 (config 4)"
timeout 600 python tools/ctmrg_bench.py --D 5 --chi 256 --sweeps 5 --backend b200 --out $E2E 2>&1 | tail -1 | cut -c1-200
timeout 600 python tools/ctmrg_bench.py --D 5 --chi 256 --sweeps 5 --backend torch --out $E2E 2>&1 | tail -1 | cut -c1-200
echo "== config 3 shape, N=20: policies"
H="--model hubbard --N 20 --D 4096 --D0 4096 --sweeps 1 --dtype complex128"
timeout 900 python tools/dmrg_bench.py $H --backend b200 --fused --chains --gemm-roofline --out $E2E 2>&1 | tail -1 | cut -c1-200
timeout 900 python tools/dmrg_bench.py $H --backend b200 --policy fuse_contracted --out $E2E 2>&1 | tail -1 | cut -c1-200
timeout 900 python tools/dmrg_bench.py $H --backend b200 --policy no_fusion --out $E2E 2>&1 | tail -1 | cut -c1-200
echo "== config 3: N=64 D=4096 complex128, FULL sweep"
timeout 1500 python tools/dmrg_bench.py --model hubbard --N 64 --D 4096 --D0 4096 --sweeps 1 --dtype complex128 --backend b200 --fused --chains --gemm-roofline --out $E2E 2>&1 | tail -1 | cut -c1-300
ls -la $OUT | tail -12
