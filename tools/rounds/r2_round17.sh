#!/bin/bash
set -u
TAG=${1:-r02z}
OUT=gpurun_out
for w in 4 8; do
  echo "== ctmrg workers $w"; timeout 600 python tools/ctmrg_bench.py --D 5 --chi 256 --sweeps 6 --backend b200 --fused --decomp-workers $w --out $OUT/${TAG}_e2e.jsonl | cut -c1-200
  echo "== ctmrg workers $w chains"; timeout 600 python tools/ctmrg_bench.py --D 5 --chi 256 --sweeps 6 --backend b200 --fused --chains --decomp-workers $w --out $OUT/${TAG}_e2e.jsonl | cut -c1-200
done
timeout 600 python tools/ctmrg_bench.py --D 5 --chi 256 --sweeps 6 --backend b200 --fused --chains --profile --out $OUT/${TAG}_e2e.jsonl | cut -c1-200
H="--model hubbard --N 20 --D 4096 --D0 4096 --sweeps 1 --dtype complex128 --backend b200 --fused --chains"
for w in 8 4; do
  echo "== hubbard workers $w"; timeout 600 python tools/dmrg_bench.py $H --decomp-workers $w --out $OUT/${TAG}_e2e.jsonl | cut -c1-160
done
echo "== hubbard f64 workers"
for w in 4 8; do
  timeout 600 python tools/dmrg_bench.py --model hubbard --N 20 --D 4096 --D0 4096 --sweeps 1 --dtype float64 --backend b200 --fused --chains --decomp-workers $w --out $OUT/${TAG}_e2e.jsonl | cut -c1-160
done
