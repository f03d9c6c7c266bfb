#!/bin/bash
set -u
TAG=${1:-r02w}
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python tools/svd_rect_probe.py | tee $OUT/${TAG}_svd_rect.jsonl
H="--model hubbard --N 20 --D 4096 --D0 4096 --sweeps 1 --dtype complex128 --backend b200 --fused --chains"
for w in 4 12 16; do
  echo "== workers $w"; YASTN_B200_DECOMP_WORKERS=$w timeout 600 python tools/dmrg_bench.py $H --decomp-workers $w --out $OUT/${TAG}_e2e.jsonl | cut -c1-160
done
