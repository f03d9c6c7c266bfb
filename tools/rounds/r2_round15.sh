#!/bin/bash
set -u
TAG=${1:-r02x}
OUT=gpurun_out
mkdir -p $OUT
H="--model hubbard --N 20 --D 4096 --D0 4096 --sweeps 1 --dtype complex128 --backend b200 --fused --chains"
for w in 2 3 4 6; do
  echo "== workers $w"; timeout 600 python tools/dmrg_bench.py $H --decomp-workers $w --out $OUT/${TAG}_e2e.jsonl | cut -c1-160
done
for w in 2 4 8; do
  echo "== fermions workers $w"; timeout 600 python tools/dmrg_bench.py --model fermions --N 64 --D 512 --sweeps 3 --backend b200 --fused --chains --decomp-workers $w --out $OUT/${TAG}_e2e.jsonl | cut -c1-160
done
echo "== reference suite"; timeout 2300 python -m pytest tests/test_reference_suite_gpu.py -m gpu -q -x 2>&1 | tail -6 | tee $OUT/${TAG}_pytest_ref.txt
