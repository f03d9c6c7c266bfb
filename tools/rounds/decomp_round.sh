#!/bin/bash
# GPU-box visit: sector-parallel decompositions (yastn_b200/decomp.py) — parity tests, then the D=4096 DMRG and chi=256 CTMRG
# targets with 1 (reference schedule) / 4 / 8 / 16 sector streams.   Usage: bash tools/decomp_round.sh <tag>
set -u
TAG=${1:-r01d}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest decomp"; timeout 600 python -m pytest tests/test_decomp.py -m gpu -x -q 2>&1 | tail -6 | tee $OUT/${TAG}_pytest.txt
E2E=$OUT/${TAG}_e2e.jsonl; rm -f $E2E
for W in 1 4 8 16; do
  echo "== DMRG Hubbard D=4096 c128 N=20, decomp workers $W"
  timeout 600 python tools/dmrg_bench.py --model hubbard --N 20 --D 4096 --D0 4096 --sweeps 1 --backend b200 --dtype complex128 --profile --decomp-workers $W --out $E2E 2>&1 | tail -1 | cut -c1-900
done
echo "== DMRG gesvdj, 8 workers"
YASTN_B200_SVD_DRIVER=gesvdj timeout 600 python tools/dmrg_bench.py --model hubbard --N 20 --D 4096 --D0 4096 --sweeps 1 --backend b200 --dtype complex128 --profile --decomp-workers 8 --out $E2E 2>&1 | tail -1 | cut -c1-900
for W in 1 8; do
  echo "== CTMRG D=5 chi=256, decomp workers $W"
  timeout 600 python tools/ctmrg_bench.py --D 5 --chi 256 --sweeps 5 --backend b200 --profile --decomp-workers $W --out $E2E 2>&1 | tail -1 | cut -c1-900
done
echo "== DMRG no profile, 8 workers"
timeout 600 python tools/dmrg_bench.py --model hubbard --N 20 --D 4096 --D0 4096 --sweeps 1 --backend b200 --dtype complex128 --decomp-workers 8 --out $E2E 2>&1 | tail -1 | cut -c1-400
