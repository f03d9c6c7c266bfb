#!/bin/bash
# GPU-box visit for the north-star end-to-end targets (BASELINE.json configs 3 and 4) next to parity + bench.
# Usage: bash tools/target_round.sh <tag>
set -u
TAG=${1:-r01t}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
nproc > $OUT/${TAG}_nproc.txt
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $OUT/${TAG}_pytest.txt
echo "== smoke"; timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3 | tee $OUT/${TAG}_smoke.txt
echo "== bench"; timeout 900 python bench.py 2>$OUT/${TAG}_bench.err | tee $OUT/${TAG}_bench.json | cut -c1-600
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>>$OUT/${TAG}_bench.err | tee $OUT/${TAG}_bench_ref.json | cut -c1-300
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'copy_kernel|gemm_kernel|match' -c 400 --csv \
   --log-file $OUT/${TAG}_launches.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/${TAG}_launches_bench.log 2>&1
E2E=$OUT/${TAG}_e2e.jsonl; rm -f $E2E
echo "== CTMRG D=5 chi=256 (config 4)"
timeout 600 python tools/ctmrg_bench.py --D 5 --chi 256 --sweeps 5 --backend b200 --out $E2E 2>&1 | tail -1 | cut -c1-500
timeout 600 python tools/ctmrg_bench.py --D 5 --chi 256 --sweeps 5 --backend b200 --profile --out $E2E 2>&1 | tail -1 | cut -c1-1500
timeout 600 python tools/ctmrg_bench.py --D 5 --chi 256 --sweeps 5 --backend torch --out $E2E 2>&1 | tail -1 | cut -c1-500
echo "== DMRG Hubbard U1xU1 D=4096 complex128 (config 3; N=${DMRG_N:-20} sites: bonds 6..N-6 carry the full D)"
timeout ${DMRG_TMO:-700} python tools/dmrg_bench.py --model hubbard --N ${DMRG_N:-20} --D 4096 --D0 4096 --sweeps 1 --backend b200 --dtype complex128 --profile --out $E2E 2>&1 | tail -1 | cut -c1-1800
ls -la $OUT
