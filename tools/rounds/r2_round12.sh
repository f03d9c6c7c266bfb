#!/bin/bash
# N GPUs: SPMD DMRG sweep (plain + device-synchronised profile); N=1 reference numbers on the same box when asked
set -u
TAG=${1:-r02q}
NG=${2:-2}
SINGLE=${3:-0}
OUT=gpurun_out
mkdir -p $OUT
H="--model hubbard --N 20 --D 4096 --D0 4096 --sweeps 1 --dtype complex128 --backend b200 --fused"
echo "== spmd hubbard x$NG"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29611 tools/dmrg_bench.py $H --spmd --out $OUT/${TAG}_e2e.jsonl 2>&1 | grep '"model"' | cut -c1-200
echo "== spmd hubbard x$NG profile"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29612 tools/dmrg_bench.py $H --spmd --profile --out $OUT/${TAG}_e2e.jsonl 2>&1 | grep '"model"' | cut -c1-200
if [ "$SINGLE" = "1" ]; then
  echo "== single hubbard (same box)"
  timeout 600 python tools/dmrg_bench.py $H --out $OUT/${TAG}_e2e.jsonl | cut -c1-200
  timeout 600 python tools/dmrg_bench.py $H --profile --out $OUT/${TAG}_e2e.jsonl | cut -c1-200
fi
