#!/bin/bash
set -u
TAG=${1:-r02s}
NG=${2:-2}
OUT=gpurun_out
mkdir -p $OUT
H="--model hubbard --N 20 --D 4096 --D0 4096 --sweeps 1 --dtype complex128 --backend b200 --fused"
echo "== multi-GPU tests"; timeout 900 python -m pytest tests/test_multigpu_gpu.py -m gpu -q -x 2>&1 | tail -5 | tee $OUT/${TAG}_pytest.txt
echo "== spmd hubbard x$NG (peer arena)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29611 tools/dmrg_bench.py $H --spmd --out $OUT/${TAG}_e2e.jsonl 2>&1 | grep '"model"\|Error\|error' | cut -c1-200
echo "== spmd hubbard x$NG (peer arena) profile"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29612 tools/dmrg_bench.py $H --spmd --profile --out $OUT/${TAG}_e2e.jsonl 2>&1 | grep '"model"\|Error\|error' | cut -c1-200
echo "== spmd hubbard x$NG (nccl broadcast)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29613 tools/dmrg_bench.py $H --spmd --spmd-nccl --out $OUT/${TAG}_e2e.jsonl 2>&1 | grep '"model"\|Error\|error' | cut -c1-200
