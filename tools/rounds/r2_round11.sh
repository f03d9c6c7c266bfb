#!/bin/bash
# N GPUs: SPMD DMRG (yastn_b200.spmd) + the multi-GPU tests
set -u
TAG=${1:-r02p}
NG=${2:-2}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi topo -m > $OUT/${TAG}_topo.txt 2>&1
echo "== multi-GPU tests"; timeout 900 python -m pytest tests/test_multigpu_gpu.py -m gpu -q -x 2>&1 | tail -5 | tee $OUT/${TAG}_pytest.txt
echo "== spmd hubbard x$NG"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29611 tools/dmrg_bench.py --model hubbard --N 20 --D 4096 --D0 4096 --sweeps 1 --dtype complex128 --backend b200 --fused --spmd --out $OUT/${TAG}_e2e.jsonl 2>&1 | grep '"model"' | cut -c1-300
echo "== spmd hubbard x$NG profile"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29612 tools/dmrg_bench.py --model hubbard --N 20 --D 4096 --D0 4096 --sweeps 1 --dtype complex128 --backend b200 --fused --spmd --profile --out $OUT/${TAG}_e2e.jsonl 2>&1 | grep '"model"' | cut -c1-300
echo "== single hubbard (same box)"
timeout 600 python tools/dmrg_bench.py --model hubbard --N 20 --D 4096 --D0 4096 --sweeps 1 --dtype complex128 --backend b200 --fused --out $OUT/${TAG}_e2e.jsonl | cut -c1-200
