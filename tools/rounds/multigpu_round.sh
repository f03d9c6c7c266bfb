#!/bin/bash
# Multi-GPU box visit (gpurun --gpus N): the sharded two-contraction chain with its three block-exchange variants, then bench.py.
# Usage: bash tools/multigpu_round.sh <tag> <ngpus>
set -u
TAG=${1:-r01m}
N=${2:-2}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi topo -m > $OUT/${TAG}_topo.txt 2>&1
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
J=$OUT/${TAG}_chain.jsonl; rm -f $J
echo "== chain D=1024 f64"; timeout 300 $RUN --master-port 29511 tools/multigpu_chain.py --case U1_D1024_chain --dtype f64 --out $J 2>&1 | grep -v "^W\|^\[W\|OMP_NUM" | tail -8 | cut -c1-600
echo "== chain D=4096 f64"; timeout 300 $RUN --master-port 29512 tools/multigpu_chain.py --case U1_D4096_chain --dtype f64 --out $J 2>&1 | grep -v "^W\|^\[W\|OMP_NUM" | tail -8 | cut -c1-600
echo "== chain D=4096 c128"; timeout 300 $RUN --master-port 29513 tools/multigpu_chain.py --case U1_D4096_chain --dtype c128 --out $J 2>&1 | grep -v "^W\|^\[W\|OMP_NUM" | tail -8 | cut -c1-600
echo "== bench N=$N"; timeout 600 $RUN --master-port 29514 bench.py --gpus $N --steps 5 --warmup 3 2>$OUT/${TAG}_bench.err | tee $OUT/${TAG}_bench_n$N.json | cut -c1-700
tail -5 $OUT/${TAG}_bench.err
ls -la $OUT
