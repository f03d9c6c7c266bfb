#!/bin/bash
# bench.py on N GPUs exactly as the driver launches it (+ the SPMD DMRG sweep it carries as second workload)
set -u
TAG=${1:-r02s}
NG=${2:-2}
OUT=gpurun_out
mkdir -p $OUT
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29701 bench.py --gpus $NG --steps 10 --warmup 3 2>$OUT/${TAG}_bench_n${NG}.err | tee $OUT/${TAG}_bench_n${NG}.json | cut -c1-400
tail -3 $OUT/${TAG}_bench_n${NG}.err | cut -c1-300
