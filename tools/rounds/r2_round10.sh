#!/bin/bash
# 1 GPU: plan creation cost at D=4096 after the pool rework; allocator knob
set -u
TAG=${1:-r02o}
OUT=gpurun_out
mkdir -p $OUT
timeout 120 python tools/plan_create_probe.py --case U1xU1_D4096_P1 | tee -a $OUT/${TAG}_plan_create.jsonl
echo "== plan trace hubbard"
timeout 600 python tools/plan_trace.py --model hubbard --N 20 --D 4096 --D0 4096 --sweeps 1 --dtype complex128 --backend b200 --fused --chains --out $OUT/${TAG}_e2e.jsonl 2>&1 | grep '"plan"\|"model"' | tee $OUT/${TAG}_plan_trace.jsonl | cut -c1-700
echo "== hubbard expandable segments"
PYTORCH_CUDA_ALLOC_CONF=expandable_segments:True timeout 600 python tools/dmrg_bench.py --model hubbard --N 20 --D 4096 --D0 4096 --sweeps 1 --dtype complex128 --backend b200 --fused --chains --out $OUT/${TAG}_e2e.jsonl | cut -c1-200
