#!/bin/bash
# DMRG end-to-end comparison on one GPU box: our backend vs the stock torch backend on the same GPU vs numpy on the host cores.
OUT=gpurun_out/dmrg_${1:-r01}.jsonl
rm -f $OUT
run() { timeout ${TMO:-600} python tools/dmrg_bench.py "$@" --out $OUT 2>&1 | tail -1 | cut -c1-400; }
run --model heisenberg --N 32 --D 64 --sweeps 3 --backend b200
run --model heisenberg --N 32 --D 64 --sweeps 3 --backend torch
run --model heisenberg --N 32 --D 64 --sweeps 3 --backend np
run --model fermions --N 64 --D 512 --sweeps 2 --backend b200
run --model fermions --N 64 --D 512 --sweeps 2 --backend torch
TMO=900 run --model fermions --N 64 --D 512 --sweeps 2 --backend np
run --model hubbard --N 32 --D 1024 --sweeps 2 --backend b200 --dtype complex128
run --model hubbard --N 32 --D 1024 --sweeps 2 --backend torch --dtype complex128
nproc
