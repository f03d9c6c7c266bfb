#!/bin/bash
set -u
TAG=${1:-r02n}
OUT=gpurun_out
mkdir -p $OUT
echo "== plan trace heisenberg"
timeout 300 python tools/plan_trace.py --model heisenberg --N 32 --D 64 --sweeps 3 --backend b200 --fused --chains 2>&1 | grep '"plan"' | tee $OUT/${TAG}_plan_trace.jsonl | cut -c1-600
echo "== fermions profile"
timeout 300 python -m cProfile -o /tmp/p.prof tools/dmrg_bench.py --model fermions --N 64 --D 512 --sweeps 3 --backend b200 --fused --chains > /dev/null 2>&1
python - <<PY > $OUT/${TAG}_prof_fermions.txt
import pstats
p = pstats.Stats('/tmp/p.prof')
p.sort_stats('tottime').print_stats(40)
p.sort_stats('cumulative').print_stats(90)
PY
echo "== hubbard profile"
timeout 600 python -m cProfile -o /tmp/p.prof tools/dmrg_bench.py --model hubbard --N 20 --D 4096 --D0 4096 --sweeps 1 --dtype complex128 --backend b200 --fused --chains > /dev/null 2>&1
python - <<PY > $OUT/${TAG}_prof_hubbard.txt
import pstats
p = pstats.Stats('/tmp/p.prof')
p.sort_stats('tottime').print_stats(40)
p.sort_stats('cumulative').print_stats(90)
PY
timeout 600 python tools/dmrg_bench.py --model hubbard --N 20 --D 4096 --D0 4096 --sweeps 1 --dtype complex128 --backend b200 --fused --chains --profile --out $OUT/${TAG}_e2e.jsonl | cut -c1-300
timeout 600 python tools/dmrg_bench.py --model hubbard --N 20 --D 4096 --D0 4096 --sweeps 1 --dtype complex128 --backend b200 --fused --chains --gemm-roofline --out $OUT/${TAG}_e2e.jsonl | cut -c1-300
