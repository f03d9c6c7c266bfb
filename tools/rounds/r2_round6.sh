#!/bin/bash
# host-time profile of the launch-bound configs (1 and 2): where does a D=64 sweep spend its wall clock?
set -u
TAG=${1:-r02k}
OUT=gpurun_out
mkdir -p $OUT
for cfg in "heisenberg 32 64" "fermions 64 512"; do
  set -- $cfg
  timeout 300 python -m cProfile -o /tmp/p_$1.prof tools/dmrg_bench.py --model $1 --N $2 --D $3 --sweeps 3 --backend b200 --fused > $OUT/${TAG}_prof_$1.json 2>&1
  python - <<PY > $OUT/${TAG}_prof_$1.txt
import pstats
p = pstats.Stats('/tmp/p_$1.prof')
p.sort_stats('tottime').print_stats(45)
p.sort_stats('cumulative').print_stats(70)
PY
  timeout 300 python tools/dmrg_bench.py --model $1 --N $2 --D $3 --sweeps 3 --backend np --out $OUT/${TAG}_e2e.jsonl | cut -c1-200
  timeout 300 python tools/dmrg_bench.py --model $1 --N $2 --D $3 --sweeps 3 --backend b200 --fused --out $OUT/${TAG}_e2e.jsonl | cut -c1-200
done
