#!/bin/bash
set -u
TAG=${1:-r02m}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_reference_suite_gpu.py 2>&1 | tail -8 | tee $OUT/${TAG}_pytest.txt
echo "== plan create probe"
for c in U1_D64_P1 U1_D1024_P1 U1xU1_D4096_P1; do timeout 120 python tools/plan_create_probe.py --case $c | tee -a $OUT/${TAG}_plan_create.jsonl; done
for cfg in "heisenberg 32 64" "fermions 64 512"; do
  set -- $cfg
  timeout 300 python tools/dmrg_bench.py --model $1 --N $2 --D $3 --sweeps 4 --backend np --gc-freeze --out $OUT/${TAG}_e2e.jsonl | cut -c1-200
  timeout 300 python tools/dmrg_bench.py --model $1 --N $2 --D $3 --sweeps 4 --backend b200 --fused --out $OUT/${TAG}_e2e.jsonl | cut -c1-200
  timeout 300 python tools/dmrg_bench.py --model $1 --N $2 --D $3 --sweeps 4 --backend b200 --fused --chains --out $OUT/${TAG}_e2e.jsonl | cut -c1-200
  timeout 300 python tools/dmrg_bench.py --model $1 --N $2 --D $3 --sweeps 4 --backend b200 --fused --chains --gc-freeze --out $OUT/${TAG}_e2e.jsonl | cut -c1-200
done
timeout 300 python -m cProfile -o /tmp/p.prof tools/dmrg_bench.py --model heisenberg --N 32 --D 64 --sweeps 4 --backend b200 --fused --chains --gc-freeze > /dev/null 2>&1
python - <<PY > $OUT/${TAG}_prof_heisenberg.txt
import pstats
p = pstats.Stats('/tmp/p.prof')
p.sort_stats('tottime').print_stats(40)
p.sort_stats('cumulative').print_stats(90)
PY
