#!/bin/bash
# chains + C meta pass + single-copy plan upload: tests, then the launch-bound configs with and without chains
set -u
TAG=${1:-r02l}
OUT=gpurun_out
mkdir -p $OUT
echo "== smoke"; timeout 120 python __graft_entry__.py --smoke 2>&1 | tail -3 | tee $OUT/${TAG}_smoke.txt
echo "== pytest"; timeout 1200 python -m pytest tests -m gpu -q -x --deselect tests/test_reference_suite_gpu.py 2>&1 | tail -8 | tee $OUT/${TAG}_pytest.txt
for cfg in "heisenberg 32 64" "fermions 64 512"; do
  set -- $cfg
  timeout 300 python tools/dmrg_bench.py --model $1 --N $2 --D $3 --sweeps 3 --backend b200 --fused --out $OUT/${TAG}_e2e.jsonl | cut -c1-200
  timeout 300 python tools/dmrg_bench.py --model $1 --N $2 --D $3 --sweeps 3 --backend b200 --fused --chains --out $OUT/${TAG}_e2e.jsonl | cut -c1-200
  YASTN_B200_CHAIN_GRAPH=1 timeout 300 python tools/dmrg_bench.py --model $1 --N $2 --D $3 --sweeps 3 --backend b200 --fused --chains --out $OUT/${TAG}_e2e.jsonl | cut -c1-200
done
timeout 300 python -m cProfile -o /tmp/p.prof tools/dmrg_bench.py --model heisenberg --N 32 --D 64 --sweeps 3 --backend b200 --fused --chains > /dev/null 2>&1
python - <<PY > $OUT/${TAG}_prof_heisenberg.txt
import pstats
p = pstats.Stats('/tmp/p.prof')
p.sort_stats('tottime').print_stats(40)
p.sort_stats('cumulative').print_stats(60)
PY
timeout 400 python tools/dmrg_bench.py --model hubbard --N 20 --D 4096 --D0 4096 --sweeps 1 --backend b200 --fused --chains --dtype complex128 --out $OUT/${TAG}_e2e.jsonl 2>&1 | tail -1 | cut -c1-300
