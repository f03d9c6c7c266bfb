set -u
OUT=gpurun_out; TAG=r01c6; mkdir -p $OUT
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $OUT/${TAG}_pytest.txt
echo "== kernel table"; timeout 900 python tools/kernel_table.py --out $OUT/${TAG}_kernel_table.json > $OUT/${TAG}_kernel_table.log 2>&1; tail -1 $OUT/${TAG}_kernel_table.log | cut -c1-100
echo "== bench"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>$OUT/${TAG}_bench.err | tee $OUT/${TAG}_bench.json | cut -c1-200
