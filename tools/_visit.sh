timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 300 python tools/dmrg_bench.py --model heisenberg --N 32 --D 64 --sweeps 2 --backend b200 2>&1 | tail -1 | cut -c1-600
