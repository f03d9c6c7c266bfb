E=gpurun_out/final3_e2e.jsonl; rm -f $E
timeout 400 python tools/dmrg_bench.py --model hubbard --N 20 --D 4096 --D0 4096 --sweeps 1 --backend b200 --dtype complex128 --out $E 2>&1 | tail -1 | cut -c1-600
timeout 300 python -m pytest tests -m gpu -q -k "vdot or tall or pipeline" 2>&1 | tail -2
