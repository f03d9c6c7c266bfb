#!/bin/bash
# Round 2, GPU call 1: parity (incl. the new skinny path and the reference's own tests), bench line, per-kernel table,
# ncu captures of the launches VERDICT r01 names (U1xU1 GEMMs and merges, P3), host profile of config 1.
set -u
TAG=${1:-r02a}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
nproc > $OUT/${TAG}_nproc.txt
echo "== pytest -m gpu (without the reference suite)"; timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_reference_suite_gpu.py 2>&1 | tail -15 | tee $OUT/${TAG}_pytest.txt
echo "== kernel table"; timeout 600 python tools/kernel_table.py --reps 5 --names U1_D16384_P3 U1_D4096_P3 U1xU1_D4096_P1 U1xU1_D4096_P2 U1xU1_D4096_P3 Z2_D512_P2 U1_D16384_P1 U1_D16384_P2 U1_D1024_P1 --out $OUT/${TAG}_kernel_table.json > $OUT/${TAG}_kernel_table.log 2>&1; tail -2 $OUT/${TAG}_kernel_table.log | cut -c1-200
echo "== bench f64"; timeout 1200 python bench.py 2>$OUT/${TAG}_bench.err | tee $OUT/${TAG}_bench_f64.json | cut -c1-400
echo "== bench reference arm"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>>$OUT/${TAG}_bench.err | tee $OUT/${TAG}_bench_ref.json | cut -c1-300
echo "== ncu full: U1xU1 P1/P2 gemm + merges, P3 skinny"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'gemm_kernel|skinny_kernel' -s 4 -c 2 -f -o $OUT/${TAG}_gemm_u1u1_p1 \
   python tools/kernel_table.py --names U1xU1_D4096_P1 --dtypes f64 --reps 1 --out $OUT/${TAG}_tmp.json > $OUT/${TAG}_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'gemm_kernel|skinny_kernel' -s 4 -c 2 -f -o $OUT/${TAG}_gemm_u1u1_p2 \
   python tools/kernel_table.py --names U1xU1_D4096_P2 --dtypes f64 --reps 1 --out $OUT/${TAG}_tmp.json > $OUT/${TAG}_ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'skinny_kernel' -s 2 -c 1 -f -o $OUT/${TAG}_skinny_p3 \
   python tools/kernel_table.py --names U1_D16384_P3 --dtypes f64 --reps 1 --out $OUT/${TAG}_tmp.json > $OUT/${TAG}_ncu3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'copy_kernel|tiled_kernel' -c 4 -f -o $OUT/${TAG}_copy_u1u1 \
   python tools/kernel_table.py --names U1xU1_D4096_P1 --dtypes f64 --reps 1 --out $OUT/${TAG}_tmp.json > $OUT/${TAG}_ncu4.log 2>&1
rm -f $OUT/${TAG}_tmp.json
echo "== host profile, config 1 (Heisenberg N=32 D=64)"
timeout 300 python -c "
import cProfile, pstats, sys, io
sys.argv=['dmrg_bench.py','--model','heisenberg','--N','32','--D','64','--sweeps','3','--backend','b200']
sys.path.insert(0,'tools')
import dmrg_bench
pr=cProfile.Profile(); pr.enable(); dmrg_bench.main(); pr.disable()
s=io.StringIO(); pstats.Stats(pr,stream=s).sort_stats('cumulative').print_stats(70); open('$OUT/${TAG}_cprofile_cfg1.txt','w').write(s.getvalue())
s=io.StringIO(); pstats.Stats(pr,stream=s).sort_stats('tottime').print_stats(50); open('$OUT/${TAG}_cprofile_cfg1_tottime.txt','w').write(s.getvalue())
" 2>&1 | tail -1 | cut -c1-300
echo "== reference suite (fuse_to_matrix incl. ctmrg)"; timeout 1500 python -m pytest tests/test_reference_suite_gpu.py -q -x -k fuse_to_matrix 2>&1 | tail -5 | tee $OUT/${TAG}_pytest_ref.txt
ls -la $OUT | grep ${TAG}
