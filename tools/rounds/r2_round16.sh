#!/bin/bash
set -u
TAG=${1:-r02y}
OUT=gpurun_out
F="tensor/test_tensordot.py tensor/test_ncon_einsum.py tensor/test_fuse_hard.py tensor/test_vdot.py"
echo "== default"; timeout 900 python tools/run_reference_tests.py --policies fuse_contracted --files $F > $OUT/${TAG}_fc_default.log 2>&1; grep -E "^E  |^FAILED|failed_ids" $OUT/${TAG}_fc_default.log | cut -c1-400 | head -30
echo "== noskip"; YB_GEMM_NOSKIP=1 timeout 900 python tools/run_reference_tests.py --policies fuse_contracted --files $F > $OUT/${TAG}_fc_noskip.log 2>&1; grep -E "^FAILED|\"failed\"" $OUT/${TAG}_fc_noskip.log | cut -c1-200 | head
echo "== classic"; YB_GEMM_CLASSIC=1 timeout 900 python tools/run_reference_tests.py --policies fuse_contracted --files $F > $OUT/${TAG}_fc_classic.log 2>&1; grep -E "^FAILED|\"failed\"" $OUT/${TAG}_fc_classic.log | cut -c1-200 | head
echo "== no_fusion"; timeout 900 python tools/run_reference_tests.py --policies no_fusion --files $F > $OUT/${TAG}_nf_default.log 2>&1; grep -E "^E  |^FAILED|\"failed\"" $OUT/${TAG}_nf_default.log | cut -c1-300 | head
