#!/bin/bash
# Install the UNMODIFIED reference into baseline/_ref (git-ignored, travels to the GPU box with gpurun) for the drop-in tests.
# The reference's own hot-path test files are placed next to it (baseline/_ref/ref_tests) so that they can be run on the box
# against our backend:  python tools/run_reference_tests.py
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
REF=${YASTN_REF:-/root/reference}
[ -d "$REF/yastn" ] || { echo "no reference checkout at $REF"; exit 0; }
rm -rf /tmp/yastn_src && cp -r "$REF" /tmp/yastn_src
python -m pip install -q --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --upgrade --target "$ROOT/baseline/_ref" /tmp/yastn_src
mkdir -p "$ROOT/baseline/_ref/ref_tests"
cp "$REF/conftest.py" "$ROOT/baseline/_ref/ref_tests/conftest.py"
cp -r "$REF/tests/tensor" "$REF/tests/mps" "$ROOT/baseline/_ref/ref_tests/"
mkdir -p "$ROOT/baseline/_ref/ref_tests/peps" && cp "$REF/tests/peps/test_ctmrg.py" "$ROOT/baseline/_ref/ref_tests/peps/" 2>/dev/null || true
rm -rf /tmp/yastn_src
echo "reference installed in $ROOT/baseline/_ref"
