#!/bin/bash
# smoke + parity tests on the default library, then bench the named variants
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke > gpurun_out/c_smoke.log 2>&1; rc=$?
tail -2 gpurun_out/c_smoke.log
if [ $rc -ne 0 ]; then echo "SMOKE FAILED rc=$rc"; exit 1; fi
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c_tests.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/c_tests.log
bash tools/r2_gpu_b.sh "$@"
