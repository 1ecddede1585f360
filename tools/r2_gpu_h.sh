#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py -m gpu -q -k "sap or brute or broad or chain or full_size or box" > gpurun_out/h_tests.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/h_tests.log
echo "== segment ordering"; timeout 200 python tools/sap_probe.py 1000000 2>&1 | tail -4
echo "== radix pair sort (round 1)"; CB2_SAP_RADIX_PAIRS=1 timeout 200 python tools/sap_probe.py 1000000 2>&1 | tail -4
timeout 200 python tools/sap_probe.py 8000000 2>&1 | tail -3
