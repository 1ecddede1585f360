#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py -m gpu -q -k "sap or brute or broad or chain or box or multi" > gpurun_out/j_tests.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/j_tests.log
timeout 300 python tools/sap_shard_probe.py 1000000
timeout 200 python tools/sap_probe.py 1000000 2>&1 | tail -2
timeout 200 python tools/sap_probe.py 8000000 2>&1 | tail -2
