#!/bin/bash
# broad phase: parity tests, then timing (1 M / 8 M boxes, ordered and as a set, one of eight shards)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -k "sap or brute or grid or shard or sweep or full_size" > gpurun_out/q_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/q_tests.log
timeout 300 python tools/sap_probe.py 1000000 2>&1 | tail -4
timeout 300 python tools/sap_probe.py 8000000 2>&1 | tail -4
timeout 300 python tools/sap_shard_probe.py 2>&1 | tail -4
