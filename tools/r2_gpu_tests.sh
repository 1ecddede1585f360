#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L
timeout 300 python __graft_entry__.py smoke > gpurun_out/t_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/t_smoke.log
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/t_tests.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/t_tests.log
