#!/bin/bash
# GJK-ahead host pipeline: parity of the host-buffer path, then e2e A/B
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "host or pipeline or e2e or chunk or parity" > gpurun_out/l_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/l_tests.log
for cfg in "1 1 1 3" "1 0 1 3" "1 1 0 3" "1 1 1 4" "1 1 1 5" "0 0 1 3"; do set -- $cfg
echo "AHEAD=$1 TAIL_PRIO=$2 RAMP=$3 NC=$4"; CB2_AHEAD=$1 CB2_TAIL_PRIO=$2 CB2_RAMP=$3 CB2_NC=$4 timeout 300 python tools/e2e_probe.py 393216 524288 786432 2>&1 | grep chunk
done
CB2_TRACE=1 timeout 300 python tools/e2e_probe.py 524288 2> gpurun_out/l_trace.txt | grep chunk
