#!/bin/bash
# e2e trace + chunk sweep at the current kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for cfg in "1 1" "1 0" "0 1"; do set -- $cfg
echo "AHEAD=$1 RAMP=$2"; CB2_AHEAD=$1 CB2_RAMP=$2 timeout 300 python tools/e2e_probe.py 393216 524288 786432 1048576 2>&1 | grep chunk
done
CB2_TRACE=1 timeout 300 python tools/e2e_probe.py 786432 2> gpurun_out/s_trace.txt | grep chunk
grep "stages" -A14 gpurun_out/s_trace.txt | tail -12
