#!/bin/bash
# e2e: ramp depth (front, tail) at the current kernels
cd "$(dirname "$0")/.."
for r in "3,3" "2,4" "3,4" "3,5" "4,4" "2,3" "4,3"; do
echo "RAMP=$r"; CB2_RAMP=$r timeout 300 python tools/e2e_probe.py 655360 786432 2>&1 | grep chunk
done
