#!/bin/bash
cd "$(dirname "$0")/.."
for nc in 3 2 4; do for ramp in 1; do
echo "NC=$nc RAMP=$ramp"; CB2_NC=$nc CB2_RAMP=$ramp timeout 300 python tools/e2e_probe.py 524288 786432 1048576 1572864 2097152 2>&1 | grep chunk
done; done
