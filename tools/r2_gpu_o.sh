#!/bin/bash
# ncu --set full + source of tier 0 (default build, 1 M pairs) -> gpurun_out/prof_o.ncu-rep
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CB2_PROFILE_PAIRS=1000000 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_epa --launch-skip 3 --launch-count 1 -f -o gpurun_out/prof_o python tools/profile_pass.py cfg3 > gpurun_out/o_ncu.log 2>&1
ls -la gpurun_out/prof_o.ncu-rep
