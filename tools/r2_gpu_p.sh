#!/bin/bash
# ncu --set full + source of a kernel ($1 = regex, $2 = launch-skip, $3 = cfg) at 1 M pairs -> gpurun_out/prof_p.ncu-rep
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CB2_PROFILE_PAIRS=1000000 timeout 900 ncu --set full --clock-control none --import-source on -k regex:$1 --launch-skip $2 --launch-count 1 -f -o gpurun_out/prof_p python tools/profile_pass.py $3 > gpurun_out/p_ncu.log 2>&1
ls -la gpurun_out/prof_p.ncu-rep
