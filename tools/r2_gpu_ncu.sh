#!/bin/bash
# ncu --set full of one kernel (regex $2) for variant $1 at 1 M pairs -> gpurun_out/prof_$3.ncu-rep
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export CB2_LIB_PATH=$PWD/collision_rs_b200/variants/libcb2_$1.so
B="python bench.py --steps 2 --warmup 3 --pairs 1000000 --no-extra --cpu-sample 50000"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$2" --launch-skip 2 --launch-count 1 \
    -f -o gpurun_out/prof_$3 $B > gpurun_out/ncu_$3.log 2>&1
ls -la gpurun_out/prof_$3.ncu-rep
