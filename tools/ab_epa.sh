#!/bin/bash
# Fast A/B builds that differ only in cb2_epa.cu: compile that one file with -D overrides and link it
# with the other objects of the base build.   tools/ab_epa.sh name1 "-D..." name2 "-D..." ...
set -e
cd "$(dirname "$0")/../collision_rs_b200/csrc"
make -s -j8
mkdir -p ../variants build/ab
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
ARCH="-gencode arch=compute_100a,code=sm_100a"
FLAGS="$ARCH -O3 -std=c++17 -lineinfo -fmad=false -prec-div=true -prec-sqrt=true -ftz=false -Xcompiler -fPIC -Xptxas -v"
others=$(ls build/*.o | grep -v cb2_epa.o)
while [ $# -gt 0 ]; do
  name=$1; extra=$2; shift 2
  ( $NVCC $FLAGS $extra -c cb2_epa.cu -o build/ab/epa_$name.o 2> build/ab/epa_$name.log || { cat build/ab/epa_$name.log; exit 1; }
    $NVCC $ARCH -shared -o ../variants/libcb2_$name.so build/ab/epa_$name.o $others -Xcompiler -fPIC
    grep -A3 "k_epa_lane" build/ab/epa_$name.log | grep -E "registers|spill" | tr '\n' ' ' | sed "s/^/$name: /"; echo ) &
done
wait
