#!/bin/bash
# Build named variants of libcollision_b200.so with -D overrides, for A/B runs on the GPU box:
#   tools/ab_variants.sh build  name1 "-DCB2_T0_F=32 ..."  name2 "..."
#   tools/ab_variants.sh run    name1 name2 ...        (on the GPU box; prints pairs/s per variant)
set -e
cd "$(dirname "$0")/.."
mode=$1; shift
if [ "$mode" = build ]; then
  mkdir -p collision_rs_b200/variants
  while [ $# -gt 0 ]; do
    name=$1; extra=$2; shift 2
    make -s -C collision_rs_b200/csrc BUILD=build/$name OUT=../variants/libcb2_$name.so EXTRA="$extra" ../variants/libcb2_$name.so
    grep -A3 "k_epa_lane" collision_rs_b200/csrc/build/$name/cb2_epa.ptxas.log | grep registers | sed "s/^/$name lane tiers: /"
  done
else
  for name in "$@"; do
    CB2_LIB_PATH=$PWD/collision_rs_b200/variants/libcb2_$name.so python bench.py --steps 5 --warmup 3 --no-extra 2>/dev/null | \
      python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$name', round(d['value']/1e6,1), 'Mpairs/s', {k: round(v,2) for k,v in d['roofline']['stage_ms'].items()})"
  done
fi
