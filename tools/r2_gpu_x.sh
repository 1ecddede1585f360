#!/bin/bash
cd "$(dirname "$0")/.."
for name in "$@"; do
  export CB2_LIB_PATH=$PWD/collision_rs_b200/variants/libcb2_$name.so
  echo "== $name"; timeout 300 python tools/sap_probe.py 1000000 2>&1 | tail -2
done
