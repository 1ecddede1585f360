#!/bin/bash
cd "$(dirname "$0")/.."
for name in "$@"; do
  if [ "$name" = default ]; then unset CB2_LIB_PATH; else export CB2_LIB_PATH=$PWD/collision_rs_b200/variants/libcb2_$name.so; fi
  timeout 300 python tools/curved_probe.py 1000000 2>&1 | tail -1
done
