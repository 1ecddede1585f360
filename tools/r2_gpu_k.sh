#!/bin/bash
# A/B of the tier-0 instruction diet: smoke + GPU tests on the default build, then the bench per variant
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/k_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/k_tests.log
run() {
  CB2_LIB_PATH=$1 timeout 300 python bench.py --steps 5 --warmup 3 --no-extra 2>/dev/null | \
      python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$2', round(d['value']/1e6,1), 'Mpairs/s', {k: round(v,2) for k,v in d['roofline']['stage_ms'].items()})"
}
run $PWD/collision_rs_b200/libcollision_b200.so default
for name in "$@"; do run $PWD/collision_rs_b200/variants/libcb2_$name.so $name; done
