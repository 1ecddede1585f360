#!/bin/bash
# A/B by bench only: variants named on the command line
cd "$(dirname "$0")/.."
run() {
  CB2_LIB_PATH=$1 timeout 300 python bench.py --steps 10 --warmup 3 --no-extra 2>/dev/null | \
      python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$2', round(d['value']/1e6,1), 'Mpairs/s', {k: round(v,2) for k,v in d['roofline']['stage_ms'].items()}, 'non_ok', d['run']['non_ok_status'])"
}
run $PWD/collision_rs_b200/libcollision_b200.so default
for name in "$@"; do run $PWD/collision_rs_b200/variants/libcb2_$name.so $name; done
