#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
for v in default r1; do
  if [ $v = r1 ]; then export CB2_LIB_PATH=$PWD/collision_rs_b200/variants/libcb2_r1.so; else unset CB2_LIB_PATH; fi
  echo "== $v"; timeout 300 python tools/e2e_probe.py 524288 786432 2>&1 | grep chunk
done
unset CB2_LIB_PATH
timeout 300 python bench.py --steps 5 --warmup 3 --no-extra --cpu-sample 100000 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('dev', round(d['value']/1e6,1), 'e2e', round(d['e2e']['value']/1e6,1), {k: round(v,2) for k,v in d['roofline']['stage_ms'].items()})"
CB2_TRACE=1 timeout 200 python tools/e2e_trace.py 2>&1 | tail -26
