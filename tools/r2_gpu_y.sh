#!/bin/bash
# thread-per-pair kernels (time of impact, distance, GJK stage): parity + the bench extras
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -k "toi or impact or distance or cfg5 or cfg2 or primitive3 or complex or hazard or settings" > gpurun_out/y_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/y_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('N=1', round(d['value']/1e6,1), 'e2e', round(d['e2e']['value']/1e6,1), {k: round(v,2) for k,v in d['roofline']['stage_ms'].items()})
for k,v in d['extra'].items():
    if isinstance(v, dict): print(' ', k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items() if a in ('ms','pairs_per_s','aabbs_per_s','distance_pairs_per_s')})
"
