#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
echo "== split tail"; timeout 300 python tools/e2e_probe.py 262144 524288 786432 1048576 2>&1 | grep chunk
echo "== one stream per chunk (round 1)"; CB2_SPLIT_TAIL=0 timeout 300 python tools/e2e_probe.py 524288 2>&1 | grep chunk
timeout 300 python bench.py --steps 5 --warmup 3 --no-extra --cpu-sample 100000 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('dev', round(d['value']/1e6,1), 'e2e', round(d['e2e']['value']/1e6,1), {k: round(v,2) for k,v in d['roofline']['stage_ms'].items()})"
CB2_TRACE=1 timeout 200 python tools/e2e_trace.py 2>&1 | tail -26
