#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 200 python __graft_entry__.py smoke > gpurun_out/i_smoke.log 2>&1; rc=$?; tail -1 gpurun_out/i_smoke.log
if [ $rc -ne 0 ]; then echo "SMOKE FAILED rc=$rc"; tail -5 gpurun_out/i_smoke.log; exit 1; fi
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/i_tests.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/i_tests.log
for i in 1 2; do
timeout 300 python bench.py --steps 10 --warmup 3 --no-extra --cpu-sample 100000 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('dev', round(d['value']/1e6,1), 'e2e', round(d['e2e']['value']/1e6,1), {k: round(v,2) for k,v in d['roofline']['stage_ms'].items()})"
done
