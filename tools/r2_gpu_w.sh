#!/bin/bash
# support-cell resolution sweep (CB2_CELL_N="N(V<=64),N(V<=512),N(beyond)")
cd "$(dirname "$0")/.."
for cn in "4,8,16" "6,8,16" "4,12,16" "6,12,16" "8,16,16" "8,12,16"; do
  CB2_CELL_N=$cn timeout 300 python bench.py --steps 10 --warmup 3 --no-extra 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$cn', round(d['value']/1e6,1), 'Mpairs/s', {k: round(v,2) for k,v in d['roofline']['stage_ms'].items()}, 'non_ok', d['run']['non_ok_status'])"
done
