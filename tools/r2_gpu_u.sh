#!/bin/bash
# parity tests + bench under a variant library
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export CB2_LIB_PATH=$PWD/collision_rs_b200/variants/libcb2_$1.so
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py -m gpu -x -q > gpurun_out/u_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/u_tests.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-extra 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['value']/1e6,1), 'Mpairs/s', {k: round(v,2) for k,v in d['roofline']['stage_ms'].items()}, 'non_ok', d['run']['non_ok_status'])"
unset CB2_LIB_PATH
timeout 300 python bench.py --steps 10 --warmup 3 --no-extra 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('default', round(d['value']/1e6,1), 'Mpairs/s', {k: round(v,2) for k,v in d['roofline']['stage_ms'].items()}, 'non_ok', d['run']['non_ok_status'])"
