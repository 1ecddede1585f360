#!/bin/bash
# bench a list of variants (tier-0 stage time is the number to compare)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for name in "$@"; do
  CB2_LIB_PATH=$PWD/collision_rs_b200/variants/libcb2_$name.so timeout 300 python bench.py --steps 5 --warmup 3 --no-extra --cpu-sample 100000 2>gpurun_out/b_bench_$name.err | tee gpurun_out/b_bench_$name.json | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$name', round(d['value']/1e6,1), 'Mpairs/s', 'e2e', round(d['e2e']['value']/1e6,1), {k: round(v,2) for k,v in d['roofline']['stage_ms'].items()}, 'bad', d['config']['non_ok_status'])" || tail -3 gpurun_out/b_bench_$name.err
done
