#!/bin/bash
# Round-2 GPU call A: smoke -> parity tests -> A/B of the lane-tier variants -> ncu of the lane kernel.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L
timeout 300 python __graft_entry__.py smoke > gpurun_out/a_smoke.log 2>&1; rc=$?
tail -3 gpurun_out/a_smoke.log
if [ $rc -ne 0 ]; then echo "SMOKE FAILED rc=$rc"; CB2_LIB_PATH=$PWD/collision_rs_b200/variants/libcb2_r1.so timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2; exit 1; fi
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/a_tests.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/a_tests.log
for name in r1 l40 l40nopf l36 l32 l28 l40t1 l32t1; do
  CB2_LIB_PATH=$PWD/collision_rs_b200/variants/libcb2_$name.so timeout 300 python bench.py --steps 5 --warmup 3 --no-extra 2>gpurun_out/a_bench_$name.err | tee gpurun_out/a_bench_$name.json | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$name', round(d['value']/1e6,1), 'Mpairs/s', 'e2e', round(d['e2e']['value']/1e6,1), {k: round(v,2) for k,v in d['roofline']['stage_ms'].items()})" || tail -3 gpurun_out/a_bench_$name.err
done
B="python bench.py --steps 2 --warmup 3 --pairs 1000000 --no-extra"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2a.csv $B > gpurun_out/a_ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_epa_lane' --launch-skip 2 --launch-count 1 \
    -f -o gpurun_out/prof_r2a $B > gpurun_out/a_ncu2.log 2>&1
ls -la gpurun_out/prof_r2a.ncu-rep
