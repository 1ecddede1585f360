#!/bin/bash
# Round-2 evidence run: tests, per-kernel counters at full size, launch lists, one --set full capture of
# the dominant kernel, then the bench (N=1 with extras).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/f_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/f_tests.log
bash tools/r2_gpu_metrics.sh
python profiles/summarize_r2.py gpurun_out/r2_metrics.csv > gpurun_out/f_kernels.txt 2>&1; tail -2 gpurun_out/f_kernels.txt
# launch list of one cfg-3 step at 4 M pairs, and of one SAP call
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_cfg3.csv python tools/profile_pass.py cfg3 > gpurun_out/f_ncu_l1.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_cfg4.csv python tools/profile_pass.py cfg4 > gpurun_out/f_ncu_l2.log 2>&1
# --set full of tier 0 at 4 M pairs (second pass of the profile script: launch-skip 1)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_epa --launch-skip 3 --launch-count 1 -f -o gpurun_out/prof_r2_tier0 python tools/profile_pass.py cfg3 > gpurun_out/f_ncu_full.log 2>&1
ls -la gpurun_out/prof_r2_tier0.ncu-rep
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/f_bench1.json 2> gpurun_out/f_bench1.err; echo "bench rc=$?"; tail -2 gpurun_out/f_bench1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/f_bench1.json').read())
print('N=1', round(d['value']/1e6,1), 'e2e', round(d['e2e']['value']/1e6,1), {k: round(v,2) for k,v in d['roofline']['stage_ms'].items()}, 'frac', d['roofline']['frac'], 'traffic', d['roofline']['traffic'])
for k,v in d['extra'].items():
    if isinstance(v, dict): print(' ', k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items() if a in ('ms','pairs_per_s','aabbs_per_s','hbm_frac')})
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 --cpu-sample 1000000 2>/dev/null | cut -c1-400
