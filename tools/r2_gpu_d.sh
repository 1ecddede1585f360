#!/bin/bash
# metrics pass with section markers, then the N=1 bench with extras, then N=2 under torchrun
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
bash tools/r2_gpu_metrics.sh
python profiles/summarize_r2.py gpurun_out/r2_metrics.csv > gpurun_out/d_kernels.txt 2>&1; tail -3 gpurun_out/d_kernels.txt
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/d_bench1.json 2> gpurun_out/d_bench1.err; echo "bench N=1 rc=$?"; tail -3 gpurun_out/d_bench1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/d_bench1.json').read())
print('N=1', round(d['value']/1e6,1), 'e2e', round(d['e2e']['value']/1e6,1), d['run'], {k: round(v,2) for k,v in d['roofline']['stage_ms'].items()})
for k in d['roofline']['kernels']:
    ex=k.get('executed') or {}
    print(' ', k['kernel'], 'ms', round(k['ms'],3), 'alg frac', k.get('frac'), 'exec frac', ex.get('frac'))
PY
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/d_bench2.json 2> gpurun_out/d_bench2.err; echo "bench N=2 rc=$?"; tail -5 gpurun_out/d_bench2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/d_bench2.json').read())
print('N=2', round(d['value']/1e6,1), 'e2e', round(d['e2e']['value']/1e6,1), d['run'])
print(json.dumps(d['extra'].get('multi'), indent=1))
PY
fi
