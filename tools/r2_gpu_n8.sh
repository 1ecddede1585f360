#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
echo "GPUs: $N"; lscpu | grep -E "^CPU\(s\)|NUMA node" | head -6
nvidia-smi topo -m 2>/dev/null | head -14
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/n8_bench.json 2> gpurun_out/n8_bench.err; echo "bench N=$N rc=$?"; tail -3 gpurun_out/n8_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/n8_bench.json').read())
print('N', d['n_gpus'], 'value', round(d['value']/1e6,1), 'e2e', round(d['e2e']['value']/1e6,1), d['run'])
print("bodies e2e", d.get("e2e_bodies")); print(json.dumps(d["extra"].get("multi"), indent=1))
PY
