for sh in 0 3 5; do
python bench.py --steps 3 --warmup 3 --no-extra --shard $sh 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('shard $sh', round(d['value']/1e6,1), 'M pairs/s', round(d['ms_per_step'],2), 'ms hit', round(d['config']['hit_fraction'],4), {k: round(v,2) for k,v in d['roofline']['stage_ms'].items()})"
done
