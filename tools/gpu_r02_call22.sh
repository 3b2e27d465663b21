#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02_bench_default.json') if l.startswith('{')][-1])
print({k: d[k] for k in ('value','ms_per_step','e2e','gpu_launches')}); print(d['roofline']); print(d['clocks'])
lg=d['lightgcn']; print({k: lg[k] for k in ('value','ms_per_step','roofline','clocks')})
print(lg['train_step']['ms_per_step'], lg['train_step']['roofline']['frac'])
for k,v in d['extra'].items(): print(k, v.get('value'), v.get('ms_per_step'), v.get('error'))
PY
tail -2 gpurun_out/r02_bench_default.err
timeout 600 python -m pytest tests -m gpu -q --timeout 600 --tb=short -k "propagate or lightgcn or bench or smoke or c2" 2>&1 | tail -3
