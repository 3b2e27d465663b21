#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 --tb=short 2>&1 | grep -v "Warning\|^  warn\|return torch" | tail -12
timeout 900 python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_default.json'))
print({k: d[k] for k in ('value','ms_per_step','e2e','gpu_launches','cpu_baseline','gpu_library_baseline')})
print(d['roofline'])
print(d['d128'])
lg=d['lightgcn']; print({k: lg[k] for k in ('value','ms_per_step','roofline','cpu_baseline','gpu_library_baseline')}); print(lg['train_step'])
for k,v in d['extra'].items(): print(k, v.get('value'), v.get('ms_per_step'), v.get('e2e',{}).get('first_call_ms'), v.get('gpu_launches'), v.get('k4_towers'), v.get('gpu_library_baseline',{}).get('value'), v.get('cpu_baseline',{}).get('value'), v.get('error'))
PY
tail -3 gpurun_out/r02_bench_default.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err; cut -c1-2500 gpurun_out/r02_bench_ref.json; tail -2 gpurun_out/r02_bench_ref.err
timeout 300 python tools/gpu_tower_probe.py 2>&1 | grep "^{" | tee gpurun_out/r02_tower_probe.jsonl
timeout 300 python tools/gpu_d128_probe.py 2>&1 | grep "^{" | tee gpurun_out/r02_d128_probe.jsonl
