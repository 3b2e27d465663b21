#!/bin/bash
# Round-end check in one 1-GPU gpurun call: what the driver runs (smoke, -m gpu suite, default bench, reference arm), the
# sanitizer passes on the kernels added this round, and the ncu captures behind roofline.traffic.
#   /usr/local/graft/bin/gpurun --timeout 3000 -- 'bash tools/gpu_round_check.sh'
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 --tb=short 2>&1 | grep -v "Warning\|^  warn\|return torch" | tail -8
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_default.json'))
print({k: d[k] for k in ('value','ms_per_step','e2e','gpu_launches','cpu_baseline','gpu_library_baseline')}); print(d['roofline']); print(d['d128']['value'], d['d128']['roofline'])
for r in d['robustness']: print(r['case'][:70], r['users_per_s'], r['sweep_tflops'], r.get('sweep_frac_of_tf32_sustained', r.get('frac_of_fp32_ffma_peak')), r.get('n_refined_all_steps'))
lg=d['lightgcn']; print({k: lg[k] for k in ('value','ms_per_step','roofline')}); print(lg['train_step'])
for k,v in d['extra'].items(): print(k, v.get('value'), v.get('ms_per_step'), v.get('e2e',{}).get('first_call_ms'), v.get('gpu_launches'), v.get('k4_towers'), v.get('gpu_library_baseline',{}).get('value'), v.get('cpu_baseline',{}).get('value'), v.get('error'))
PY
tail -2 gpurun_out/r02_bench_default.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err; cut -c1-600 gpurun_out/r02_bench_ref.json
for what in towers; do
  timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python tools/gpu_sanitize.py $what > gpurun_out/r02_memcheck_$what.log 2>&1; tail -3 gpurun_out/r02_memcheck_$what.log
  timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python tools/gpu_sanitize.py $what > gpurun_out/r02_racecheck_$what.log 2>&1; tail -3 gpurun_out/r02_racecheck_$what.log
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:score_sweep_tc -s 1 -c 1 -f -o gpurun_out/r02_sweep \
    python bench.py --steps 1 --warmup 1 --workload score --no-cpu-baseline --no-configs --no-robustness --no-d128 > gpurun_out/r02_ncu_sweep.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmm_rows_grouped -s 3 -c 1 -f -o gpurun_out/r02_spmm \
    python bench.py --steps 1 --warmup 1 --workload lightgcn --no-cpu-baseline --no-train --no-configs > gpurun_out/r02_ncu_spmm.log 2>&1
ls -la gpurun_out/r02_sweep.ncu-rep gpurun_out/r02_spmm.ncu-rep
