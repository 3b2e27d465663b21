#!/bin/bash
# 1 GPU: parity suite after the d=128 / flag-bitmap / compaction-remap / balanced-warp changes, sanitizer on the new paths, robustness lines
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 --tb=short -x 2>&1 | grep -v "Warning\|^  warn\|return torch" | tail -12
for what in sweep spmm; do
  timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/gpu_sanitize.py $what > gpurun_out/r02_memcheck_$what.log 2>&1
  tail -5 gpurun_out/r02_memcheck_$what.log
done
timeout 600 compute-sanitizer --tool racecheck --print-limit 10 python tools/gpu_sanitize.py spmm > gpurun_out/r02_racecheck_spmm.log 2>&1; tail -3 gpurun_out/r02_racecheck_spmm.log
timeout 900 python bench.py --workload score --no-configs --no-cpu-baseline > gpurun_out/r02_bench_score_b.json 2> gpurun_out/r02_bench_score_b.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_score_b.json'))
print(d['value'], d['roofline']['achieved'], d['roofline']['frac'])
for r in d['robustness']: print(r['case'][:60], r['users_per_s'], r['sweep_tflops'], r['sweep_frac_of_tf32_sustained'], r['n_refined_all_steps'])
PY
tail -3 gpurun_out/r02_bench_score_b.err
