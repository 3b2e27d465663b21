#!/bin/bash
# 1 GPU: TC towers with K-block drains (parity), sweep A/B on one box (current / head-seeded / previous revision), robustness lines, C3 line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 800 --tb=short -k "tower or c3 or golden or seed_phase or synthetic" 2>&1 | grep -v "Warning\|^  warn\|return torch" | tail -12
for v in default headseed prev default headseed prev; do
  if [ $v == default ]; then unset CR_LIB_PATH; else export CR_LIB_PATH=$PWD/coldrec_b200/csrc/variants/lib_$v.so; fi
  timeout 300 python tools/gpu_shard_probe.py 1,8 2>&1 | grep '^{' | tee -a gpurun_out/r02_sweep_ab2.jsonl
done
unset CR_LIB_PATH
timeout 900 python bench.py --workload score --configs C3 --no-cpu-baseline > gpurun_out/r02_bench_score_c.json 2> gpurun_out/r02_bench_score_c.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_score_c.json'))
print(d['value'], d['roofline']['achieved'], d['roofline']['frac'])
for r in d['robustness']: print(r['case'][:60], r['users_per_s'], r['sweep_tflops'], r['sweep_frac_of_tf32_sustained'], r['n_refined_all_steps'])
print(json.dumps(d['extra']['C3'])[:1500])
PY
tail -3 gpurun_out/r02_bench_score_c.err
