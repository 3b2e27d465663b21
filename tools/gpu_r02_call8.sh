#!/bin/bash
# 1 GPU: tensor-core towers (parity + sanitizer), same-box A/B of the sweep against the previous revision, C3 line with TC towers
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 800 --tb=short -x -k "tower or c3 or golden or dropin" 2>&1 | grep -v "Warning\|^  warn\|return torch" | tail -15
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python tools/gpu_sanitize.py towers > gpurun_out/r02_memcheck_towers.log 2>&1; tail -4 gpurun_out/r02_memcheck_towers.log
for v in default prev default prev; do
  if [ $v == default ]; then unset CR_LIB_PATH; else export CR_LIB_PATH=$PWD/coldrec_b200/csrc/variants/lib_$v.so; fi
  timeout 300 python tools/gpu_shard_probe.py 1,8 2>&1 | grep '^{' | tee -a gpurun_out/r02_sweep_ab.jsonl
done
unset CR_LIB_PATH
timeout 600 python bench.py --workload score --no-robustness --configs C3 --steps 3 --warmup 3 > gpurun_out/r02_bench_c3.json 2> gpurun_out/r02_bench_c3.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_c3.json'))
print(json.dumps(d['extra']['C3'])[:1800])
PY
tail -3 gpurun_out/r02_bench_c3.err
