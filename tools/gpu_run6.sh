#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
for m in 0 1; do
  echo "mode $m" >> gpurun_out/modes6.log
  CR_TC_DEBUG_MODE=$m timeout 300 python bench.py --workload score --no-cpu-baseline --steps 3 --warmup 2 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['achieved'], d['roofline']['launch_ms'], d['clocks'])" >> gpurun_out/modes6.log
done
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench6.json 2> gpurun_out/bench6.err
tail -4 gpurun_out/pytest_gpu.log; cat gpurun_out/modes6.log; cat gpurun_out/bench6.json; tail -2 gpurun_out/bench6.err
