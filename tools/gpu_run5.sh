#!/bin/bash
mkdir -p gpurun_out
for m in 0 4 1 5 2; do
  echo "mode $m" >> gpurun_out/modes.log
  CR_TC_DEBUG_MODE=$m timeout 300 python bench.py --workload score --no-cpu-baseline --steps 3 --warmup 2 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['achieved'], d['roofline']['launch_ms'], d['clocks'])" >> gpurun_out/modes.log
done
cat gpurun_out/modes.log
