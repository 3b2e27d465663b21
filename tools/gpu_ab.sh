#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/ab.log
run() { python bench.py --workload score --no-cpu-baseline "$@" 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['roofline']['launch_ms'], d['roofline']['achieved'], d['clocks']['sm_mhz'])"; }
for rep in 1 2; do
 for seed in 0 128; do
  echo "seed=$seed default:  $(CR_TC_SEED_TILES=$seed run --steps 3 --warmup 2)" >> gpurun_out/ab.log
  echo "seed=$seed shard8:   $(CR_TC_SEED_TILES=$seed run --steps 2 --warmup 2 --n-items 1250000 --users-per-step 606208)" >> gpurun_out/ab.log
 done
done
cat gpurun_out/ab.log
