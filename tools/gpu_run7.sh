#!/bin/bash
mkdir -p gpurun_out
# the per-GPU work of the 8-GPU run (606,208 users x 1.25M-item shard) on one GPU
timeout 600 python bench.py --workload score --no-cpu-baseline --n-items 1250000 --users-per-step 606208 --steps 3 --warmup 2 > gpurun_out/bench_shard8.json 2> gpurun_out/bench_shard8.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:score_sweep_tc -s 1 -c 1 -o gpurun_out/prof_shard8 \
    python bench.py --steps 1 --warmup 1 --workload score --no-cpu-baseline --n-items 1250000 --users-per-step 606208 > gpurun_out/ncu_shard8.log 2>&1
cat gpurun_out/bench_shard8.json | cut -c1-1800
