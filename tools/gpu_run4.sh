#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -k "tc_raw or tf32 or sharding" 2>&1 | tail -30 > gpurun_out/pytest_tc.log
timeout 600 python bench.py --workload score --no-cpu-baseline > gpurun_out/bench_score.json 2> gpurun_out/bench_score.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:score_sweep_tc -s 1 -c 1 -o gpurun_out/prof_sweep3 \
    python bench.py --steps 1 --warmup 1 --workload score --no-cpu-baseline > gpurun_out/ncu_sweep3.log 2>&1
tail -5 gpurun_out/pytest_tc.log; cat gpurun_out/bench_score.json; tail -3 gpurun_out/bench_score.err
