#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -k "tc_raw or tf32 or sharding or seed_phase" 2>&1 | tail -12 > gpurun_out/pytest_tc.log
timeout 600 python bench.py --workload score --no-cpu-baseline --steps 4 --warmup 3 > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err
timeout 600 python bench.py --workload score --no-cpu-baseline --n-items 1250000 --users-per-step 606208 --steps 3 --warmup 2 > gpurun_out/bench_b.json 2>> gpurun_out/bench_a.err
cat gpurun_out/pytest_tc.log; for f in a b; do python -c "
import json,sys; d=json.load(open('gpurun_out/bench_$f.json')); print(d['value'], d['ms_per_step'], d['roofline']['achieved'], d['roofline']['launch_ms'], d['clocks'], d['check'])"; done; tail -3 gpurun_out/bench_a.err
