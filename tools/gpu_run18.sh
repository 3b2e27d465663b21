#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 --tb=short 2>&1 | grep -v "Warning\|^  " | tail -40 > gpurun_out/pytest_gpu_full.log
tail -25 gpurun_out/pytest_gpu_full.log
( time timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err ) 2> gpurun_out/bench_default.time
cat gpurun_out/bench_default.json; tail -3 gpurun_out/bench_default.err; cat gpurun_out/bench_default.time
