#!/bin/bash
# What the driver runs at round end, in one gpurun call: smoke, the -m gpu suite, the default bench and the reference arm.
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash tools/gpu_check.sh'
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 --tb=short 2>&1 | grep -v "Warning\|^  " | tail -15
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; cat gpurun_out/bench_default.json
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
