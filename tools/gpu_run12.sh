#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/exp12.log
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -q --timeout 600 --tb=short 2>&1 | grep -v "^  \|Warning" | tail -60 > gpurun_out/pytest_train.log
cat gpurun_out/pytest_train.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -q --timeout 600 -x 2>&1 | tail -5 > gpurun_out/pytest_score.log; cat gpurun_out/pytest_score.log
run() { python bench.py --workload score --no-cpu-baseline "$@" 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['roofline']['launch_ms'], d['roofline']['achieved'], d['ms_per_step'], d['clocks']['sm_mhz'], d['check'])"; }
echo "default: $(run --steps 4 --warmup 3)" >> gpurun_out/exp12.log
echo "shard8: $(run --steps 2 --warmup 2 --n-items 1250000 --users-per-step 606208)" >> gpurun_out/exp12.log
echo "shard8 dbg=1: $(CR_TC_DEBUG_MODE=1 run --steps 2 --warmup 2 --n-items 1250000 --users-per-step 606208)" >> gpurun_out/exp12.log
echo "1wave x1.25M: $(run --steps 4 --warmup 2 --n-items 1250000 --users-per-step 37888)" >> gpurun_out/exp12.log
echo "default: $(run --steps 4 --warmup 3)" >> gpurun_out/exp12.log
cat gpurun_out/exp12.log
