#!/bin/bash
# training-side kernels: GPU tests + probe; scorer debug-mode experiments for the many-users / small-shard regime
mkdir -p gpurun_out; rm -f gpurun_out/exp11.log
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -q --timeout 600 2>&1 | tail -25 > gpurun_out/pytest_train.log
cat gpurun_out/pytest_train.log
timeout 600 python tools/gpu_train_probe.py > gpurun_out/train_probe.log 2>&1; cat gpurun_out/train_probe.log
run() { python bench.py --workload score --no-cpu-baseline "$@" 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['roofline']['launch_ms'], d['roofline']['achieved'], d['ms_per_step'], d['clocks']['sm_mhz'])"; }
for m in 0 1 2; do
  echo "shard8 dbg=$m: $(CR_TC_DEBUG_MODE=$m run --steps 2 --warmup 2 --n-items 1250000 --users-per-step 606208)" >> gpurun_out/exp11.log
done
echo "default dbg=1: $(CR_TC_DEBUG_MODE=1 run --steps 3 --warmup 2)" >> gpurun_out/exp11.log
echo "default dbg=2: $(CR_TC_DEBUG_MODE=2 run --steps 3 --warmup 2)" >> gpurun_out/exp11.log
echo "1wave x1.25M dbg=0: $(run --steps 4 --warmup 2 --n-items 1250000 --users-per-step 37888)" >> gpurun_out/exp11.log
echo "2wave x1.25M dbg=0: $(run --steps 4 --warmup 2 --n-items 1250000 --users-per-step 75776)" >> gpurun_out/exp11.log
echo "4wave x1.25M dbg=0: $(run --steps 4 --warmup 2 --n-items 1250000 --users-per-step 151552)" >> gpurun_out/exp11.log
echo "2wave x2.5M dbg=0: $(run --steps 4 --warmup 2 --n-items 2500000 --users-per-step 75776)" >> gpurun_out/exp11.log
cat gpurun_out/exp11.log
