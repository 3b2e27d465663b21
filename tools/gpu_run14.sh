#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/exp14.log
timeout 600 python -m pytest tests/test_gpu_train.py -m gpu -q --timeout 600 --tb=short 2>&1 | grep -v "Warning\|^  " | tail -40 > gpurun_out/pytest_train.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -q --timeout 600 2>&1 | tail -8 > gpurun_out/pytest_score.log; cat gpurun_out/pytest_score.log
L=coldrec_b200/csrc/libcoldrec_b200.so
cp $L /tmp/lib_new.so
run() { python bench.py --workload score --no-cpu-baseline "$@" 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['roofline']['launch_ms'], d['roofline']['achieved'], d['ms_per_step'], d['clocks']['sm_mhz'], d['check'])"; }
for seed in 128 0 64; do
  echo "ms4 seed=$seed default: $(CR_TC_SEED_TILES=$seed run --steps 4 --warmup 3)" >> gpurun_out/exp14.log
  echo "ms4 seed=$seed shard8:  $(CR_TC_SEED_TILES=$seed run --steps 2 --warmup 2 --n-items 1250000 --users-per-step 606208)" >> gpurun_out/exp14.log
done
cp tools/lib_ms8.so $L
echo "ms8 seed=128 default: $(run --steps 4 --warmup 3)" >> gpurun_out/exp14.log
echo "ms8 seed=128 shard8:  $(run --steps 2 --warmup 2 --n-items 1250000 --users-per-step 606208)" >> gpurun_out/exp14.log
cp /tmp/lib_new.so $L
cat gpurun_out/exp14.log
CR_TC_DEBUG_MODE=8 timeout 300 python tools/gpu_timeline.py 37888 1250000 100 > gpurun_out/timeline_seed.log 2>&1; cat gpurun_out/timeline_seed.log
cat gpurun_out/pytest_train.log
