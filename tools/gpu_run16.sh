#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/exp16.log gpurun_out/exp16.err
L=coldrec_b200/csrc/libcoldrec_b200.so
cp $L /tmp/lib_new.so
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -q --timeout 600 2>&1 | tail -3 > gpurun_out/pytest_score.log; cat gpurun_out/pytest_score.log
run() { python bench.py --workload score --no-cpu-baseline "$@" 2>>gpurun_out/exp16.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['roofline']['launch_ms'], d['roofline']['achieved'], d['ms_per_step'], d['clocks']['sm_mhz'], d['check']['n_refined_last_step'])"; }
for rep in 1 2; do
for lib in v12 new ms8; do
  if [ $lib = new ]; then cp /tmp/lib_new.so $L; else cp tools/lib_$lib.so $L; fi
  echo "$lib default: $(run --steps 4 --warmup 3)" >> gpurun_out/exp16.log
  echo "$lib shard8:  $(run --steps 2 --warmup 2 --n-items 1250000 --users-per-step 606208)" >> gpurun_out/exp16.log
done
done
cp /tmp/lib_new.so $L
echo "new seed=0 default: $(CR_TC_SEED_TILES=0 run --steps 4 --warmup 3)" >> gpurun_out/exp16.log
echo "new seed=0 shard8:  $(CR_TC_SEED_TILES=0 run --steps 2 --warmup 2 --n-items 1250000 --users-per-step 606208)" >> gpurun_out/exp16.log
cat gpurun_out/exp16.log; grep -v "^  \|Traceback\|json" gpurun_out/exp16.err | tail -5
CR_TC_DEBUG_MODE=8 timeout 300 python tools/gpu_timeline.py 37888 1250000 100 > gpurun_out/timeline_seed.log 2>&1; cat gpurun_out/timeline_seed.log
