#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/timeline.log
timeout 600 python -m pytest tests/test_gpu_train.py -m gpu -q --timeout 600 --tb=short 2>&1 | tail -4
for args in "37888 1250000 100" "37888 1250000 0" "75776 10000000 100"; do
  for m in 8 9; do CR_TC_DEBUG_MODE=$m timeout 300 python tools/gpu_timeline.py $args >> gpurun_out/timeline.log 2>&1; done
done
cat gpurun_out/timeline.log
