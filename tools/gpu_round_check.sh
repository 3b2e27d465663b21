#!/bin/bash
# Round-end check in one gpurun call: what the driver runs (smoke, -m gpu suite, default bench, reference arm) + the ncu launch list.
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash tools/gpu_round_check.sh'
mkdir -p gpurun_out
bash tools/gpu_check.sh 2>&1 | cut -c1-6000
timeout 200 python tools/gpu_train_probe.py 2>&1 | grep "C2" | head -1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r01_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
ls -la gpurun_out/r01_launches.csv
