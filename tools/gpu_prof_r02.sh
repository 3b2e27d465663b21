#!/bin/bash
# round-2 profiles in one 1-GPU call: launch list of the default bench + one `ncu --set full` capture of each dominant kernel
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-configs --no-robustness --no-d128 > gpurun_out/r02_bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:score_sweep_tc -s 1 -c 1 -f -o gpurun_out/r02_sweep \
    python bench.py --steps 1 --warmup 1 --workload score --no-cpu-baseline --no-configs --no-robustness --no-d128 > gpurun_out/r02_ncu_sweep.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmm_rows_grouped -s 3 -c 1 -f -o gpurun_out/r02_spmm \
    python bench.py --steps 1 --warmup 1 --workload lightgcn --no-cpu-baseline --no-train --no-configs > gpurun_out/r02_ncu_spmm.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tower_layer_tc -s 8 -c 3 -f -o gpurun_out/r02_tower \
    python tools/gpu_tower_probe.py > gpurun_out/r02_ncu_tower.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:score_sweep_tc -s 1 -c 1 -f -o gpurun_out/r02_sweep128 \
    python tools/gpu_d128_probe.py > gpurun_out/r02_ncu_sweep128.log 2>&1
ls -la gpurun_out/*.ncu-rep
