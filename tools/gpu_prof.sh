#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r01_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:score_sweep_tc -s 1 -c 1 -o gpurun_out/r01_sweep \
    python bench.py --steps 1 --warmup 1 --workload score --no-cpu-baseline > gpurun_out/ncu_sweep.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmm_rows_grouped -s 3 -c 1 -o gpurun_out/r01_spmm \
    python bench.py --steps 1 --warmup 1 --workload lightgcn --no-cpu-baseline > gpurun_out/ncu_spmm.log 2>&1
timeout 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
ls -la gpurun_out; cat gpurun_out/bench_final.json
