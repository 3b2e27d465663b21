#!/bin/bash
# tests + bench + ncu launch list + full captures of the two dominant kernels
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:score_sweep_tc -s 1 -c 1 -o gpurun_out/prof_sweep \
    python bench.py --steps 1 --warmup 1 --workload score --no-cpu-baseline > gpurun_out/ncu_sweep.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmm_rows -s 3 -c 1 -o gpurun_out/prof_spmm \
    python bench.py --steps 1 --warmup 1 --workload lightgcn --no-cpu-baseline > gpurun_out/ncu_spmm.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err; cat gpurun_out/bench_ref.json; tail -2 gpurun_out/ncu_sweep.log gpurun_out/ncu_spmm.log; ls -la gpurun_out
