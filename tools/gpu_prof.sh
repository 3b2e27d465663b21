#!/bin/bash
# validation of the SpMM threshold change + profiles of the final kernels (launch list, one --set full capture each)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 900 --tb=short -k "spmm or propagate or lightgcn or train or c2 or databuilder or knn" 2>&1 | grep -v "Warning\|^  " | tail -8
timeout 600 python tools/gpu_train_probe.py 2>&1 | tee gpurun_out/train_probe.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r01_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:score_sweep_tc -s 1 -c 1 -f -o gpurun_out/r01_sweep \
    python bench.py --steps 1 --warmup 1 --workload score --no-cpu-baseline > gpurun_out/ncu_sweep.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmm_rows_grouped -s 3 -c 1 -f -o gpurun_out/r01_spmm \
    python bench.py --steps 1 --warmup 1 --workload lightgcn --no-cpu-baseline --no-train > gpurun_out/ncu_spmm.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"adam_step|bpr_backward" -s 2 -c 2 -f -o gpurun_out/r01_train \
    python bench.py --steps 1 --warmup 1 --workload lightgcn --no-cpu-baseline > gpurun_out/ncu_train.log 2>&1
ls -la gpurun_out/*.ncu-rep
