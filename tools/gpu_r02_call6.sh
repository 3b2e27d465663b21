#!/bin/bash
# 1 GPU: balanced vs fixed rows per warp at C4 (N=1), new GPU tests, default bench with the C1-C3 / robustness lines
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 800 --tb=short -x -k "spmm or propagate or lightgcn or train or c2 or dropin or reevaluation or bench" 2>&1 | grep -v "Warning\|^  warn" | tail -12
timeout 300 python bench.py --workload lightgcn --steps 10 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/r02_lg_n1_balanced.json 2>gpurun_out/r02_lg_n1_balanced.err; cut -c1-1500 gpurun_out/r02_lg_n1_balanced.json
CR_SPMM_FIXED_ROWS=1 timeout 300 python bench.py --workload lightgcn --steps 10 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/r02_lg_n1_fixed.json 2>gpurun_out/r02_lg_n1_fixed.err; cut -c1-700 gpurun_out/r02_lg_n1_fixed.json
timeout 900 python bench.py --workload score > gpurun_out/r02_bench_score_full.json 2> gpurun_out/r02_bench_score_full.err; tail -c 6000 gpurun_out/r02_bench_score_full.json; tail -5 gpurun_out/r02_bench_score_full.err
