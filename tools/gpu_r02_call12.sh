#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 800 --tb=short -k "knn or spmm or propagate or lightgcn or bench" 2>&1 | grep -v "Warning\|^  warn\|return torch" | tail -8
timeout 300 python tools/gpu_knn_probe.py 2>&1 | grep "^{\|Error" | tee gpurun_out/r02_knn_probe.jsonl
bash tools/gpu_prof_r02.sh 2>&1 | tail -8
