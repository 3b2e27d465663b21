#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -q --timeout 600 --tb=short 2>&1 | grep -v "Warning\|^  " | tail -30
timeout 600 python tools/gpu_train_probe.py 2>&1 | tee gpurun_out/train_probe.log
