#!/bin/bash
mkdir -p gpurun_out
for v in default seedskip default seedskip; do
  if [ $v == default ]; then unset CR_LIB_PATH; else export CR_LIB_PATH=$PWD/coldrec_b200/csrc/variants/lib_$v.so; fi
  timeout 300 python tools/gpu_flag_probe.py 2>&1 | grep '^{' | tee -a gpurun_out/r02_flag_probe.jsonl
done
unset CR_LIB_PATH
timeout 600 python -m pytest tests -m gpu -q --timeout 600 --tb=short -k "seed_phase or synthetic or evaluate_and or aldi or c1 or c3" 2>&1 | grep -v "Warning\|^  warn\|return torch" | tail -6
