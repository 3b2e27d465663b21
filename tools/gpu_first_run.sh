#!/bin/bash
# first GPU session: exact-path tests, TC tests (separate process: a trap kills the context), probes
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -k "not tf32 and not tc_raw" --timeout 300 2>&1 | tail -60 > gpurun_out/pytest_exact.log
timeout 600 python -m pytest tests -m gpu -q -k "tc_raw" --timeout 120 2>&1 | tail -40 > gpurun_out/pytest_tcprobe.log
timeout 900 python -m pytest tests -m gpu -q -k "tf32" --timeout 300 2>&1 | tail -80 > gpurun_out/pytest_tf32.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
timeout 900 python tools/gpu_probe.py score > gpurun_out/probe_score.log 2>&1
timeout 900 python tools/gpu_probe.py spmm > gpurun_out/probe_spmm.log 2>&1
tail -5 gpurun_out/pytest_exact.log gpurun_out/pytest_tcprobe.log gpurun_out/pytest_tf32.log gpurun_out/smoke.log gpurun_out/probe_score.log gpurun_out/probe_spmm.log
