#!/bin/bash
# round 2, call 1: scorer tile-geometry A/B at the sharded shapes + compute-sanitizer on small shapes
mkdir -p gpurun_out
for v in default bn64a3 bn64a3s6 bn32a6; do
  if [ $v == default ]; then unset CR_LIB_PATH; else export CR_LIB_PATH=$PWD/coldrec_b200/csrc/variants/lib_$v.so; fi
  timeout 300 python tools/gpu_shard_probe.py 1,8,4 2>&1 | grep '^{' | tee -a gpurun_out/r02_shard_probe.jsonl
done
unset CR_LIB_PATH
for what in sweep spmm train towers; do
  timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python tools/gpu_sanitize.py $what > gpurun_out/r02_memcheck_$what.log 2>&1
  tail -4 gpurun_out/r02_memcheck_$what.log
done
for what in spmm train; do
  timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python tools/gpu_sanitize.py $what > gpurun_out/r02_racecheck_$what.log 2>&1
  tail -4 gpurun_out/r02_racecheck_$what.log
done
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python tools/gpu_sanitize.py sweep > gpurun_out/r02_racecheck_sweep.log 2>&1
tail -4 gpurun_out/r02_racecheck_sweep.log
timeout 600 compute-sanitizer --tool synccheck --print-limit 20 python tools/gpu_sanitize.py sweep > gpurun_out/r02_synccheck_sweep.log 2>&1
tail -4 gpurun_out/r02_synccheck_sweep.log
