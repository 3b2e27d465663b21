#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/exp15.log
L=coldrec_b200/csrc/libcoldrec_b200.so
cp $L /tmp/lib_new.so
run() { python bench.py --workload score --no-cpu-baseline "$@" 2>>gpurun_out/exp15.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['roofline']['launch_ms'], d['roofline']['achieved'], d['ms_per_step'], d['clocks']['sm_mhz'])"; }
nvidia-smi --query-gpu=name,power.limit,clocks.max.sm,clocks.max.mem,temperature.gpu --format=csv >> gpurun_out/exp15.log
for rep in 1 2; do
for lib in v12 new ms8; do
  if [ $lib = new ]; then cp /tmp/lib_new.so $L; else cp tools/lib_$lib.so $L; fi
  echo "$lib default: $(run --steps 4 --warmup 3)" >> gpurun_out/exp15.log
  echo "$lib shard8:  $(run --steps 2 --warmup 2 --n-items 1250000 --users-per-step 606208)" >> gpurun_out/exp15.log
done
done
cp /tmp/lib_new.so $L
cat gpurun_out/exp15.log; tail -5 gpurun_out/exp15.err
