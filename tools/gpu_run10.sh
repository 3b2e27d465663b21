#!/bin/bash
# full GPU test suite + A/B of the scorer (f2c8b3b's score_tc.cu vs HEAD) and of the SpMM variants, same box, interleaved
mkdir -p gpurun_out; rm -f gpurun_out/ab2.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x 2>&1 | tail -15 > gpurun_out/pytest_gpu_full.log
cat gpurun_out/pytest_gpu_full.log
L=coldrec_b200/csrc/libcoldrec_b200.so
cp $L /tmp/lib_new.so
run() { python bench.py --workload score --no-cpu-baseline "$@" 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['roofline']['launch_ms'], d['roofline']['achieved'], d['ms_per_step'], d['clocks']['sm_mhz'], d['check'])"; }
for rep in 1 2; do
 for lib in old new; do
  if [ $lib = old ]; then cp tools/lib_old_scorer.so $L; else cp /tmp/lib_new.so $L; fi
  echo "$lib default: $(run --steps 4 --warmup 3)" >> gpurun_out/ab2.log
  echo "$lib shard8:  $(run --steps 2 --warmup 2 --n-items 1250000 --users-per-step 606208)" >> gpurun_out/ab2.log
 done
done
cp /tmp/lib_new.so $L
for v in 0 1 2 3 4 0; do
  echo "spmm variant $v: $(CR_SPMM_VARIANT=$v python bench.py --workload lightgcn --no-cpu-baseline --steps 6 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); d=d.get('lightgcn',d); print(d['ms_per_step'], d['roofline']['achieved'], d['roofline'].get('rows_kernel_ms'))")" >> gpurun_out/ab2.log
done
cat gpurun_out/ab2.log
