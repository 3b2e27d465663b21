mkdir -p gpurun_out
bash tools/gpu_check.sh 2>&1 | cut -c1-6000
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r01_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmm_rows_grouped -s 3 -c 1 -f -o gpurun_out/r01_spmm python bench.py --steps 1 --warmup 1 --workload lightgcn --no-cpu-baseline --no-train > gpurun_out/ncu_spmm.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/r01_launches.csv
