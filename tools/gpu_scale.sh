#!/bin/bash
# SCALE-like lines at N GPUs (what the driver runs at round end) + the alternative layouts:  tools/gpu_scale.sh <N>
N=${1:-2}
mkdir -p gpurun_out
run() { timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N "$@"; }
run --steps 20 --warmup 5 > gpurun_out/r02_bench_n${N}.json 2> gpurun_out/r02_bench_n${N}.err; tail -c 4000 gpurun_out/r02_bench_n${N}.json; grep -v "^\*\*\*\|OMP_NUM\|Warning\|warn" gpurun_out/r02_bench_n${N}.err | tail -3
if [ "$N" == "4" ] || [ "$N" == "8" ]; then
  run --steps 10 --warmup 3 --workload score --min-shard-items 2500000 > gpurun_out/r02_bench_n${N}_shards4.json 2> gpurun_out/r02_bench_n${N}_shards4.err; tail -c 1500 gpurun_out/r02_bench_n${N}_shards4.json
fi
if [ "$N" == "8" ]; then
  run --steps 10 --warmup 3 --workload score --shard items > gpurun_out/r02_bench_n8_shard_items.json 2> gpurun_out/r02_bench_n8_shard_items.err; tail -c 1800 gpurun_out/r02_bench_n8_shard_items.json
  run --steps 10 --warmup 3 --workload lightgcn --prop-result users > gpurun_out/r02_bench_n8_prop_users.json 2> gpurun_out/r02_bench_n8_prop_users.err; tail -c 1500 gpurun_out/r02_bench_n8_prop_users.json
fi
