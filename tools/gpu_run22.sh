#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 600 --tb=short 2>&1 | grep -v "Warning\|^  " | tail -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 4 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
cat gpurun_out/bench_n2.json; tail -3 gpurun_out/bench_n2.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 4 --warmup 3 --workload score --shard users --no-cpu-baseline > gpurun_out/bench_n2_users.json 2> gpurun_out/bench_n2_users.err
cat gpurun_out/bench_n2_users.json; tail -3 gpurun_out/bench_n2_users.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err
cat gpurun_out/bench_ref_n2.json; tail -2 gpurun_out/bench_ref_n2.err
