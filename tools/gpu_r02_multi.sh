#!/bin/bash
# multi-GPU validation + A/B of the peer-store paths: tools/gpu_r02_multi.sh <N> [tests]
N=${1:-2}
mkdir -p gpurun_out
if [ "$2" == "tests" ]; then
  timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 800 --tb=short -x 2>&1 | grep -v "Warning\|^  warn" | tail -15
fi
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@"; }
run --steps 10 --warmup 3 --workload lightgcn > gpurun_out/r02_lg_n${N}_tma.json 2> gpurun_out/r02_lg_n${N}_tma.err; tail -c 1800 gpurun_out/r02_lg_n${N}_tma.json; tail -3 gpurun_out/r02_lg_n${N}_tma.err
run --steps 10 --warmup 3 --workload lightgcn --peer-store st > gpurun_out/r02_lg_n${N}_st.json 2> gpurun_out/r02_lg_n${N}_st.err; tail -c 1500 gpurun_out/r02_lg_n${N}_st.json
run --steps 5 --warmup 3 --workload score > gpurun_out/r02_score_n${N}.json 2> gpurun_out/r02_score_n${N}.err; tail -c 900 gpurun_out/r02_score_n${N}.json; tail -3 gpurun_out/r02_score_n${N}.err
