#!/bin/bash
mkdir -p gpurun_out
for st in "8 3" "20 5" "40 5"; do set -- $st
  timeout 300 python bench.py --workload lightgcn --steps $1 --warmup $2 --no-cpu-baseline --no-configs --no-train 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('lightgcn only steps',d['steps'], d['ms_per_step'], d['roofline']['rows_kernel_ms'], d['roofline']['share_of_step'], d['gpu_launches'])"
done
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-configs --no-robustness --no-d128 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); lg=d['lightgcn']; print('both, no extras steps',lg['steps'], lg['ms_per_step'], lg['roofline']['rows_kernel_ms'], lg['roofline']['share_of_step'])"
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-configs 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); lg=d['lightgcn']; print('both + robustness + d128 steps',lg['steps'], lg['ms_per_step'], lg['roofline']['rows_kernel_ms'], lg['roofline']['share_of_step'])"
timeout 300 python tools/gpu_tower_probe.py 2>&1 | grep "^{" | cut -c1-330
timeout 300 python -m pytest tests -m gpu -q --timeout 600 --tb=short -k "tower or knn or c3 or golden" 2>&1 | tail -3
