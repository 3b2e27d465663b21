#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/small_launches.csv python tools/gpu_small_probe.py > gpurun_out/small_probe.log 2>&1
tail -2 gpurun_out/small_probe.log
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/small_launches.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
last=rows[-40:]
for r in last: print(r[ki][:70], r[vi])
PY
timeout 600 python tools/gpu_train_probe.py 2>&1 | head -2
