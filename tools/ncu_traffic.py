#!/usr/bin/env python
"""Write profiles/r02_traffic.json from `ncu --set full` captures:  tools/ncu_traffic.py <kernel_key>=<rep>:<src.cu> ...
e.g.  tools/ncu_traffic.py score_sweep_tc_kernel=gpurun_out/r02_sweep.ncu-rep:score_tc.cu spmm_rows_grouped_kernel=gpurun_out/r02_spmm.ncu-rep:spmm.cu
Each entry records dram__bytes_read.sum + dram__bytes_write.sum of the capture's first launch and the sha256 of the kernel
source it was taken on (bench.py withholds the figure when the source has changed since)."""
import csv, hashlib, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
out_path = os.path.join(ROOT, "profiles", "r02_traffic.json")
out = json.load(open(out_path)) if os.path.exists(out_path) else {}
for arg in sys.argv[1:]:
    key, rest = arg.split("=", 1)
    rep, src = rest.rsplit(":", 1)
    rows = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
    hdr, units, first = rows[0], rows[1], rows[2]
    d = {h: (v, u) for h, v, u in zip(hdr, first, units)}
    tot = sum(float(d[m][0].replace(",", "")) * UNIT[d[m][1]] for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    out[key] = {"bytes_per_launch": tot, "kernel": d["Kernel Name"][0][:80], "capture": os.path.basename(rep),
                "source": f"coldrec_b200/csrc/{src}", "source_sha256": hashlib.sha256(open(os.path.join(ROOT, "coldrec_b200", "csrc", src), "rb").read()).hexdigest()}
json.dump(out, open(out_path, "w"), indent=1)
print(json.dumps(out, indent=1))
