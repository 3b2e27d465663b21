#!/usr/bin/env python
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import csv, sys
from collections import defaultdict
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]; ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
tot, cnt = defaultdict(float), defaultdict(int)
for r in rows[1:]:
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    tot[r[ki]] += v; cnt[r[ki]] += 1
own = lambda k: "<unnamed>::" in k
T, To = sum(tot.values()), sum(v for k, v in tot.items() if own(k))
print(" ".join(sys.argv[2:]))
print("(per-launch times are cold-cache and serialised: compare shares, not absolutes; torch kernels are the synthetic data generators)")
print(f"total {T / 1e6:.1f} ms over {sum(cnt.values())} launches; coldrec_b200 kernels {To / 1e6:.1f} ms\n")
print(f"{'ns':>14s} {'share':>7s} {'share(own)':>10s} {'n':>5s}  kernel")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:28]:
    print(f"{v:14.0f} {100 * v / T:6.2f}% {(f'{100 * v / To:9.2f}%' if own(k) else ' ' * 10)} x{cnt[k]:4d}  {k[:90]}")
