#!/usr/bin/env python
"""Per-kernel shares from an ncu launch list (developer tool):
    ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline
    python tools/launch_shares.py launches.csv > launch_shares.csv
cuBLAS DGEMM launches of bench.py's in-run fp64 peak measurement and torch fill / copy kernels are listed but the shares are over
the repo's own kernels (k_* / kl_*)."""
import csv
import re
import sys
from collections import defaultdict

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if not l.startswith("==")) if r]
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = defaultdict(float), defaultdict(int)
for r in rows[1:]:
    if len(r) <= iv:
        continue
    name = re.sub(r"\(.*", "", r[ik]).replace("void ", "").replace("<unnamed>::", "")
    v = float(r[iv].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3}.get(r[iu], 1.0)
    tot[name] += v
    cnt[name] += 1
own = {k: v for k, v in tot.items() if k.startswith(("k_", "kl_"))}
s = sum(own.values())
print("kernel,launches,total_us,share_pct_of_own_kernels")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"{k},{cnt[k]},{v:.1f},{100 * v / s:.2f}" if k in own else f"{k},{cnt[k]},{v:.1f},")
