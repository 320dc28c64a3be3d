"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel: launches, total and share of time."""
import csv, re, sys
from collections import defaultdict
path = sys.argv[1]
rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10]
hdr = rows[0]
iK, iM, iV, iU = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    if r[iM] != "gpu__time_duration.sum":
        continue
    v = float(r[iV].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[iU], 1.0)
    name = re.sub(r"\(.*", "", r[iK]).replace("void ", "").replace("ivlm::", "")
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
print(f"| kernel | launches | total ms | share |\n|---|---|---|---|")
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k}` | {n} | {us / 1e3:.2f} | {us / tot:.3f} |")
print(f"\ntotal {sum(v[0] for v in agg.values())} launches, {tot / 1e3:.1f} ms of kernel time (cold-cache, serialised: compare shares, not absolutes)")
