"""Group an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: launches, total time, share.
usage: python scripts/summarize_launches.py <launch csv> [first_fraction_to_skip, default 0.5 = keep the second (warm) half]"""
import csv
import re
import sys

rows = list(csv.DictReader(l for l in open(sys.argv[1]) if l.startswith('"')))
skip = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
rows = rows[int(len(rows) * skip):]
scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}
agg = {}
for r in rows:
    name = re.sub(r"<.*", "", r["Kernel Name"].replace("void ", "")).split("(")[0].strip()
    t = float(r["Metric Value"].replace(",", "")) * scale[r["Metric Unit"]]
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += t
tot = sum(a[1] for a in agg.values())
print(f"{len(rows)} launches, {tot / 1e3:.2f} ms of kernel time (ncu per-launch times: serialised, cold caches)")
for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"  {name[:70]:70s} n={n:4d}  {t / 1e3:8.3f} ms  {100 * t / tot:5.1f}%")
