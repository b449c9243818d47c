"""Summarise an `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv` log of one conv kernel's
launches into profiles/conv_traffic.json (second half of the log = the warm decoder step).  bench.py reads the entry keyed
"<workload>:<batch>:<precision>:<kernel>" as `roofline.traffic`.
usage: python scripts/summarize_traffic.py <ncu csv log> <json> <workload> <batch> <precision> <kernel: conv_tc3 | conv_tc2>"""
import csv
import json
import os
import sys

src, dst, workload, batch, prec, kern = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]), sys.argv[5], sys.argv[6]
lines = [l for l in open(src) if l.startswith('"')]
rows = list(csv.DictReader(lines))
per = {}
for r in rows:
    per.setdefault(r["ID"], {})[r["Metric Name"]] = (float(r["Metric Value"].replace(",", "")), r["Metric Unit"])
ids = sorted(per, key=int)
ids = ids[len(ids) // 2:]
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "nsecond": 1e-9, "usecond": 1e-6}
rd = sum(per[i]["dram__bytes_read.sum"][0] * scale[per[i]["dram__bytes_read.sum"][1]] for i in ids)
wr = sum(per[i]["dram__bytes_write.sum"][0] * scale[per[i]["dram__bytes_write.sum"][1]] for i in ids)
t = sum(per[i]["gpu__time_duration.sum"][0] * scale[per[i]["gpu__time_duration.sum"][1]] for i in ids)
out = {"kernel": f"pdae::{kern}_kernel", "workload": workload, "batch": batch, "precision": prec, "launches": len(ids),
       "dram_bytes_read_per_launch": rd / len(ids), "dram_bytes_write_per_launch": wr / len(ids),
       "traffic_bytes_per_launch": (rd + wr) / len(ids), "ncu_time_us_per_launch_cold": 1e6 * t / len(ids),
       "source": f"ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none "
                 f"-k regex:{kern} python scripts/ncu_step.py {workload} {batch} {prec} 2 (second of two decoder steps)"}
allj = json.load(open(dst)) if os.path.exists(dst) else {}
allj[f"{workload}:{batch}:{prec}:{kern}"] = out
json.dump(allj, open(dst, "w"), indent=1)
print(out)
