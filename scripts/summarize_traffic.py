"""Summarise an `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv` log of conv_tc2 launches
into profiles/conv_tc2_traffic.json (second half of the log = the warm decoder step)."""
import csv
import json
import sys

src, dst, workload, batch = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4])
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
out = {"kernel": "pdae::conv_tc2_kernel", "workload": workload, "batch": batch, "launches": len(ids),
       "dram_bytes_read_per_launch": rd / len(ids), "dram_bytes_write_per_launch": wr / len(ids),
       "traffic_bytes_per_launch": (rd + wr) / len(ids), "ncu_time_us_per_launch_cold": 1e6 * t / len(ids),
       "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none "
                 "-k regex:conv_tc2 python scripts/ncu_step.py (second of two decoder steps)"}
json.dump(out, open(dst, "w"), indent=1)
print(out)
