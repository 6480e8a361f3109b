"""Per-step DRAM traffic and kernel shares from an ncu csv with gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum
(one row per launch and metric): python scripts/traffic_summary.py launches.csv fits_per_step launches_per_eval [out.json]
Takes the LAST evaluation in the file (its last `launches_per_eval` launches: both stream groups, enqueued one after the other)."""
import csv, json, re, sys
from collections import defaultdict
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
fits = int(sys.argv[2])
launch = {}
order = []
for r in rows:
    i = int(r[0])
    if i not in launch:
        launch[i] = {"name": re.sub(r"\(.*", "", r[4]).replace("gps::", "").replace("void ", ""), "grid": r[8]}
        order.append(i)
    v = float(r[14].replace(",", ""))
    unit = r[13]
    if r[12].startswith("gpu__time"):
        launch[i]["us"] = v / 1e3 if unit == "ns" else (v if unit == "us" else v * 1e3)
    else:
        mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
        launch[i]["bytes"] = launch[i].get("bytes", 0.0) + v * mult
per = int(sys.argv[3])
sel = order[-per:]
tot_us = sum(launch[i]["us"] for i in sel)
tot_b = sum(launch[i].get("bytes", 0.0) for i in sel)
by = defaultdict(lambda: [0.0, 0, 0.0])
for i in sel:
    k = launch[i]["name"][:90]
    by[k][0] += launch[i]["us"]; by[k][1] += 1; by[k][2] += launch[i].get("bytes", 0.0)
print(f"one step: {len(sel)} launches, {tot_us/1e3:.2f} ms serialised under ncu, {tot_b/1e9:.2f} GB DRAM")
print("| share | time | launches | DRAM | achieved DRAM GB/s | kernel |\n|---:|---:|---:|---:|---:|---|")
for k, (us, cnt, byt) in sorted(by.items(), key=lambda kv: -kv[1][0]):
    print(f"| {100*us/tot_us:.1f} % | {us/1e3:.3f} ms | {cnt} | {byt/1e9:.2f} GB | {byt/us/1e3 if us else 0:.0f} | `{k}` |")
if len(sys.argv) > 4:
    json.dump({"fits_per_step": fits, "traffic_bytes_per_step": tot_b, "launches_per_step": len(sel), "ncu_ms_per_step": tot_us / 1e3,
               "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none python scripts/profile_batch.py %d (graphs off), one full evaluation of both stream groups" % fits},
              open(sys.argv[4], "w"), indent=1)
