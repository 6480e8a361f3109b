"""Print the per-launch durations of the LAST evaluation in an ncu launch list (csv from --metrics gpu__time_duration.sum)."""
import csv, re, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
names = [re.sub(r"\(.*", "", r[4]).replace("gps::", "").replace("void ", "") for r in rows]
idx = [i for i, n in enumerate(names) if n.startswith("hyp_prep")]
a, b = idx[-2], idx[-1]
tot = 0.0
for i in range(a, b):
    t = int(rows[i][14]) / 1000
    tot += t
    print(i - a, names[i][:64], rows[i][7], rows[i][8], f"{t:.1f}us")
print("total", round(tot, 1), "us")
