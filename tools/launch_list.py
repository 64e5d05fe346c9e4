"""Last step of an ncu launch list (gpu__time_duration.sum csv): python tools/launch_list.py FILE.csv"""
import csv, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
rows = list(csv.DictReader(lines))
names = [(r["Kernel Name"], float(r["Metric Value"])) for r in rows]
idx = [i for i, (n, v) in enumerate(names) if n.startswith(("k_advect", "k_keys"))]
tot = 0
for n, v in names[idx[-1]:]:
    print(f"{n[:44]:44s} {v / 1000:9.1f} us"); tot += v
print("sum", tot / 1000, "us")
