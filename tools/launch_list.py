"""Per-kernel summary of an ncu launch list (gpu__time_duration.sum csv): python tools/launch_list.py FILE.csv [OUT.csv]
Launch count, median and total duration per kernel name, plus the kernels of the last complete timestep."""
import csv, sys, statistics, collections
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
rows = list(csv.DictReader(lines))
names = [(r["Kernel Name"].split("(")[0].replace("void ", "").split("<")[0], float(r["Metric Value"]) / 1000) for r in rows]
per = collections.OrderedDict()
for n, v in names: per.setdefault(n, []).append(v)
out = [("kernel", "launches", "median_us", "total_us")]
for n, v in per.items(): out.append((n, len(v), round(statistics.median(v), 1), round(sum(v), 1)))
for o in out: print("%-28s %9s %10s %10s" % o)
if len(sys.argv) > 2:
    csv.writer(open(sys.argv[2], "w")).writerows(out)
starts = [i for i, (n, v) in enumerate(names) if n in ("k_advect", "k_keys")]
steps = [names[a:b] for a, b in zip(starts, starts[1:]) if any(n.startswith("k_detect") for n, v in names[a:b])]
if steps:
    st = min(steps, key=lambda q: sum(v for n, v in q))
    print("fastest complete timestep of the run (%d launches, %.1f us; each kernel cold after ncu's cache flush):" % (len(st), sum(v for n, v in st)))
    for n, v in st: print("   %-26s %9.1f us" % (n, v))
