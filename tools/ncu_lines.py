#!/usr/bin/env python
"""Per-source-line share of executed warp instructions / stall samples of one kernel in an ncu report
(needs --import-source on and -lineinfo).   python tools/ncu_lines.py REPORT.ncu-rep KERNEL [min_pct]"""
import csv
import io
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
minp = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern, "--print-source=cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
h = rows[hi]
iI, iS, iT = h.index("Instructions Executed"), h.index("# Samples"), h.index("Thread Instructions Executed")
stalls = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
agg, src = {}, {}
for r in rows[hi + 1:]:
    if r and r[0] == "Line No":
        break                      # next launch of the same kernel
    try:
        ln = int(r[0]); inst = int(r[iI]); samp = int(r[iS]); thr = int(r[iT])
    except Exception:
        continue
    src[ln] = r[1]
    a = agg.setdefault(ln, [0, 0, 0, {}])
    a[0] += inst; a[1] += samp; a[2] += thr
    for c in stalls:
        try:
            a[3][c] = a[3].get(c, 0) + int(r[h.index(c)])
        except Exception:
            pass
tot, ts = sum(a[0] for a in agg.values()), sum(a[1] for a in agg.values())
print("total warp instructions", tot, "samples", ts)
for ln in sorted(agg):
    a = agg[ln]
    if a[0] > tot * minp / 100 or a[1] > ts * minp / 100:
        top = sorted(a[3].items(), key=lambda kv: -kv[1])[:2]
        print("%5d inst %5.1f%% samp %5.1f%% lanes %4.1f %-38s| %s" % (ln, 100 * a[0] / tot, 100 * a[1] / max(ts, 1), a[2] / max(a[0], 1),
              " ".join("%s:%d" % (k[6:], v) for k, v in top if v), src[ln][:100]))
