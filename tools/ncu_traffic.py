#!/usr/bin/env python
"""profiles/ncu_traffic.json from an `ncu --set full` capture: per kernel, dram__bytes_read.sum + dram__bytes_write.sum
and the duration of its first captured launch.  bench.py reads the file for roofline.traffic.

usage: python tools/ncu_traffic.py REPORT.ncu-rep PARTICLES SOURCE_LABEL [kernel-name-prefix ...]"""
import csv
import io
import json
import os
import subprocess
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main():
    rep, particles, label = sys.argv[1], int(sys.argv[2]), sys.argv[3]
    want = sys.argv[4:] or ["k_detect", "k_keys", "k_scatter_advect", "k_pairs"]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units = rows[0], rows[1]
    col = {n: i for i, n in enumerate(head)}
    out, seen = [], set()
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        short = name.replace("void ", "").split("(")[0].split("<")[0]
        if short in seen or not any(short.startswith(w) for w in want):
            continue
        seen.add(short)
        b = sum(float(r[col[m]]) * UNIT[units[col[m]]] for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        out.append({"kernel": short, "particles": particles, "dram_bytes": b, "duration_us": float(r[col["gpu__time_duration.sum"]]),
                    "source": label})
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_traffic.json")
    with open(path, "w") as f:
        json.dump({"kernels": out}, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
