"""Hot spots of one kernel from an ncu report: python tools/ncu_hot.py REPORT.ncu-rep KERNEL [TOP]
Prints the SASS instructions of the source page ordered by address with executed counts and stall
samples, keeping only those above 0.4 % of either total, plus per-opcode totals."""
import csv, subprocess, sys, collections
rep, kern = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kern], capture_output=True, text=True).stdout
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith('"Kernel Name"')), len(lines))
rows = list(csv.DictReader(lines[start:end]))
tot_i = sum(float(r["Instructions Executed"]) for r in rows)
tot_s = sum(float(r["# Samples"]) for r in rows)
print("instructions", tot_i, "samples", tot_s, "sass lines", len(rows))
ops = collections.Counter(); ops_s = collections.Counter()
for k, r in enumerate(rows):
    ie, sm = float(r["Instructions Executed"]), float(r["# Samples"])
    op = r["Source"].split()[1 if r["Source"].strip().startswith("@") else 0].split(".")[0]
    ops[op] += ie; ops_s[op] += sm
    if ie > 0.004 * tot_i or sm > 0.004 * tot_s:
        stalls = {c[6:]: int(r[c]) for c in r if c.startswith("stall_") and "Not Issued" not in c and r[c] not in ("", "0")}
        top = sorted(stalls.items(), key=lambda kv: -kv[1])[:3]
        print(f"{k:5d} {100*ie/tot_i:5.2f}%i {100*sm/tot_s:5.2f}%s thr {r['Avg. Threads Executed']:>5s} {r['Source'].strip()[:70]:70s} {top}")
print("per opcode (% instr, % samples):")
for op, v in ops.most_common(25):
    print(f"  {op:10s} {100*v/tot_i:5.1f} {100*ops_s[op]/tot_s:5.1f}")
