"""Pick one op per distinct kernel name (the longest launch) from an ncu launch list of a whole pass.

    python scripts/ncu_pick.py launches.csv oplist.txt  ->  prints "i,j,k" (op indices for ncu_capture.py --ops)
"""
import csv
import sys

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") == "gpu__time_duration.sum":
        rows.append((r["Kernel Name"], float(r["Metric Value"].replace(",", ""))))
ops = []
for l in open(sys.argv[2]):
    p = l.split()
    if len(p) >= 3 and p[0].isdigit():
        ops += [int(p[0])] * int(p[2])          # an op with two launches owns two consecutive rows
if len(ops) != len(rows):
    raise SystemExit(f"{len(rows)} profiled launches but the plan lists {len(ops)}")
best = {}
for (name, dur), op in zip(rows, ops):
    if name not in best or dur > best[name][0]:
        best[name] = (dur, op)
print(",".join(str(op) for _, op in sorted(best.values(), key=lambda t: t[1])))
