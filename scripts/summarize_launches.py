"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel + grid."""
import collections
import csv
import re
import sys

path = sys.argv[1]
with open(path) as f:
    lines = [l for l in f if not l.startswith('==')]
rows = list(csv.DictReader(lines))
tot = collections.defaultdict(lambda: [0, 0.0])
for row in rows:
    k = row['Kernel Name']
    v = float(row['Metric Value'])
    m = re.search(r'tapgemm_sm100_kernel<(\d+), (\d+), (\d+)(?:, (\d+))?>', k)
    if m:
        name = f"tapgemm<bn{m.group(1)},cs{m.group(4) or 1}> grid={row['Grid Size']}"
    else:
        name = re.sub(r'\(.*', '', k).replace('its::', '').replace('void ', '')[:44]
    tot[name][0] += 1
    tot[name][1] += v
for k, (n, v) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:52s} n={n:4d} total={v/1e3:9.1f} us  avg={v/n/1e3:7.1f}")
print("sum us", round(sum(float(r['Metric Value']) for r in rows) / 1e3, 1), "launches", len(rows))
