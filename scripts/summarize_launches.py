"""Summarise an `ncu --metrics ... --csv` launch list: one row per launch (joined with the op list of
scripts/profile_pass.py --list when given) and totals per kernel.
Usage: python scripts/summarize_launches.py launches.csv [kinds.txt [traffic.json]]
With a third argument the per-family DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum) of one UNet
pass is written as JSON: bench.py reports it as `roofline.traffic`."""
import collections
import csv
import re
import sys

path = sys.argv[1]
with open(path) as f:
    lines = [l for l in f if not l.startswith('==')]
rows = list(csv.DictReader(lines))
launch = collections.OrderedDict()
for r in rows:
    d = launch.setdefault(r['ID'], {'name': r['Kernel Name'], 'grid': r['Grid Size']})
    try:
        d[r['Metric Name']] = float(r['Metric Value'].replace(',', ''))
    except ValueError:
        pass
    d['unit:' + r['Metric Name']] = r['Metric Unit']
kinds = []
if len(sys.argv) > 2:
    for l in open(sys.argv[2]).read().splitlines()[1:]:
        m = re.match(r'(\d+) (\w+) (\d+) (\d+) ?(.*)', l)
        if m:
            kinds += [(m.group(2), int(m.group(4)), m.group(5))] * int(m.group(3))
tot = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])


def short(k):
    m = re.search(r'tapgemm_persist_kernel<(\d+), (\d+), (\d+), (\d+)', k)
    if m:
        return f"tapgemm_persist<bn{m.group(1)},st{m.group(2)},mt{m.group(3)},ks{m.group(4)}>"
    m = re.search(r'tapgemm_sm100_kernel<(\d+)', k)
    if m:
        return f"tapgemm_sm100<bn{m.group(1)}>"
    return re.sub(r'\(.*', '', k).replace('its::', '').replace('void ', '')[:40]


print("idx kernel grid time_us tensor_pipe_%elapsed dram_MB(read+write) flops_TF/s op")
for i, d in enumerate(launch.values()):
    t = d.get('gpu__time_duration.sum', 0.0)
    t_us = t / 1e3 if d.get('unit:gpu__time_duration.sum', 'ns').startswith('n') else t
    tp = d.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 0.0)
    scale = {'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1.0, 'Gbyte': 1e3}
    dr = sum(d.get(k, 0.0) * scale.get(d.get('unit:' + k, 'byte'), 1e-6) for k in ('dram__bytes_read.sum', 'dram__bytes_write.sum'))
    name = short(d['name'])
    op = kinds[i] if i < len(kinds) else ('', 0, '')
    tf = op[1] / t_us / 1e6 if t_us and op[1] else 0.0
    print(f"{i:3d} {name:44s} {d['grid']:14s} {t_us:7.1f} {tp:6.1f} {dr:8.2f} {tf:7.0f}  {op[0]} {op[2]}")
    a = tot[name]
    a[0] += 1; a[1] += t_us; a[2] += tp * t_us; a[3] += dr
print()
print("kernel                                         n   total_us  avg_us  time-weighted tensor_pipe%  dram_MB")
for k, (n, t, tp, dr) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:44s} {n:4d} {t:9.1f} {t/n:7.1f} {tp/t if t else 0:10.1f} {dr:10.1f}")
print("sum us", round(sum(v[1] for v in tot.values()), 1), "launches", len(launch))

if len(sys.argv) > 3:
    import json
    fam = collections.defaultdict(lambda: {"launches": 0, "time_us_under_ncu": 0.0, "dram_bytes": 0.0})
    for k, (n, t, tp, dr) in tot.items():
        f = "tapgemm" if k.startswith("tapgemm") else k
        fam[f]["launches"] += n
        fam[f]["time_us_under_ncu"] += t
        fam[f]["dram_bytes"] += dr * 1e6
    json.dump({"source": path, "note": "one UNet pass; ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,"
               "gpu__time_duration.sum --clock-control none (cold-cache, serialised launches)", "families": fam},
              open(sys.argv[3], "w"), indent=1)
