"""Condense an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv`) into one line per
kernel: launches, total and mean device time, share of the listed time.
    python tools/launch_summary.py profiles/r2_launches_raw.csv > profiles/r2_launch_list_summary.csv"""
import collections
import csv
import re
import sys

rows = [l for l in open(sys.argv[1]) if l.startswith('"')]
stats = collections.OrderedDict()
for rec in csv.DictReader(rows):
    name = re.sub(r'^void ', '', rec['Kernel Name'])
    name = re.sub(r'\(anonymous namespace\)::|at::', '', name)
    name = re.sub(r'\(.*$', '', name)                       # drop the parameter list
    us = float(rec['Metric Value']) / 1e3
    n, total = stats.get(name, (0, 0.0))
    stats[name] = (n + 1, total + us)
listed = sum(t for _, t in stats.values())
out = csv.writer(sys.stdout, quoting=csv.QUOTE_MINIMAL)
out.writerow(['kernel', 'launches', 'total_us', 'mean_us', 'share_of_listed_time'])
for name, (n, total) in sorted(stats.items(), key=lambda kv: -kv[1][1]):
    out.writerow([name, n, f'{total:.1f}', f'{total / n:.2f}', f'{total / listed:.4f}'])
