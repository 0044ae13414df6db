#!/usr/bin/env python
"""Print the per-launch table of an `ncu --metrics gpu__time_duration.sum --csv` log."""
import csv, sys
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
rows = []
for row in csv.DictReader(lines):
    if row.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    v = float(row['Metric Value'].replace(',', ''))
    u = row['Metric Unit']
    v = v / 1e3 if u == 'ns' else (v * 1e3 if u == 'ms' else v)
    rows.append((row['Kernel Name'][:56], row.get('Grid Size'), row.get('Block Size'), v))
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
for r in rows[skip:]:
    print(f'{r[0]:58s} {r[1]:>14s} {r[2]:>12s} {r[3]:10.1f} us')
print('total us', sum(r[3] for r in rows[skip:]))
