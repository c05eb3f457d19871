#!/usr/bin/env python
"""Launch list of the LAST step in an `ncu --metrics gpu__time_duration.sum --csv` log of bench.py: everything from
the last k_prep launch on, with each kernel's share of the step."""
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
last = max(i for i, r in enumerate(rows) if r[4].startswith('k_prep'))
step = rows[last:]
total = sum(float(r[14]) for r in step)
for r in step:
    name = r[4].split('(')[0]
    print('%-34s grid %-17s block %-14s %10.3f ms %6.1f%%' % (name, r[8], r[7], float(r[14]) / 1e6, 100 * float(r[14]) / total))
print('%-34s %54.3f ms' % ('total', total / 1e6))
