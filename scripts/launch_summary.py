#!/usr/bin/env python
"""Launch lists in an `ncu --metrics gpu__time_duration.sum --csv` log of bench.py: the LAST pileup step (k_prep ...
k_tile<variation_strand> ... k_tile<coverage_ext>) and the LAST binned sweep (k_prep ... k_tile<coverage_gc>, k_gc,
k_binned_*), with each kernel's share of its step."""
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
starts = [i for i, r in enumerate(rows) if r[4].startswith('k_prep')] + [len(rows)]
steps = [rows[a:b] for a, b in zip(starts[:-1], starts[1:])]
for label, key in (('pileup step (variation_strand + coverage_ext)', 'k_tile<2, 5'), ('binned sweep (coverage_gc + coverage / mapq / tlen binned)', 'k_binned_acc')):
    sel = [s for s in steps if any(key in r[4] for r in s)]
    if not sel:
        continue
    step = sel[-1]
    total = sum(float(r[14]) for r in step)
    print('== ' + label)
    for r in step:
        name = r[4].split('(')[0]
        print('%-34s grid %-17s block %-14s %10.3f ms %6.1f%%' % (name, r[8], r[7], float(r[14]) / 1e6, 100 * float(r[14]) / total))
    print('%-34s %54.3f ms' % ('total', total / 1e6))
