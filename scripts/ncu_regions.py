#!/usr/bin/env python
"""Executed warp instructions / active threads / stall samples per block of SASS of one
profiled kernel (ncu --page source --csv): where the instructions of a kernel go."""
import subprocess, csv, sys
rep = sys.argv[1]
step = int(sys.argv[2]) if len(sys.argv) > 2 else 120
raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
# the report may hold several kernels: "Kernel Name" row, header row, instruction rows; argv[3] picks one by substring
want = sys.argv[3] if len(sys.argv) > 3 else ''
blocks, cur = [], None
for r in rows:
    if r and r[0] == 'Kernel Name':
        cur = {'name': r[1], 'hdr': None, 'rows': []}
        blocks.append(cur)
    elif cur is not None and cur['hdr'] is None:
        cur['hdr'] = r
    elif cur is not None:
        cur['rows'].append(r)
blk = [b for b in blocks if want in b['name']][-1]
print(blk['name'])
hdr = blk['hdr']
ia, ie, it, iss = (hdr.index(k) for k in ('Source', 'Instructions Executed', 'Thread Instructions Executed',
                                           'Warp Stall Sampling (All Samples)'))
data = [(int(r[ie]), int(r[it]), int(r[iss]), r[ia]) for r in blk['rows'] if len(r) > it]
tot = sum(d[0] for d in data)
first = data[0][0]
print('total warp instructions', tot, '| SASS', len(data), '| per CTA (first instruction count = warps launched):', tot / max(1, first) * 1)
for i in range(0, len(data), step):
    seg = data[i:i + step]
    n = sum(d[0] for d in seg)
    print('%5d %6.1f%%  thr/inst %5.1f  stalls %7d  %s' % (i, 100 * n / tot, sum(d[1] for d in seg) / max(1, n),
                                                        sum(d[2] for d in seg), seg[0][3].strip()[:70]))
