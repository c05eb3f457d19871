#!/usr/bin/env python
"""Attribute ncu warp-stall samples / executed instructions to CUDA source lines.

usage: ncu_lines.py <report.ncu-rep> <lib.so> <mangled-kernel-substring> [source-file-substring]
Joins `ncu --page source --csv` (per SASS instruction, in address order) with
`nvdisasm -g` of the cubin inside the shared library (same instruction order)."""
import csv, glob, os, re, subprocess, sys, tempfile
rep, lib, kern = sys.argv[1:4]
srcsub = sys.argv[4] if len(sys.argv) > 4 else '.cu'
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(lib)], cwd=tmp, capture_output=True)
lines = []          # (line or None) per instruction of the kernel
for cub in glob.glob(tmp + '/*.cubin'):
    txt = subprocess.run(['nvdisasm', '-g', '-c', cub], capture_output=True, text=True).stdout
    cur, infn = None, False
    for ln in txt.splitlines():
        if ln.startswith('.text.'):
            infn = kern in ln
            cur = None
            continue
        if not infn:
            continue
        m = re.match(r'\s*//## File "(.*)", line (\d+)', ln)
        if m:
            cur = (m.group(1), int(m.group(2)))
            continue
        if re.match(r'\s*/\*[0-9a-f]{4,}\*/', ln):
            lines.append((cur, ln.split('*/', 1)[1].strip()))
    if lines:
        break
raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
# the page holds one section per profiled launch: "Kernel Name" row, header row, SASS rows
secs = [k for k, r in enumerate(rows) if r and r[0] == 'Kernel Name']
want = int(os.environ.get('NCU_LAUNCH', '0'))
lo = secs[want]
hi = secs[want + 1] if want + 1 < len(secs) else len(rows)
print('launch %d of %d: %s' % (want, len(secs), rows[lo][1]))
hdr = rows[lo + 1]
si, ii = hdr.index('# Samples'), hdr.index('Instructions Executed')
body = [r for r in rows[lo + 2:hi] if len(r) > max(si, ii)]
print('sass in report %d, in cubin %d' % (len(body), len(lines)))
agg = {}
tot = toti = 0
for k, r in enumerate(body):
    s, i = int(r[si]), int(r[ii])
    tot += s; toti += i
    loc = lines[k][0] if k < len(lines) else None
    key = loc if (loc and srcsub in loc[0]) else ('(other) ' + (os.path.basename(loc[0]) if loc else '?'), 0)
    a = agg.setdefault(key, [0, 0, 0])
    a[0] += s; a[1] += i; a[2] += 1
src = {}
print('%-8s %7s %7s %5s  %s' % ('line', 'stall%', 'instr%', 'nsass', 'source'))
for key in sorted(agg, key=lambda k: (str(k[0]), k[1])):
    s, i, n = agg[key]
    if s < tot * 0.004 and i < toti * 0.004:
        continue
    f, l = key
    text = ''
    if l:
        if f not in src:
            try: src[f] = open(f).read().splitlines()
            except OSError: src[f] = []
        text = src[f][l - 1].strip()[:100] if l - 1 < len(src[f]) else ''
    print('%-8s %6.2f%% %6.2f%% %5d  %s' % (l or os.path.basename(str(f))[:8], 100.0 * s / tot, 100.0 * i / toti, n, text or f))
