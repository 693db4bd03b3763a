#!/usr/bin/env python
"""Summarise an ncu report per CUDA source line: samples, instructions, top stall reasons.
usage: tools/ncu_lines.py <report.ncu-rep> <kernel-regex> [top_n]"""
import csv
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass', '--kernel-name',
                      'regex:' + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fname, hdr, data = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path':
        fname = r[1].split('/')[-1]
        continue
    if r[0] == 'Line No':
        hdr = r
        continue
    if hdr is None or r[0] in ('', 'Function Name'):
        continue
    try:
        ln = int(r[0])
    except ValueError:
        continue
    d = dict(zip(hdr[2:], r[2:]))
    try:
        s = int(d['# Samples'])
        ie = int(d['Instructions Executed'])
    except (KeyError, ValueError):
        continue
    stalls = sorted(((int(v), k) for k, v in d.items()
                     if k.startswith('stall_') and 'Not Issued' not in k and v.isdigit() and int(v) > 0), reverse=True)
    data.append((s, ie, fname, ln, r[1].strip()[:90], stalls[:3]))
tot_s = sum(d[0] for d in data) or 1
tot_i = sum(d[1] for d in data) or 1
print('total samples {}  total warp-instructions {}'.format(tot_s, tot_i))
for s, ie, f, ln, src, st in sorted(data, reverse=True)[:top]:
    print('{:5.1f}% smp {:5.1f}% ins  {}:{}  {}   {}'.format(100.0 * s / tot_s, 100.0 * ie / tot_i, f, ln, src,
                                                          ' '.join('{}={}'.format(k[6:], v) for v, k in st)))
