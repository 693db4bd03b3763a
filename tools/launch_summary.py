#!/usr/bin/env python
"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel name, launches and mean us."""
import collections
import csv
import sys


def summarise(fn, skip_torch=True):
    rows = list(csv.reader(open(fn)))
    hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
    hdr = rows[hi]
    ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
    agg = collections.OrderedDict()
    for r in rows[hi + 2:]:
        if len(r) <= vi:
            continue
        n = r[ki]
        if skip_torch and ('at::' in n or 'distribution' in n):
            continue
        n = n.split('(')[0][:70]
        agg.setdefault(n, []).append(float(r[vi].replace(',', '')))
    return agg


if __name__ == '__main__':
    for fn in sys.argv[1:]:
        print(fn)
        for n, v in summarise(fn).items():
            print('  %-70s n=%3d mean=%9.1f us  max=%9.1f' % (n, len(v), sum(v) / len(v) / 1e3, max(v) / 1e3))
