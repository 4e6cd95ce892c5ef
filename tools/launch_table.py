#!/usr/bin/env python
"""Pivot an ncu --csv launch log (one row per launch x metric) into one line per launch."""
import csv
import re
import sys
from collections import OrderedDict


def main(path):
    lines = [l for l in open(path) if l.startswith('"')]
    r = csv.reader(lines)
    hdr = next(r)
    ix = {n: i for i, n in enumerate(hdr)}
    rows = OrderedDict()
    for x in r:
        k = x[ix['ID']]
        d = rows.setdefault(k, {'name': x[ix['Kernel Name']], 'grid': x[ix['Grid Size']]})
        d[x[ix['Metric Name']]] = float(x[ix['Metric Value']].replace(',', ''))
    print('%4s %-28s %-16s %9s %9s %9s %9s %6s' % ('id', 'kernel', 'grid', 'us', 'dramR MB', 'dramW MB', 'L2 MB', 'regs'))
    tot = 0.0
    for k, d in rows.items():
        n = d['name']
        m = re.search(r'rib::(\w+)(<\d>)?', n)
        short = (m.group(1) + (m.group(2) or '')) if m else n[:28]
        us = d.get('gpu__time_duration.sum', 0) / 1e3
        tot += us
        print('%4s %-28s %-16s %9.1f %9.1f %9.1f %9.1f %6d' % (
            k, short, d['grid'], us, d.get('dram__bytes_read.sum', 0) / 1e6, d.get('dram__bytes_write.sum', 0) / 1e6,
            d.get('lts__t_bytes.sum', 0) / 1e6, d.get('launch__registers_per_thread', 0)))
    print('total %.1f us' % tot)


if __name__ == '__main__':
    main(sys.argv[1])
