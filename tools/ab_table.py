#!/usr/bin/env python
"""Per-layer table of a tools/gpu_ab_libs.sh run: mean of the two repetitions per build, layers that moved by > 3 %.
    python tools/ab_table.py <tag> <variant> [<variant> ...]"""
import sys


def load(f):
    d = []
    for l in open(f):
        a = l.split()
        if len(a) == 2:
            d.append((a[0], float(a[1])))
    return d


def main(tag, variants):
    data = {}
    for v in ['base'] + variants:
        a, b = load('gpurun_out/conv_events_%s_%s_1.txt' % (tag, v)), load('gpurun_out/conv_events_%s_%s_2.txt' % (tag, v))
        data[v] = [(k, (x + b[i][1]) / 2) for i, (k, x) in enumerate(a)]
    for i, (k, x) in enumerate(data['base']):
        ys = [data[v][i][1] for v in variants]
        if any(abs(y - x) > 0.03 * x for y in ys) or k in ('total', 'forward_wall'):
            print('%-20s base %8.1f  ' % (k, x) + '  '.join('%s %8.1f (%+6.1f)' % (v, y, y - x) for v, y in zip(variants, ys)))


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2:])
