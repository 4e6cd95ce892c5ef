#!/usr/bin/env python
"""Top stall sites of an `ncu --page source --csv` export: prints the N most-sampled SASS instructions with
their dominant stall reasons and a little context.   python tools/src_top.py file.source.csv [N] [context]"""
import csv
import sys


def main(path, top=25, ctx=2):
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    ix = {n: i for i, n in enumerate(hdr)}
    body = rows[2:]
    stall_cols = [n for n in hdr if n.startswith('stall_') and 'Not Issued' not in n]
    tot = sum(int(r[ix['# Samples']] or 0) for r in body)
    order = sorted(range(len(body)), key=lambda i: -int(body[i][ix['# Samples']] or 0))[:top]
    print('total samples', tot, 'instructions', len(body))
    for i in sorted(order):
        r = body[i]
        n = int(r[ix['# Samples']] or 0)
        st = sorted(((int(r[ix[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:3]
        for k in range(max(0, i - ctx), i):
            print('        %5d  %s' % (k, body[k][ix['Source']].strip()[:90]))
        print('%5.1f%%  %5d  %-70s exec=%s  %s' % (100.0 * n / max(tot, 1), i, r[ix['Source']].strip()[:70],
                                                   r[ix['Instructions Executed']], ' '.join('%s:%d' % (c, v) for v, c in st if v)))
    return 0


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25, int(sys.argv[3]) if len(sys.argv) > 3 else 2)
