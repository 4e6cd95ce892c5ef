#!/usr/bin/env python
"""Where the warps of a conv_gemm launch spend their time, by role: an `ncu --page source --csv` export is split into
code regions at the role boundaries of the kernel (TMA producer, MMA issuer, transform warps, epilogue) and the stall
samples and executed instructions are summed per region.  The regions are found from landmark instructions: the first
UTMALDG (producer), the first UTCHMMA (MMA issuer), the first LDTM (epilogue).

    python tools/src_roles.py file.source.csv
"""
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    ix = {n: i for i, n in enumerate(hdr)}
    body = rows[2:]
    stall_cols = [n for n in hdr if n.startswith('stall_') and 'Not Issued' not in n]
    src = [r[ix['Source']].strip() for r in body]

    def first(pat, start=0):
        for i in range(start, len(src)):
            if pat in src[i]:
                return i
        return None
    marks = [('prologue', 0)]
    for name, pat in (('tma_producer', 'UTMALDG'), ('mma_issuer', 'UTCHMMA'), ('epilogue', 'LDTM')):
        i = first(pat)
        if i is not None:
            marks.append((name, i))
    marks.sort(key=lambda m: m[1])
    bounds = [(n, a, (marks[k + 1][1] if k + 1 < len(marks) else len(body))) for k, (n, a) in enumerate(marks)]
    tot = sum(int(r[ix['# Samples']] or 0) for r in body)
    print('total samples %d, %d SASS instructions' % (tot, len(body)))
    for name, a, b in bounds:
        seg = body[a:b]
        n = sum(int(r[ix['# Samples']] or 0) for r in seg)
        ex = sum(int(r[ix['Instructions Executed']] or 0) for r in seg)
        st = sorted(((sum(int(r[ix[c]] or 0) for r in seg), c[6:]) for c in stall_cols), reverse=True)[:6]
        print('%-13s sass %5d-%5d  samples %6d (%4.1f%%)  warp-instr executed %10d   %s' % (
            name, a, b, n, 100.0 * n / max(tot, 1), ex, ' '.join('%s:%d' % (c, v) for v, c in st if v)))
    # instruction mix of the epilogue by opcode (executed warp-instructions)
    for name, a, b in bounds:
        if name != 'epilogue':
            continue
        mix = {}
        for r in body[a:b]:
            op = r[ix['Source']].strip().split()
            op = op[1] if op and op[0].startswith('@') and len(op) > 1 else (op[0] if op else '?')
            op = op.split('.')[0]
            mix[op] = mix.get(op, 0) + int(r[ix['Instructions Executed']] or 0)
        tot_ex = sum(mix.values())
        print('epilogue mix:', ' '.join('%s:%.1f%%' % (k, 100.0 * v / max(tot_ex, 1)) for k, v in sorted(mix.items(), key=lambda kv: -kv[1])[:16]))


if __name__ == '__main__':
    main(sys.argv[1])
