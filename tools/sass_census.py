#!/usr/bin/env python
"""tcgen05 / TMA / TMEM instruction census of the shipped library, per kernel (cuobjdump -sass, runs anywhere).
    python tools/sass_census.py > profiles/<tag>_sass_tcgen05_tma.txt"""
import os
import re
import subprocess
import sys
from collections import Counter, OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'render-in-between_b200', 'rib', 'librib_b200.so')
PAT = re.compile(r'\b(UTCHMMA(?:\.2CTA)?|UTMALDG\.\dD(?:\.2CTA)?|UTMASTG[.\w]*|UTCBAR(?:\.2CTA\.MULTICAST)?|UTCATOMSWS(?:\.2CTA)?|LDTM|HMMA|IMMA|'
                 r'UCGABAR_ARV|LDGSTS|SYNCS)\b')


def main():
    out = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True, check=True).stdout
    per = OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r'\s*Function : (\S+)', line)
        if m:
            name = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = per.setdefault(re.sub(r'\(.*', '', name), Counter())
            continue
        if cur is None:
            continue
        for k in PAT.findall(line):
            cur[k] += 1
    print('# cuobjdump -sass render-in-between_b200/rib/librib_b200.so: tcgen05 / TMA / TMEM instructions per kernel (sm_100a).')
    print('# UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), UTCBAR = tcgen05.commit (.2CTA.MULTICAST = multicast::cluster),')
    print('# UTMALDG = cp.async.bulk.tensor (TMA load; .2CTA = cta_group::2 form), LDTM = tcgen05.ld, SYNCS = mbarrier ops,')
    print('# UTCATOMSWS = tcgen05.alloc/dealloc, UCGABAR_ARV = barrier.cluster, LDGSTS = cp.async.  HMMA / IMMA = mma.sync.')
    tot = Counter()
    for name, c in per.items():
        if not c:
            continue
        tot.update(c)
        print('%-70s %s' % (name[:70], ' '.join('%s=%d' % kv for kv in sorted(c.items()))))
    print('library totals: ' + ' '.join('%s=%d' % kv for kv in sorted(tot.items())))
    print('mma.sync instructions (HMMA + IMMA): %d' % (tot.get('HMMA', 0) + tot.get('IMMA', 0)))


if __name__ == '__main__':
    sys.exit(main())
