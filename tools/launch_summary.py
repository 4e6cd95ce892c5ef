#!/usr/bin/env python
"""Per-kernel summary of an ncu launch list of `bench.py --steps K --warmup W` (tools/launch_table.py pivot):
launches, time and DRAM bytes per step.  Writes profiles/<tag>_launch_summary.json and refreshes
profiles/conv_gemm_traffic.json, the file bench.py reads `roofline.traffic` from.

    python tools/launch_summary.py gpurun_out/launches_<tag>.csv <tag> <steps run under ncu>
"""
import csv
import json
import os
import re
import sys
from collections import OrderedDict, defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main(path, tag, steps):
    lines = [l for l in open(path) if l.startswith('"')]
    r = csv.reader(lines)
    hdr = next(r)
    ix = {n: i for i, n in enumerate(hdr)}
    rows = OrderedDict()
    for x in r:
        d = rows.setdefault(x[ix['ID']], {'name': x[ix['Kernel Name']]})
        d[x[ix['Metric Name']]] = float(x[ix['Metric Value']].replace(',', ''))
    # A step starts with the key-frame pass of ClipRenderer.render (a composite launch that follows a non-raster
    # kernel); only complete steps with the modal launch count are kept, which drops model creation and the
    # auto-tuner's candidate launches of the first step.
    allrows = list(rows.values())
    def short(d):
        m = re.search(r'rib::(\w+)', d['name'])
        return m.group(1) if m else d['name'][:40]
    starts = [i for i, d in enumerate(allrows) if short(d) == 'frames_from_u8_kernel']   # r2+: uint8 key frames first
    if not starts:
        starts = [i for i, d in enumerate(allrows) if short(d) == 'composite_kernel' and
                  (i + 1 < len(allrows) and short(allrows[i + 1]) != 'composite_kernel')]
    segs = [allrows[a:b] for a, b in zip(starts, starts[1:] + [len(allrows)])]
    from collections import Counter
    modal = Counter(len(x) for x in segs).most_common(1)[0][0]
    segs = [x for x in segs if len(x) == modal]
    steps = len(segs)
    rows = OrderedDict((i, d) for i, d in enumerate(d for seg in segs for d in seg))
    agg = defaultdict(lambda: {'launches': 0, 'us': 0.0, 'dram_read': 0.0, 'dram_write': 0.0})
    for d in rows.values():
        m = re.search(r'rib::(\w+)', d['name'])
        a = agg[m.group(1) if m else d['name'][:40]]
        a['launches'] += 1
        a['us'] += d.get('gpu__time_duration.sum', 0) / 1e3
        a['dram_read'] += d.get('dram__bytes_read.sum', 0)
        a['dram_write'] += d.get('dram__bytes_write.sum', 0)
    once = ('sn_sigma_inv_kernel', 'sn_sigma_rows_kernel', 'sn_sigma_final_kernel', 'pack_weight_kernel', 'pack_weight_subpix_kernel')   # model creation, not per step
    out = {'source': os.path.basename(path), 'clean_steps_averaged': steps, 'launches_per_step': modal, 'per_step': {}, 'once': {}}
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]['us']):
        if k in once:
            out['once'][k] = {'launches': a['launches'], 'us': round(a['us'], 1)}
        else:
            out['per_step'][k] = {'launches': a['launches'] / steps, 'us': round(a['us'] / steps, 1),
                                  'dram_read_bytes': round(a['dram_read'] / steps), 'dram_write_bytes': round(a['dram_write'] / steps)}
    tot = sum(v['us'] for v in out['per_step'].values())
    for v in out['per_step'].values():
        v['share_of_step'] = round(v['us'] / tot, 4)
    out['per_step_total_us'] = round(tot, 1)
    json.dump(out, open(os.path.join(ROOT, 'profiles', '%s_launch_summary.json' % tag), 'w'), indent=1)
    c = out['per_step'].get('conv_gemm_kernel')
    if c:
        json.dump({'source': 'profiles/%s_launch_summary.json (ncu dram__bytes_read.sum + dram__bytes_write.sum of every '
                             'conv_gemm launch of a bench step)' % tag,
                   'launches_per_step': c['launches'], 'dram_bytes_per_step': c['dram_read_bytes'] + c['dram_write_bytes'],
                   'dram_bytes_per_launch': (c['dram_read_bytes'] + c['dram_write_bytes']) / c['launches'],
                   'share_of_step_under_ncu': c['share_of_step']},
                  open(os.path.join(ROOT, 'profiles', 'conv_gemm_traffic.json'), 'w'), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]))
