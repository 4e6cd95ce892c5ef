#!/usr/bin/env python
"""Joins gpurun_out/plan_B*_*.txt (rib_generator_plan_text) with an ncu raw-page CSV of the conv_gemm launches
(tools/ncu_conv.sh) and prints one line per layer: measured time, tensor-pipe %, DRAM / L2 traffic and the
per-layer roofline time max(FLOPs / TC peak, algorithmic bytes / HBM peak).

    python tools/conv_report.py gpurun_out/plan_B32_512.txt gpurun_out/conv_v3.csv
"""
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.isfile(p):
        d = json.load(open(p))
        return d['bf16_tflops'] * 1e12, d['hbm_gbs'] * 1e9
    return 1.59e15, 6.65e12


def parse_plan(path):
    out = []
    for line in open(path):
        if not line.startswith('gemm '):
            continue
        parts = line.split()
        d = {'name': parts[1]}
        for kv in parts[2:]:
            k, v = kv.split('=')
            d[k] = float(v) if k == 'flops' else int(v)
        out.append(d)
    return out


def alg_bytes(d):
    """Algorithmic HBM bytes: inputs read once + outputs written once (16-bit), weights once."""
    px = d['B'] * d['H'] * d['W']
    s = d['stride']
    rd = px * s * s * d['cin0'] * 2 + px * d['cin1'] * 2
    if d['mode'] in (1, 3):      # SPADE: reads x (C = N/2 per quantity... N = nq*2*C) and writes nq maps of C
        c_total = d['N'] // 2
        rd += px * c_total * 2 * 0 + px * (d['N'] // 2) * 2 * 0   # x is re-read per quantity; counted below
        wr = px * (d['N'] // 2) * 2
        rd += px * (d['N'] // 2) * 2                                # upper bound (x read once per quantity)
    elif d['mode'] == 2:
        wr = px * d['nvalid'] * 4
    else:
        # plain store: the real output columns (a sub-pixel conv's N = 4 x Cout are all real; merged convs pad N)
        wr = px * (d['N'] if d['taps'] == 4 else d['nvalid']) * 2
    return rd + wr + d['N'] * d['K'] * 2


def main(plan_path, csv_path):
    tc, hbm = peaks()
    plan = parse_plan(plan_path)
    rows = list(csv.reader(open(csv_path)))
    hdr = rows[0]
    ix = {n: i for i, n in enumerate(hdr)}
    data = [r for r in rows[2:] if len(r) == len(hdr) and 'conv_gemm' in r[ix['Kernel Name']]]

    units = rows[1]
    scale = {'ms': 1e3, 'us': 1.0, 'ns': 1e-3, 's': 1e6, 'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0, 'Tbyte': 1e12}

    def col(r, name, default=0.0):
        if name not in ix:
            return default
        try:
            return float(r[ix[name]].replace(',', '')) * scale.get(units[ix[name]], 1.0)
        except ValueError:
            return default
    n = min(len(plan), len(data))
    print('%-22s %3s %4s %5s %5s %3s %2s | %7s %7s %5s | %5s %5s %5s | %6s %6s %6s' % (
        'layer', 'md', 'HxW', 'N', 'K', 'BN', 'MT', 'us', 'ideal', 'eff', 'TC%', 'DRAM%', 'L2%', 'dramMB', 'algMB', 'L2MB'))
    tot = tot_ideal = 0.0
    for d, r in zip(plan[:n], data[:n]):
        us = col(r, 'gpu__time_duration.sum')
        tcp = col(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', -1)
        dram = col(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')
        l2 = col(r, 'lts__throughput.avg.pct_of_peak_sustained_elapsed')
        dmb = (col(r, 'dram__bytes_read.sum') + col(r, 'dram__bytes_write.sum')) / 1e6
        l2mb = col(r, 'lts__t_bytes.sum') / 1e6
        ab = alg_bytes(d)
        ideal = max(d['flops'] / tc, ab / hbm) * 1e6
        tot += us
        tot_ideal += ideal
        print('%-22s %3d %4d %5d %5d %3d %2d | %7.1f %7.1f %5.2f | %5.1f %5.1f %5.1f | %6.0f %6.0f %6.0f' % (
            d['name'][:22], d['mode'], d['H'], d['N'], d['K'], d['BN'], d['MT'], us, ideal, ideal / us, tcp, dram, l2,
            dmb, ab / 1e6, l2mb))
    print('total %.1f us, per-layer roofline %.1f us (%.2f)' % (tot, tot_ideal, tot_ideal / tot))


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2])
