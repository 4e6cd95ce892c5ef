#!/usr/bin/env python
"""Device time of one forward of the motion Transformer (csrc/motion.cu) at the configuration's longest clip (321 frames),
beside the same dataflow run by ATen on the same GPU (the oracle's functional restatement on CUDA tensors) and by the CPU
oracle on the host cores.   python tools/motion_bench.py [--iters N] [--batch B]"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'render-in-between_b200'))


def gpu_ms(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--iters', type=int, default=50)
    ap.add_argument('--length', type=int, default=321)
    ap.add_argument('--rate', type=int, default=16)
    a = ap.parse_args()
    from oracle import motion_oracle as mo
    from rib.motion import MotionTransformer
    dev = torch.device('cuda:0')
    sd = mo.synth_state_dict(1)
    m = MotionTransformer(mo.CFG['input_joints'])
    m.load_state_dict(sd, strict=True)
    m = m.to(dev).eval()
    sd_gpu = {k: v.to(dev) for k, v in sd.items()}
    out = {'length': a.length, 'rate': a.rate, 'iters': a.iters}
    c = mo.CFG
    e, ff, j, l = c['hidden_dim'], c['dim_feedforward'], c['input_joints'], a.length
    enc = 2 * l * e * 3 * e + 4 * l * l * e + 2 * l * e * e + 4 * l * e * ff
    dec = enc + 2 * l * e * 3 * e + 4 * l * l * e + 2 * l * e * e
    out['flop_per_sequence'] = c['enc_layers'] * enc + c['dec_layers'] * dec + 3 * 2 * l * e * j
    for batch in (1, 8, 64):
        seqs = [mo.synth_motion(a.length, a.rate, seed=s) for s in range(batch)]
        src = torch.stack([s[0] for s in seqs]).to(dev)
        sm = torch.stack([s[1] for s in seqs]).to(dev)
        tm = torch.stack([s[2] for s in seqs]).to(dev)
        pos = mo.position_encoding(batch, a.length).to(dev)
        ours = gpu_ms(lambda: m(src, sm, pos, None, tm, pos, a.rate), a.iters)
        g = torch.cuda.CUDAGraph()
        m(src, sm, pos, None, tm, pos, a.rate)
        torch.cuda.synchronize()
        with torch.cuda.graph(g):
            m(src, sm, pos, None, tm, pos, a.rate)
        graph = gpu_ms(g.replay, a.iters)

        def aten():
            # the oracle builds its masks on the host; move the inputs once and patch torch.zeros / eye through a device context
            with torch.device(dev):
                return mo.transformer_forward(sd_gpu, src, sm, pos, tm, pos, a.rate)
        with torch.no_grad():
            at = gpu_ms(aten, max(3, a.iters // 5))
        out['batch_%d' % batch] = {'ours_ms': ours, 'ours_cuda_graph_ms': graph, 'aten_gpu_ms': at,
                                   'sequences_per_s': batch / (graph * 1e-3),
                                   'gflops_in_graph': out['flop_per_sequence'] * batch / (graph * 1e-3) / 1e9}
    data, em, dm = mo.synth_motion(a.length, a.rate, seed=0)
    pos = mo.position_encoding(1, a.length)
    with torch.no_grad():
        mo.transformer_forward(sd, data[None], em[None], pos, dm[None], pos, a.rate)
        t0 = time.perf_counter()
        for _ in range(5):
            mo.transformer_forward(sd, data[None], em[None], pos, dm[None], pos, a.rate)
        out['cpu_oracle_ms'] = (time.perf_counter() - t0) / 5 * 1e3
    out['cpu_threads'] = torch.get_num_threads()
    print(json.dumps(out, indent=1))


if __name__ == '__main__':
    main()
