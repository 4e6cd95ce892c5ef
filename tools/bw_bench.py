#!/usr/bin/env python
"""Device time (CUDA events, median of N) of the bandwidth-bound kernels at the bench shape (32 generated frames of a 65-frame
2x clip at 512x512): flow warp, uint8 -> network input, rasteriser, composite.   python tools/bw_bench.py [--iters N]"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'render-in-between_b200'))


def timed(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--iters', type=int, default=20)
    a = ap.parse_args()
    import rib
    from rib import ops
    from rib.synth import synth_flow, synth_image, synth_joints
    dev = torch.device('cuda:0')
    b, h, w = 32, 512, 512
    src = synth_image(b, h, w, seed=1).to(dev)
    flow = synth_flow(b, h, w, seed=2).to(dev)
    flow16 = flow.half()
    out = torch.empty_like(src)
    joints = torch.from_numpy(synth_joints(65, h, w, seed=0)[1::2]).to(dev)
    keys = torch.randint(0, 256, (33, h, w, 3), dtype=torch.uint8, device=dev)
    kf = torch.empty(33, 3, h, w, dtype=torch.float32, device=dev)
    ku = torch.empty(33, h, w, 3, dtype=torch.uint8, device=dev)
    mask = torch.rand(b, 1, h, w, device=dev)
    u8 = torch.empty(b, h, w, 3, dtype=torch.uint8, device=dev)
    scratch = torch.empty(64 << 20, dtype=torch.uint8, device=dev)   # touched between runs: nothing stays in L2 by luck
    res = {}
    res['warp f32 flow'] = timed(lambda: ops.warp(src, flow, out=out), a.iters)
    res['warp f16 flow'] = timed(lambda: ops.warp(src, flow16, out=out), a.iters)
    res['frames_from_u8 (33 key frames)'] = timed(lambda: ops.frames_from_u8(keys, out=kf, out_u8=ku), a.iters)
    res['rasterize (planar + f32 label)'] = timed(lambda: ops.rasterize(joints, h, w), a.iters)
    res['composite (+u8)'] = timed(lambda: ops.composite(src, mask, out, want_u8=True, out=kf[:b], out_u8=u8), a.iters)
    for k, v in res.items():
        print('%-36s %8.1f us' % (k, v))


if __name__ == '__main__':
    main()
