#!/usr/bin/env python
"""Per-layer CUDA-event timing of the implicit-GEMM launches of one generator forward (B=32, 512x512 by default),
joined with the plan text.  RIB_LIB=<path to another build> python tools/conv_bench.py compares kernel variants.
Events are recorded on the launching stream around every launch (rib_profile_enable); no profiler involved."""
import argparse
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'render-in-between_b200')):
    if p not in sys.path:
        sys.path.insert(0, p)
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=32)
    ap.add_argument('--size', type=int, default=512)
    ap.add_argument('--iters', type=int, default=5)
    ap.add_argument('--out', default='')
    a = ap.parse_args()
    from rib._lib import lib
    from rib.arch import Arch
    from rib.config import default_gen_cfg
    from rib.generator import Generator
    from rib.synth import synth_image, synth_joints, synth_state_dict
    import rib
    dev = torch.device('cuda:0')
    cfg = default_gen_cfg()
    gen = Generator(cfg)
    gen.load_state_dict(synth_state_dict(Arch(cfg), seed=0, power_iters=5), strict=True)
    gen = gen.to(dev).eval()
    b, h, w = a.batch, a.size, a.size
    label = rib.rasterize(torch.from_numpy(synth_joints(b, h, w, seed=3)).to(dev), h, w)
    fake, prev = synth_image(b, h, w, seed=1).to(dev), synth_image(b, h, w, seed=2).to(dev)
    with torch.no_grad():
        for _ in range(2):
            gen(label, None, fake, prev)
        torch.cuda.synchronize()
        names = [l.split()[1] for l in gen.plan_text().splitlines() if l.startswith('gemm ')]
        acc = [0.0] * len(names)
        for _ in range(a.iters):
            lib.rib_profile_enable(1)
            gen(label, None, fake, prev)
            torch.cuda.synchronize()
            ms, n = C.c_double(), C.c_longlong()
            buf = (C.c_float * 256)()
            lib.rib_profile_collect_launches(C.byref(ms), C.byref(n), buf, 256)
            lib.rib_profile_enable(0)
            assert n.value == len(names), (n.value, len(names))
            for i in range(len(names)):
                acc[i] += buf[i] * 1e3 / a.iters
        # the whole forward as the caller sees it (launch gaps and the non-conv kernels included): CUDA events around
        # `iters` back-to-back forwards without the per-launch instrumentation
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(a.iters):
            gen(label, None, fake, prev)
        e1.record()
        torch.cuda.synchronize()
        wall_us = e0.elapsed_time(e1) * 1e3 / a.iters
    lines = ['%-24s %9.1f' % (nm, us) for nm, us in zip(names, acc)]
    lines.append('%-24s %9.1f' % ('total', sum(acc)))
    lines.append('%-24s %9.1f' % ('forward_wall', wall_us))
    txt = '\n'.join(lines)
    print(txt.splitlines()[-1], 'lib', os.environ.get('RIB_LIB', 'default'))
    if a.out:
        open(a.out, 'w').write(txt + '\n')
        open(a.out + '.plan', 'w').write(gen.plan_text())
        open(a.out + '.tune', 'w').write(gen.tune_log())


if __name__ == '__main__':
    main()
