#!/usr/bin/env python
"""Runs N generator forwards (B, H, W) on cuda:0 — the command ncu wraps for the per-kernel captures
under profiles/ (B200_PROFILING.md recipe).  No timing is reported from here."""
import argparse
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
    ap.add_argument('--iters', type=int, default=2)
    ap.add_argument('--clip', action='store_true', help='render one bench clip (65 frames, 2x) instead of a bare forward')
    a = ap.parse_args()
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
    if a.clip:
        from rib.clip import ClipRenderer
        from rib.synth import synth_flow
        nkey, rate = b + 1, 2
        t = (nkey - 1) * rate + 1
        key = synth_image(nkey, h, w, seed=0).to(dev)
        joints = torch.from_numpy(synth_joints(t, h, w, seed=0)).to(dev)
        flows = synth_flow(t, h, w, seed=0).to(dev)
        r = ClipRenderer(gen, sample_rate=rate)
        with torch.no_grad():
            r.render(key, joints, flows=flows, want_u8=True, want_fuse=False)
            torch.cuda.synchronize()
            torch.cuda.profiler.start()
            for _ in range(a.iters):
                out = r.render(key, joints, flows=flows, want_u8=True, want_fuse=False)
            torch.cuda.synchronize()
            torch.cuda.profiler.stop()
        print('ok clip', int(out['u8'].sum()))
        return
    label = rib.rasterize(torch.from_numpy(synth_joints(b, h, w, seed=3)).to(dev), h, w)
    fake, prev = synth_image(b, h, w, seed=1).to(dev), synth_image(b, h, w, seed=2).to(dev)
    with torch.no_grad():
        # one untimed forward first (plan build, auto-tuner candidates); with `ncu --profile-from-start off` only the
        # forwards after cudaProfilerStart are seen, so `-s <i>` is the index of a launch in the plan
        gen(label, None, fake, prev)
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        for _ in range(a.iters):
            img, mask = gen(label, None, fake, prev)
            rib.composite(img, mask, fake)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    with open(os.path.join(ROOT, 'gpurun_out', 'plan_B%d_%d.txt' % (b, h)), 'w') as f:
        f.write(gen.plan_text())
    print('ok', float(img.abs().mean()), float(mask.mean()))


if __name__ == '__main__':
    main()
