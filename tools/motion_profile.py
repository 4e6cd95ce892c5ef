#!/usr/bin/env python
"""One forward of the motion Transformer per batch size after cudaProfilerStart (for `ncu --profile-from-start off`)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'render-in-between_b200'))
from oracle import motion_oracle as mo  # noqa: E402
from rib.motion import MotionTransformer  # noqa: E402

dev = torch.device('cuda:0')
m = MotionTransformer(mo.CFG['input_joints'])
m.load_state_dict(mo.synth_state_dict(1), strict=True)
m = m.to(dev).eval()
for batch in [int(a) for a in sys.argv[1:]] or [1, 64]:
    seqs = [mo.synth_motion(321, 16, seed=s) for s in range(batch)]
    src = torch.stack([s[0] for s in seqs]).to(dev)
    sm = torch.stack([s[1] for s in seqs]).to(dev)
    tm = torch.stack([s[2] for s in seqs]).to(dev)
    pos = mo.position_encoding(batch, 321).to(dev)
    m(src, sm, pos, None, tm, pos, 16)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    m(src, sm, pos, None, tm, pos, 16)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
