"""CPU: the launch plan for a sweep of shapes through rib_plan_dry_run (no GPU): every layer's tiling must fit the
shared-memory / TMEM budgets, and the plan for the bench shape must be the one a B200 produced (profiles/)."""
import os
import re

import pytest

from rib.config import default_gen_cfg
from rib.generator import Generator

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SMEM_MAX = 227 * 1024


def _gemm_lines(text):
    out = []
    for line in text.splitlines():
        if line.startswith('gemm '):
            parts = line.split()
            d = {'name': parts[1]}
            d.update({k: v for k, v in (kv.split('=') for kv in parts[2:])})
            out.append(d)
    return out


@pytest.fixture(scope='module')
def gen():
    return Generator(default_gen_cfg())


@pytest.mark.parametrize('b,h,w', [(1, 16, 16), (2, 64, 96), (1, 48, 80), (2, 320, 480), (16, 256, 256), (32, 512, 512),
                                   (48, 512, 512), (4, 512, 768), (1, 1024, 1024), (4, 1024, 1024), (1, 1024, 768), (64, 128, 128)])
def test_plan_dry_run_fits_every_shape(gen, b, h, w):
    need, text = gen.plan_dry_run(b, h, w)
    g = _gemm_lines(text)
    assert len(g) == 75, len(g)                      # 96 reference convs after the fusions of DESIGN.md section 4
    assert need > 0
    for d in g:
        assert int(d['smem']) <= SMEM_MAX, d
        assert int(d['MT']) in (1, 2) and 2 <= int(d['aring']) <= 8, d
        assert 2 * int(d['MT']) * int(d['BN']) <= 512, d       # TMEM columns of the two accumulator buffers
        assert int(d['BKc']) in (16, 32, 64) and int(d['B']) == b, d
    flops = sum(float(d['flops']) for d in g)
    # executed MACs: below the reference's 883.5 kFLOP per pixel (sub-pixel convs run 4/9 of theirs), above 80 % of it
    assert 0.80 * 883.5e3 * b * h * w <= flops <= 1.02 * 883.5e3 * b * h * w, flops / (b * h * w)
    # workspace: activations are 16-bit; the bench shape needs about 17 GB, far below 180 GB of HBM
    assert need < 180e9


@pytest.mark.parametrize('b,h,w', [(1, 40, 40), (0, 64, 64), (1, 8, 8), (1, 64, 72)])
def test_plan_dry_run_rejects_bad_shapes(gen, b, h, w):
    with pytest.raises(RuntimeError):
        gen.plan_dry_run(b, h, w)


def test_plan_dry_run_reproduces_the_measured_plan(gen):
    """The plan text committed from the GPU run (tuned tilings included) is reproduced field by field on the CPU."""
    path = os.path.join(ROOT, 'profiles', 'r2_plan_B32_512.txt')
    if not os.path.isfile(path):
        pytest.skip('no committed plan')
    want = _gemm_lines(open(path).read())
    _, text = gen.plan_dry_run(32, 512, 512)
    got = _gemm_lines(text)
    assert [d['name'] for d in got] == [d['name'] for d in want]
    for a, bb in zip(got, want):
        assert a == bb, (a, bb)
    assert re.findall(r'^(\w+)', text, flags=re.M).count('in_apply') == 13 and 'pack_images' in text
