import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'render-in-between_b200')):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (B200); run with -m gpu')


def pytest_collection_modifyitems(config, items):
    """A plain `pytest tests` on a box without a GPU skips the gpu-marked tests instead of erroring in their fixtures.
    With a GPU present nothing is skipped: a missing or broken extension must fail there, not hide."""
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason='needs a CUDA device (B200); there is no CPU path')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope='session')
def golden_dir():
    return GOLDEN


@pytest.fixture(scope='session')
def arch():
    from rib.arch import Arch
    from rib.config import default_gen_cfg
    return Arch(default_gen_cfg())


@pytest.fixture(scope='session')
def synth_sd(arch):
    from rib.synth import synth_state_dict
    return synth_state_dict(arch, seed=0)
