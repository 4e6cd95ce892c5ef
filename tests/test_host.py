"""CPU: host-side logic and the C-ABI surface (no compute calls)."""
import ctypes
import json
import os
import re

import pytest
import torch

import rib
from rib.arch import Arch
from rib.config import AttrDict, default_gen_cfg
from rib.synth import synth_joints, synth_state_dict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_state_spec_matches_committed_reference_keys(golden_dir, arch):
    ref = json.load(open(os.path.join(golden_dir, 'state_dict_keys.json')))
    spec = arch.state_spec()
    assert [k for k, _, _ in spec] == list(ref.keys())
    assert all(list(s) == ref[k] for k, s, _ in spec)


def test_generator_module_has_reference_state_dict(golden_dir):
    from rib.generator import Generator
    g = Generator(default_gen_cfg())
    ref = json.load(open(os.path.join(golden_dir, 'state_dict_keys.json')))
    sd = g.state_dict()
    assert list(sd.keys()) == list(ref.keys())
    assert all(list(v.shape) == ref[k] for k, v in sd.items())
    assert sum(p.numel() for p in g.parameters()) == 34563733
    g.load_state_dict(synth_state_dict(Arch(default_gen_cfg()), 0), strict=True)


def test_generator_has_no_cpu_path():
    from rib.generator import Generator
    g = Generator(default_gen_cfg()).eval()
    x = torch.zeros(1, 22, 32, 32)
    im = torch.zeros(1, 3, 32, 32)
    with pytest.raises(RuntimeError):
        g(x, None, im, im)
    g.train()
    with pytest.raises(RuntimeError):
        g(x, None, im, im)


def test_config_family_guard():
    cfg = default_gen_cfg()
    cfg.embed.arch = 'unet'
    with pytest.raises(NotImplementedError):
        Arch(cfg)
    a = AttrDict({'a': {'b': 1}})
    assert a.a.b == 1 and vars(a)['a']['b'] == 1


def test_abi_library_exports_every_declared_symbol():
    from rib import _lib
    header = open(os.path.join(ROOT, 'include', 'rib_b200.h')).read()
    declared = set(re.findall(r'\b(rib_[a-z0-9_]+)\s*\(', header))
    declared -= {'rib_gen_config', 'rib_tensor', 'rib_generator'}
    assert declared == set(_lib.SYMBOLS)
    dll = ctypes.CDLL(_lib.LIB_PATH)
    for s in declared:
        assert hasattr(dll, s), s
    assert _lib.lib.rib_abi_version() == 6


def test_synth_is_deterministic():
    a, b = synth_joints(4, 64, 96, seed=3), synth_joints(4, 64, 96, seed=3)
    assert (a == b).all() and a.shape == (4, 19, 3)
    assert ((a[..., :2] % 1.0) != 0).all()    # generic floats, never grid-aligned


def test_tune_table_import_export_roundtrip():
    """The tuning table is plain text (key, mt, policy); import parses without a GPU and export returns it."""
    import ctypes as C
    from rib import _lib
    n = _lib.lib.rib_tune_import(b'unit-test-shape-a\t2\t1\nunit-test-shape-b\t0\t0\nmalformed line\nunit-test-bad\t7\t9\n')
    assert n == 2
    buf = C.create_string_buffer(1 << 20)
    assert _lib.lib.rib_tune_export(buf, len(buf)) == 0
    lines = buf.value.decode().splitlines()
    assert 'unit-test-shape-a\t2\t1\t0' in lines and 'unit-test-shape-b\t0\t0\t0' in lines
    assert not any(l.startswith('unit-test-bad') for l in lines)
    # the shipped table parses completely
    import os
    if os.path.isfile(_lib.TUNE_TABLE):
        text = open(_lib.TUNE_TABLE, 'rb').read()
        assert _lib.lib.rib_tune_import(text + b'\0') == len(text.strip().splitlines())
