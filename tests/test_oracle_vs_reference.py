"""CPU, build container only: the oracle restatements against the live, unmodified reference."""
import numpy as np
import pytest
import torch

from oracle import generator_oracle as go
from oracle import raster_oracle as ro
from oracle import ref_import
from rib.synth import synth_image

pytestmark = pytest.mark.skipif(not ref_import.available(), reason='/root/reference not present on this box')


def test_state_dict_spec_matches_reference(arch):
    G = ref_import.make_generator()
    ref = G.state_dict()
    spec = arch.state_spec()
    assert [k for k, _, _ in spec] == list(ref.keys())
    assert all(tuple(ref[k].shape) == s for k, s, _ in spec)
    assert len(spec) == 372


def test_raster_oracle_bit_exact_vs_reference():
    h, w = 80, 112
    ds = ref_import.make_dataset(h, w)
    rng = np.random.default_rng(5)
    for case in range(4):
        lm = [(float(rng.uniform(-4, w + 4)), float(rng.uniform(-4, h + 4))) for _ in range(19)]
        conf = [1.0 if rng.uniform() > 0.1 else 0.0 for _ in range(19)]
        assert np.array_equal(ds._generate_pose_map(lm, conf, h, w), ro.pose_map(lm, conf, h, w))
        assert np.array_equal(ds._generate_skeleton(lm, conf, h, w), ro.skeleton(lm, conf, h, w))


def test_generator_oracle_equals_reference(arch, synth_sd):
    G = ref_import.make_generator()
    G.load_state_dict(synth_sd, strict=True)
    G.eval()
    b, h, w = 1, 32, 48
    label = torch.rand(b, 22, h, w)
    fake, prev = synth_image(b, h, w, 3), synth_image(b, h, w, 4)
    with torch.no_grad():
        ri, rm = G(label, None, fake, prev)
        oi, om = go.generator_forward(synth_sd, arch, label, fake, prev)
    assert torch.equal(ri, oi) and torch.equal(rm, om)


def test_to_uint8_equals_tensor2images():
    ns = ref_import.load()
    x = torch.randn(1, 3, 16, 24)
    assert np.array_equal(ns.tensor2images(x), go.to_uint8(x)[0].numpy())
