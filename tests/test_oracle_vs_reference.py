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


# ---------------------------------------------------------------- motion model
def _reference_motion_model():
    import sys
    ref = '/root/reference/Human_Motion_Modelling'
    if ref not in sys.path:
        sys.path.insert(0, ref)
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == 'models' or k.startswith('models.')}
    try:
        from models.position_encoding import PositionEmbeddingSine_1D
        from models.transformer import Transformer
    finally:
        for k in [k for k in sys.modules if k == 'models' or k.startswith('models.')]:
            del sys.modules[k]
        sys.modules.update(saved)
        sys.path.remove(ref)
    from oracle import motion_oracle as mo
    c = mo.CFG
    m = Transformer(c['input_joints'], d_model=c['hidden_dim'], nhead=c['nheads'], num_encoder_layers=c['enc_layers'],
                    num_decoder_layers=c['dec_layers'], dim_feedforward=c['dim_feedforward'], dropout=0.1,
                    activation='leaky_relu', normalize_before=True, return_intermediate_dec=False, two_stage=True).eval()
    return m, PositionEmbeddingSine_1D(c['hidden_dim'] // 2, normalize=True)


def test_motion_oracle_equals_reference():
    from oracle import motion_oracle as mo
    m, pe = _reference_motion_model()
    assert [(k, tuple(v.shape)) for k, v in m.state_dict().items()] == [(k, tuple(s)) for k, s in mo.state_spec()]
    sd = mo.synth_state_dict(11)
    m.load_state_dict(sd, strict=True)
    for length, rate in [(17, 4), (65, 8), (321, 16)]:
        data, em, dm = mo.synth_motion(length, rate, seed=length)
        src, sm, tm = data[None], em[None], dm[None]
        pos = pe(sm)
        assert torch.equal(pos, mo.position_encoding(1, length))
        with torch.no_grad():
            j, r = m(src, sm, pos, src.clone(), tm, pe(tm), rate)
        j2, r2 = mo.transformer_forward(sd, src, sm, pos, tm, pos, rate)
        assert (j - j2).abs().max().item() <= 1e-5 and (r - r2).abs().max().item() <= 1e-5
