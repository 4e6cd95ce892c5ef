"""CPU: the motion-model oracle against the committed fixture (made by the reference, oracle/make_golden_motion.py), and the
host mirror's surface (state-dict keys, option checks).  SURVEY.md section 8f rank 3."""
import os

import numpy as np
import pytest
import torch

from oracle import motion_oracle as mo


def _cases(golden_dir):
    z = np.load(os.path.join(golden_dir, 'motion_case.npz'))
    for i in range(int(z['n_cases'])):
        yield (int(z['weight_seed']), torch.from_numpy(z['c%d_data' % i]), torch.from_numpy(z['c%d_enc_mask' % i]),
               torch.from_numpy(z['c%d_dec_mask' % i]), int(z['c%d_rate' % i]), torch.from_numpy(z['c%d_joints' % i]),
               torch.from_numpy(z['c%d_reco' % i]))


def test_motion_oracle_matches_reference_fixture(golden_dir):
    n = 0
    for seed, data, em, dm, rate, joints, reco in _cases(golden_dir):
        sd = mo.synth_state_dict(seed)
        pos = mo.position_encoding(1, data.shape[1])
        j, r = mo.transformer_forward(sd, data[None], em[None], pos, dm[None], pos, rate)
        assert j.shape == joints.shape and torch.isfinite(j).all()
        assert (j - joints).abs().max().item() <= 2e-5, (j - joints).abs().max().item()   # float32, values up to ~6
        assert (r - reco).abs().max().item() <= 2e-5
        assert torch.equal(mo.inference(sd, data, em, dm, rate), j.permute(1, 2, 0))
        n += 1
    assert n == 4


def test_interpolate_embedding_is_the_linear_blend():
    x = torch.randn(17, 1, 38)
    y = mo.interpolate_embedding(x, 4)
    assert torch.equal(y[0::4], x[0::4] / 4 * 4 + x[[4, 8, 12, 16, 16]] / 4 * 0)      # key frames: prev / rate * rate
    assert torch.allclose(y[2], 0.5 * (x[0] + x[4]), atol=1e-6)


def test_host_mirror_state_dict_and_options():
    from rib.motion import MotionTransformer, state_spec
    c = mo.CFG
    spec = state_spec(c['input_joints'], c['hidden_dim'], c['dim_feedforward'], c['enc_layers'], c['dec_layers'])
    assert spec == [(k, tuple(s)) for k, s in mo.state_spec()]
    assert len(spec) == 188
    m = MotionTransformer(c['input_joints'])
    assert [(k, tuple(v.shape)) for k, v in m.state_dict().items()] == spec
    m.load_state_dict(mo.synth_state_dict(0), strict=True)
    for bad in (dict(activation='relu'), dict(normalize_before=False), dict(two_stage=False), dict(return_intermediate_dec=True),
                dict(nhead=4)):
        with pytest.raises(ValueError):
            MotionTransformer(c['input_joints'], **bad)
    with pytest.raises((ValueError, RuntimeError)):                       # no CPU path
        m(torch.zeros(1, 38, 9), None, torch.zeros(9, 1, 128), None, None, torch.zeros(9, 1, 128), 2)


def test_host_position_encoding_equals_oracle():
    """PositionEmbeddingSine_1D (position_encoding.py:26-56) as the host mirror computes it == the oracle (which is pinned to
    the reference module in tests/test_oracle_vs_reference.py), for the lengths the model is used at."""
    from rib.motion import PositionEmbeddingSine1D
    pe = PositionEmbeddingSine1D(mo.CFG['hidden_dim'] // 2)
    for n, length in [(1, 9), (2, 33), (1, 321)]:
        got = pe(torch.zeros(n, length, dtype=torch.bool))
        assert tuple(got.shape) == (length, n, mo.CFG['hidden_dim'])
        assert torch.equal(got, mo.position_encoding(n, length))
