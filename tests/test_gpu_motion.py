"""GPU parity tests of the motion Transformer (csrc/motion.cu) through the C ABI: against the
fixture written by the reference itself and against the CPU oracle.  Tolerance: float32 arithmetic with another summation
order on values of magnitude ~6 -> max-abs 3e-5 (measured 3-5e-6)."""
import os

import numpy as np
import pytest
import torch

from oracle import motion_oracle as mo

pytestmark = pytest.mark.gpu
TOL = 3e-5


@pytest.fixture(scope='module')
def dev():
    assert torch.cuda.is_available(), 'GPU tests need a CUDA device'
    return torch.device('cuda:0')


def _model(dev, seed):
    from rib.motion import MotionTransformer
    m = MotionTransformer(mo.CFG['input_joints'])
    m.load_state_dict(mo.synth_state_dict(seed), strict=True)
    return m.to(dev).eval()


def _log(text):
    d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'gpurun_out')
    if os.path.isdir(d):
        with open(os.path.join(d, 'parity_report.txt'), 'a') as f:
            f.write(text + '\n')
    print(text)


def test_motion_matches_reference_fixture(dev, golden_dir):
    z = np.load(os.path.join(golden_dir, 'motion_case.npz'))
    m = _model(dev, int(z['weight_seed']))
    for i in range(int(z['n_cases'])):
        data, em, dm = (torch.from_numpy(z['c%d_%s' % (i, k)]) for k in ('data', 'enc_mask', 'dec_mask'))
        rate = int(z['c%d_rate' % i])
        pos = mo.position_encoding(1, data.shape[1]).to(dev)
        j, r = m(data[None].to(dev), em[None].to(dev), pos, None, dm[None].to(dev), pos, rate)
        ej = (j.cpu() - torch.from_numpy(z['c%d_joints' % i])).abs().max().item()
        er = (r.cpu() - torch.from_numpy(z['c%d_reco' % i])).abs().max().item()
        _log('motion fixture case %d (L=%d, rate %d): max|d joints| %.3g  max|d reco| %.3g' % (i, data.shape[1], rate, ej, er))
        assert ej <= TOL and er <= TOL


@pytest.mark.parametrize('length,rate', [(321, 16), (321, 8), (129, 4), (3, 2)], ids=lambda v: str(v))
def test_motion_matches_oracle_batched(dev, length, rate):
    """The longest clip of the configuration (max_seq_length 321) and a batch of three different sequences in one call:
    every sequence equals the oracle run on it alone."""
    m = _model(dev, 5)
    sd = mo.synth_state_dict(5)
    seqs = [mo.synth_motion(length, rate, seed=40 + s) for s in range(3)]
    src = torch.stack([s[0] for s in seqs])
    sm = torch.stack([s[1] for s in seqs])
    tm = torch.stack([s[2] for s in seqs])
    pos = mo.position_encoding(3, length)
    j, r = m(src.to(dev), sm.to(dev), pos.to(dev), None, tm.to(dev), pos.to(dev), rate)
    torch.cuda.synchronize()
    worst = 0.0
    for b in range(3):
        jo, ro_ = mo.transformer_forward(sd, src[b:b + 1], sm[b:b + 1], pos[:, :1], tm[b:b + 1], pos[:, :1], rate)
        worst = max(worst, (j[:, b].cpu() - jo[:, 0]).abs().max().item(), (r[:, b].cpu() - ro_[:, 0]).abs().max().item())
    _log('motion L=%d rate %d batch 3: max-abs error vs the oracle %.3g' % (length, rate, worst))
    assert worst <= TOL
    j2, r2 = m(src.to(dev), sm.to(dev), pos.to(dev), None, tm.to(dev), pos.to(dev), rate)
    assert torch.equal(j, j2) and torch.equal(r, r2)                      # no atomics: bit-reproducible


def test_motion_inference_entry_and_masks(dev):
    """Model_inference.inference (inference.py:21-43) with the host mirror's position encoding; masks given as None equal
    all-visible masks; a sequence length that is not rate * n + 1 is refused as the reference's indexing would fail."""
    from rib.motion import MotionInference, PositionEmbeddingSine1D
    m = _model(dev, 7)
    sd = mo.synth_state_dict(7)
    data, em, dm = mo.synth_motion(33, 8, seed=9)
    pe = PositionEmbeddingSine1D(mo.CFG['hidden_dim'] // 2)
    assert torch.equal(pe(em[None].to(dev)).cpu(), mo.position_encoding(1, 33))
    pred = MotionInference(pe, m).inference(data, None, em, dm, 8)
    assert tuple(pred.shape) == (1, 38, 33)
    assert (pred.cpu() - mo.inference(sd, data, em, dm, 8)).abs().max().item() <= TOL
    pos = mo.position_encoding(1, 33).to(dev)
    none_mask = torch.zeros(1, 33, dtype=torch.bool, device=dev)
    a = m(data[None].to(dev), None, pos, None, None, pos, 8)
    b = m(data[None].to(dev), none_mask, pos, None, none_mask, pos, 8)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    with pytest.raises(RuntimeError):
        m(data[None, :, :32].contiguous().to(dev), None, pos[:32].contiguous(), None, None, pos[:32].contiguous(), 8)
