"""TEST INFRASTRUCTURE — writes tests/golden/motion_case.npz by running the UNMODIFIED reference motion model
(/root/reference/Human_Motion_Modelling/models/transformer.py `Transformer`, position_encoding.py
`PositionEmbeddingSine_1D`) on seeded synthetic joint sequences with the seeded synthetic weights of
oracle/motion_oracle.synth_state_dict (8 MB of weights are regenerated from the seed, only inputs / outputs are stored).

    python oracle/make_golden_motion.py            (build container only: needs /root/reference)

Cases: the quick-start shape (rate 8), a short clip (rate 4), and one with extra hidden key frames in the encoder mask and
padded frames at the end of the decoder mask (both key_padding_mask paths of nn.MultiheadAttention).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = '/root/reference/Human_Motion_Modelling'

from oracle import motion_oracle as mo  # noqa: E402

WEIGHT_SEED = 3
CASES = [dict(length=33, rate=8, seed=1, hide=[], pad=0), dict(length=17, rate=4, seed=2, hide=[], pad=0),
         dict(length=41, rate=8, seed=3, hide=[16], pad=5), dict(length=9, rate=2, seed=4, hide=[], pad=0)]


def reference_model():
    sys.path.insert(0, REF)
    from models.position_encoding import PositionEmbeddingSine_1D
    from models.transformer import Transformer
    c = mo.CFG
    m = Transformer(c['input_joints'], d_model=c['hidden_dim'], nhead=c['nheads'], num_encoder_layers=c['enc_layers'],
                    num_decoder_layers=c['dec_layers'], dim_feedforward=c['dim_feedforward'], dropout=0.1,
                    activation='leaky_relu', normalize_before=True, return_intermediate_dec=False, two_stage=True).eval()
    return m, PositionEmbeddingSine_1D(c['hidden_dim'] // 2, normalize=True)


def case_inputs(case):
    data, enc_mask, dec_mask = mo.synth_motion(case['length'], case['rate'], seed=case['seed'])
    for i in case['hide']:                       # a key frame the encoder may not look at either
        enc_mask[i] = True
        data[:, i] = 0.0
    if case['pad']:
        dec_mask[-case['pad']:] = True           # padded tail (AMASS_dataset.generate_training_mask convention)
        enc_mask[-case['pad']:] = True
        data[:, -case['pad']:] = 0.0
    return data, enc_mask, dec_mask


def main():
    m, pe = reference_model()
    m.load_state_dict(mo.synth_state_dict(WEIGHT_SEED), strict=True)
    out = {'weight_seed': np.int64(WEIGHT_SEED), 'n_cases': np.int64(len(CASES))}
    for i, case in enumerate(CASES):
        data, em, dm = case_inputs(case)
        src, sm, tm = data[None], em[None], dm[None]
        with torch.no_grad():
            joints, reco = m(src, sm, pe(sm), src.clone(), tm, pe(tm), case['rate'])
        out.update({'c%d_data' % i: data.numpy(), 'c%d_enc_mask' % i: em.numpy(), 'c%d_dec_mask' % i: dm.numpy(),
                    'c%d_rate' % i: np.int64(case['rate']), 'c%d_joints' % i: joints.numpy(), 'c%d_reco' % i: reco.numpy()})
        print('case %d: L=%d rate=%d |joints| max %.3f' % (i, case['length'], case['rate'], joints.abs().max()))
    path = os.path.join(ROOT, 'tests', 'golden', 'motion_case.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, os.path.getsize(path), 'bytes')


if __name__ == '__main__':
    main()
