"""GPU: the input side of the path — the bicubic resize kernel against its oracle and against cv2's
committed outputs, and the folder-driven entry against frames written by the reference's own evaluate_from_folder
(fixture made by oracle/make_golden_folder.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import resize_oracle as rz

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def dev():
    assert torch.cuda.is_available(), 'GPU tests need a CUDA device'
    return torch.device('cuda:0')


def test_resize_kernel_equals_oracle_and_tracks_cv2(dev, golden_dir):
    import rib
    z = np.load(os.path.join(golden_dir, 'resize_cases.npz'))
    for i in range(int(z['n'])):
        img, want = z['in_%d' % i], z['out_%d' % i]
        got = rib.resize_cubic_u8(torch.from_numpy(img)[None].to(dev), want.shape[0], want.shape[1])[0].cpu().numpy()
        assert np.array_equal(got, rz.resize_cubic_u8(img, want.shape[0], want.shape[1])), 'case %d: kernel != oracle' % i
        d = np.abs(got.astype(np.int32) - want.astype(np.int32))       # cv2 4.13 (IPP) outputs committed with the fixture
        assert d.max() <= 1 and (d != 0).mean() <= 5e-4, 'case %d vs cv2' % i
    # batches, down- and up-scaling, exact 2x (ties of the rounding), identity size
    rng = np.random.default_rng(3)
    for (b, h, w, H, W) in [(3, 90, 70, 160, 240), (2, 128, 192, 64, 96), (1, 200, 300, 64, 96), (2, 64, 96, 64, 96)]:
        img = rng.integers(0, 256, (b, h, w, 3), dtype=np.uint8)
        got = rib.resize_cubic_u8(torch.from_numpy(img).to(dev), H, W).cpu().numpy()
        for n in range(b):
            assert np.array_equal(got[n], rz.resize_cubic_u8(img[n], H, W)), (b, h, w, H, W, n)


def _psnr_u8(a, b):
    mse = ((a.astype(np.float64) - b.astype(np.float64)) ** 2).mean() / 255.0 ** 2
    return 10 * np.log10(1.0 / mse) if mse > 0 else float('inf')


def test_evaluate_from_folder_matches_reference_frames(dev, golden_dir, synth_sd, tmp_path):
    """inputs/ DAIN/ Predict_motion/ -> Generated_frames/: same file names as the reference; key frames equal to the
    reference's (cv2-resized) ones within the resize tolerance; generated frames >= 45 dB against the frames the
    reference's CPU loop wrote with the same weights."""
    from PIL import Image
    import rib
    from rib.config import default_gen_cfg
    from rib.folder import clip_layout, evaluate_from_folder
    from rib.generator import Generator
    z = np.load(os.path.join(golden_dir, 'folder_case.npz'))
    sub = str(z['sub'])
    dirs = {k: tmp_path / k / sub for k in ('inputs', 'DAIN', 'Predict_motion')}
    for d in dirs.values():
        d.mkdir(parents=True)
    for n, a in zip(z['key_names'], z['keys']):
        Image.fromarray(a).save(str(dirs['inputs'] / str(n)))
    for n, a in zip(z['dain_names'], z['dain']):
        Image.fromarray(a).save(str(dirs['DAIN'] / str(n)))
    for n, doc in zip(z['pose_names'], z['pose_docs']):
        (dirs['Predict_motion'] / str(n)).write_text(str(doc))
    g = Generator(default_gen_cfg())
    g.load_state_dict(synth_sd, strict=True)
    g = g.to(dev).eval()
    h, w = (int(v) for v in z['model_hw'])
    written = evaluate_from_folder(g, str(tmp_path / 'inputs'), str(tmp_path / 'DAIN'), str(tmp_path / 'Predict_motion'),
                                   str(tmp_path / 'Generated_frames'), height=h, width=w, verbose=False)
    assert list(written) == [sub]
    names = [os.path.basename(p) for p in written[sub]]
    assert names == [str(n) for n in z['out_names']]
    rate, seq_len = clip_layout(len(z['key_names']), len(z['pose_names']))
    assert (rate, seq_len) == (2, 5)
    want = z['out']
    for i, p in enumerate(written[sub]):
        got = np.asarray(Image.open(p))
        assert got.shape == want[i].shape
        if i % rate == 0:      # key frame: cv2 resize -> to_tensor_norm -> tensor2images in the reference
            d = np.abs(got.astype(np.int32) - want[i].astype(np.int32))
            assert d.max() <= 1 and (d != 0).mean() <= 1e-3, 'key frame %d' % i
        else:
            p_db = _psnr_u8(got, want[i])
            assert p_db >= 45.0, 'generated frame %d: %.2f dB' % (i, p_db)
