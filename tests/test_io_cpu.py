"""CPU: OpenPose JSON -> joints against the reference's own reader and the committed fixtures."""
import json
import os

import numpy as np
import pytest

from rib import io as rio


def _cases(golden_dir):
    return json.load(open(os.path.join(golden_dir, 'keypoints_cases.json')))['cases']


def test_read_keypoints_matches_golden(golden_dir, tmp_path):
    cases = _cases(golden_dir)
    assert len(cases) >= 9
    for c in cases:
        want = np.asarray(c['expected'], dtype=np.float64)
        got = rio.read_keypoints(c['json'])
        assert got.shape == (19, 3) and got.dtype == np.float64, c['name']
        assert np.array_equal(got, want), c['name']            # bit-exact: same numpy operations in the same order
        path = tmp_path / 'k.json'
        path.write_text(json.dumps(c['json']))
        assert np.array_equal(rio.read_keypoints(str(path)), want), c['name']


def test_read_keypoints_matches_reference_when_present(golden_dir, tmp_path):
    from oracle import ref_import
    if not ref_import.available():
        pytest.skip('reference tree not present (GPU box): covered by the golden fixture')
    ref_import.load()
    from utils.utils import read_json_keypoint
    rng = np.random.default_rng(3)
    for t in range(20):
        people = []
        for _ in range(int(rng.integers(0, 4))):
            body = rng.uniform(0, 500, (25, 3))
            body[:, 2] = rng.uniform(0, 1, 25) * (rng.uniform() > 0.2)
            hl, hr = rng.uniform(0, 500, (21, 3)), rng.uniform(0, 500, (21, 3))
            hl[:, 2] = (rng.uniform(0, 1, 21) > rng.uniform()) * rng.uniform(0, 1, 21)
            hr[:, 2] = (rng.uniform(0, 1, 21) > rng.uniform()) * rng.uniform(0, 1, 21)
            people.append({'pose_keypoints_2d': body.reshape(-1).tolist(), 'hand_left_keypoints_2d': hl.reshape(-1).tolist(),
                           'hand_right_keypoints_2d': hr.reshape(-1).tolist()})
        path = tmp_path / ('r%d.json' % t)
        path.write_text(json.dumps({'people': people}))
        assert np.array_equal(rio.read_keypoints(str(path)), np.asarray(read_json_keypoint(str(path)), dtype=np.float64)), t


def test_scale_and_clip_joints(golden_dir, tmp_path):
    c = _cases(golden_dir)[0]
    k = rio.read_keypoints(c['json'])
    s = rio.scale_keypoints(k, (480, 640), (512, 512))
    assert np.allclose(s[:, 0], k[:, 0] * 512 / 640) and np.allclose(s[:, 1], k[:, 1] * 512 / 480)
    assert np.array_equal(s[:, 2], k[:, 2])
    paths = []
    for i in range(3):
        p = tmp_path / ('f%d.json' % i)
        p.write_text(json.dumps(c['json']))
        paths.append(str(p))
    j = rio.clip_joints(paths, (480, 640), (512, 512))
    assert j.shape == (3, 19, 3) and j.dtype == np.float64 and np.array_equal(j[1], s)


def test_save_frames_roundtrip_and_reference_bytes(tmp_path):
    """PNG is lossless: what save_frames writes decodes to the same pixels as the reference's per-frame
    Image.fromarray(...).save(...) (evaluator.py:265-266), under the reference's file names."""
    from PIL import Image
    rng = np.random.default_rng(0)
    frames = rng.integers(0, 256, size=(5, 48, 64, 3), dtype=np.uint8)
    dain = ['/data/clip/DAIN/%05d.png' % i for i in range(5)]
    names = rio.frame_names(dain, str(tmp_path))
    assert names[3] == os.path.join(str(tmp_path), '00003.png')
    out = rio.save_frames(frames, names, workers=4)
    assert out == names
    for i, n in enumerate(names):
        assert np.array_equal(np.asarray(Image.open(n)), frames[i])
    ref = tmp_path / 'ref.png'
    Image.fromarray(frames[2]).save(str(ref))
    assert ref.read_bytes() == open(names[2], 'rb').read()
    import torch
    rio.save_frames(torch.from_numpy(frames[:2]), names[:2], workers=1)
    with pytest.raises(ValueError):
        rio.save_frames(frames.astype(np.float32), names)


def test_clip_layout_follows_the_reference_formula():
    """sample rate and sequence length from the file counts, evaluator.py:187-191 (surplus poses are ignored)."""
    from rib.folder import clip_layout
    assert clip_layout(3, 5) == (2, 5)
    assert clip_layout(33, 65) == (2, 65)
    assert clip_layout(17, 65) == (4, 65)
    assert clip_layout(3, 6) == (2, 5)          # 2 ** int(log2(5 / 2)) = 2
    assert clip_layout(2, 9) == (8, 9)
    import pytest
    with pytest.raises(ValueError):
        clip_layout(1, 5)
